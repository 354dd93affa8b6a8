#!/usr/bin/env python
"""Soak: thousands of frames of the reference loop through mtFrameEx with resizes, storage formats, flag sets and context churn in
between; device memory in use must come back to where it started and every call must succeed.   python tools/soak.py [--frames 3000]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from meteoros_b200 import api, scene, textures  # noqa: E402


def used_mb():
    free, total = torch.cuda.mem_get_info(0)
    return (total - free) / 2**20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3000)
    a = ap.parse_args()
    torch.cuda.init()
    noise = textures.load_noise()
    sky = scene.Sky()
    sizes = [(1920, 1080), (1284, 720), (640, 360), (130, 70), (2560, 1440)]
    configs = [(0, 0), (2, 0), (0, api.FLAG_HW_CONE_FILTER), (1, api.FLAG_NO_FUSED_TONEMAP), (0, api.FLAG_TOP_DOWN | api.FLAG_NO_CONE_RF)]
    base = None
    done = 0
    rng = np.random.default_rng(0)
    for round_ in range(len(configs) * 2):
        storage, flags = configs[round_ % len(configs)]
        w, h = sizes[0]
        cam, sc = scene.Camera(w, h), scene.Scene()
        with api.CloudRenderer(w, h, storage=storage, flags=flags) as r:
            r.upload_noise(noise)
            r.set_sun_and_sky(sky.ubo())
            old = cam.ubo()
            per = a.frames // (len(configs) * 2)
            for f in range(per):
                if f and f % 97 == 0:  # resize in the middle of a sequence
                    w, h = sizes[int(rng.integers(len(sizes)))]
                    r.resize(w, h)
                    cam = scene.Camera(w, h)
                    old = cam.ubo()
                if f % 211 == 0:
                    tun = scene.default_tuning()
                    tun["coverage"] = float(0.3 + 0.6 * rng.random())
                    r.set_tuning(tun)
                cam.rotate_about_up(0.25)
                sc.update_time(1 / 60)
                r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
                r.frame(with_godrays=(f % 3 != 0), with_txaa=(f % 2 == 0))
                old = cam.ubo()
                if f % 53 == 0:
                    r.dispatch_cloud_full()
                    img = r.read_image(api.IMAGE_CLOUD_CUR)
                    assert np.isfinite(img.astype(np.float32)).all()
                done += 1
            r.synchronize()
            ldr = r.read_image(api.IMAGE_LDR_PREV)
            assert ldr[..., :3].any()
        torch.cuda.synchronize()
        m = used_mb()
        if base is None:
            base = m
        print(f"round {round_}: storage {storage} flags {flags}: {done} frames so far, device memory in use after close {m:.0f} MiB (first round {base:.0f})", flush=True)
        assert abs(m - base) < 64, "device memory did not come back"
    print("soak ok")


if __name__ == "__main__":
    main()
