#!/usr/bin/env python
"""Writes the oracle-generated regression fixtures under tests/golden/ (NOT reference vectors: the reference
cannot be run).  Re-run only when the canonical semantics change on purpose."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from conftest import default_scene  # noqa: E402
from meteoros_b200 import scene, textures  # noqa: E402


def main():
    noise = textures.load_noise()
    w, h = 64, 36
    frame_id, total_time, yaw = 3, 2.5, 1.5
    cam, tm, _, tun = default_scene(w, h, frame_id=frame_id, total_time=total_time, yaw=yaw)
    r = oracle.cloud(cam, tm, tun, noise, w, h, full=True, debug=True)
    np.savez_compressed(ROOT / "tests/golden/cloud_64x36.npz", hdr=r["hdr"], mask=r["mask"], steps=r["debug"]["steps"],
                        jitter_hash=r["debug"]["jitter_hash"], accum=r["debug"]["accum"], frame_id=frame_id,
                        total_time=total_time, yaw=yaw)

    # 4-frame pan with the full frame loop (REPROJ, CLOUD, GODRAYS, TONEMAP, swap), main.cpp:172-194
    w, h = 96, 54
    cam = scene.Camera(w, h)
    sc, sky = scene.Scene(), scene.Sky()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    mask = np.zeros((h, w, 4), np.float32)
    cur = 0
    cam_old = cam.ubo()
    ldrs, hdrs = [], []
    for _ in range(4):
        cam.rotate_about_up(0.25)
        sc.update_time(1 / 60)
        c, t = cam.ubo(), sc.ubo()
        img[cur] = oracle.reproject(c, cam_old, t, img[cur ^ 1])
        oracle.cloud(c, t, tun, noise, w, h, full=False, hdr=img[cur], mask=mask)
        img[cur] = oracle.godrays(c, sky.ubo(), mask, img[cur])
        ldrs.append(oracle.tonemap(t, img[cur]))
        hdrs.append(img[cur].copy())
        cur ^= 1
        cam_old = c
    np.savez_compressed(ROOT / "tests/golden/sequence_96x54.npz", ldr=np.stack(ldrs), hdr=np.stack(hdrs))
    print("goldens written")


if __name__ == "__main__":
    main()
