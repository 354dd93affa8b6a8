#!/usr/bin/env python
"""Warm device time of every pass of one reference frame (CUDA events on the context's stream, MT_FLAG_PASS_TIMING).
usage: python tools/pass_times.py [--width 1920 --height 1080] [--sequential]"""
import argparse
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from meteoros_b200 import api, scene, textures  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--sequential", action="store_true")
    ap.add_argument("--ctx-flags", type=int, default=0)
    a = ap.parse_args()
    w, h = a.width, a.height
    flags = api.FLAG_PASS_TIMING | (api.FLAG_SEQUENTIAL_MARCH if a.sequential else 0) | a.ctx_flags
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    names = ["reproject", "cloud(1/16)", "godrays", "tonemap", "txaa"]
    t = {n: [] for n in names}
    with api.CloudRenderer(w, h, flags=flags) as r:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        old = cam.ubo()
        for f in range(a.frames):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
            r.frame(True, True)
            old = cam.ubo()
            if f >= 16:
                for k, n in enumerate(names):
                    t[n].append(r.last_pass_ms(k))
    tot = 0.0
    for n in names:
        m = statistics.median(t[n])
        tot += m
        print(f"{n:12s} {1e3 * m:9.1f} us")
    print(f"{'frame':12s} {1e3 * tot:9.1f} us   ({w}x{h}, {'sequential' if a.sequential else 'step-parallel'} 1/16 march)")


if __name__ == "__main__":
    main()
