#!/usr/bin/env python
"""Device time of whole frames of the reference loop (mtFrameEx: reprojection + 1-of-16 Cloud + god rays + tone map [+ TXAA]), WITHOUT
per-pass events -- so that the passes mtFrameEx runs side by side do.  CUDA events around `--frames` consecutive warm frames.
usage: python tools/frame_time.py [--width 1920 --height 1080] [--txaa] [--no-godrays]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from meteoros_b200 import api, scene, textures  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--txaa", action="store_true")
    ap.add_argument("--no-godrays", action="store_true")
    a = ap.parse_args()
    w, h = a.width, a.height
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    with api.CloudRenderer(w, h) as r:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        old = cam.ubo()
        ubos = []
        for f in range(16 + a.frames):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            ubos.append((cam.ubo(), old, sc.ubo()))
            old = cam.ubo()
        for f, (c, o, t) in enumerate(ubos):
            if f == 16:
                r.event_record(0)
            r.set_camera(c); r.set_camera_old(o); r.set_time(t)
            r.frame(not a.no_godrays, a.txaa)
        r.event_record(1)
        r.synchronize()
        ms = r.event_elapsed_ms(0, 1)
    what = "reproject + cloud 1/16" + ("" if a.no_godrays else " + god rays") + " + tone map" + (" + TXAA" if a.txaa else "")
    print(f"frame {1e3 * ms / a.frames:9.1f} us   ({w}x{h}, {what}; {a.frames} warm frames)")


if __name__ == "__main__":
    main()
