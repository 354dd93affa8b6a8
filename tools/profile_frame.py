#!/usr/bin/env python
"""Small driver for ncu: a few launches of each pass at a chosen size (default 3840x2160 full-quality cloud pass).

  ncu --set full --clock-control none --import-source on -k regex:cloud_raymarch -s 1 -c 1 -o gpurun_out/prof \
      python tools/profile_frame.py --passes cloud
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from meteoros_b200 import api, scene, textures  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--passes", default="cloud", help="comma list of cloud,cloud16,reproject,godrays,tonemap,frame")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--flags", type=int, default=0, help="MtConfig.flags of the context (e.g. 128 = MT_FLAG_HW_CONE_FILTER)")
    a = ap.parse_args()
    w, h = a.width, a.height
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    sc.update_time(1 / 60)
    with api.CloudRenderer(w, h, flags=a.flags) as r:
        r.upload_noise(textures.load_noise())
        r.set_camera(cam.ubo()); r.set_camera_old(cam.ubo()); r.set_time(sc.ubo()); r.set_sun_and_sky(sky.ubo())
        for _ in range(a.reps):
            for p in a.passes.split(","):
                {"cloud": r.dispatch_cloud_full, "cloud16": r.dispatch_cloud, "reproject": r.dispatch_reprojection,
                 "godrays": r.dispatch_god_rays, "tonemap": r.dispatch_tone_map, "frame": lambda: r.frame(True)}[p]()
            r.synchronize()


if __name__ == "__main__":
    main()
