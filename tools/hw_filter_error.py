#!/usr/bin/env python
"""What MT_FLAG_HW_CONE_FILTER (the opt-in texture-unit mode) costs in radiance across the 256-view sweep of BASELINE config 5: for
a few views (sun elevation x coverage), the full-quality 1920x1080 frame of the mode against the default path's frame -- max relative
error, pixels beyond the 1e-3 bar, mask / alpha equality -- and both device times.   python tools/hw_filter_error.py  (on the GPU box)"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import scene_for_view  # noqa: E402
from meteoros_b200 import api, textures  # noqa: E402


def main():
    w, h = 1920, 1080
    noise = textures.load_noise()
    views = [0, 15, 96, 127, 128, 143, 240, 255]
    out = {}
    rs = {f: api.CloudRenderer(w, h, flags=f) for f in (0, api.FLAG_HW_CONE_FILTER)}
    for r in rs.values():
        r.upload_noise(noise)
    for v in views:
        cam, tm, sky, tun = scene_for_view(v, w, h, sweep=True)
        img, ms = {}, {}
        for f, r in rs.items():
            r.set_camera(cam); r.set_camera_old(cam); r.set_time(tm); r.set_sun_and_sky(sky); r.set_tuning(tun)
            r.dispatch_cloud_full()
            r.event_record(0)
            r.dispatch_cloud_full()
            r.event_record(1)
            ms[f] = r.event_elapsed_ms(0, 1)
            img[f] = (r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK))
        (ex, exm), (hw, hwm) = img[0], img[api.FLAG_HW_CONE_FILTER]
        a, b = hw[..., :3].astype(np.float64), ex[..., :3].astype(np.float64)
        rel = (np.abs(a - b) / np.maximum(np.abs(b), 1e-6)).max(axis=-1)
        out[v] = {"coverage": round(float(tun["coverage"]), 3), "ms_exact": round(ms[0], 4), "ms_hw": round(ms[api.FLAG_HW_CONE_FILTER], 4),
                  "max_rel_err": float(f"{rel.max():.3e}"), "pixels_over_1e-3": int((rel > 1e-3).sum()),
                  "p99.9_rel_err": float(f"{np.quantile(rel, 0.999):.3e}"),
                  "mask_equal": bool(np.array_equal(hwm, exm)), "alpha_equal": bool(np.array_equal(hw[..., 3], ex[..., 3]))}
        print(v, json.dumps(out[v]), flush=True)
    for r in rs.values():
        r.close()


if __name__ == "__main__":
    main()
