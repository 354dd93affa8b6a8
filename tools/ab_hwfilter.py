#!/usr/bin/env python
"""A/B: the light-cone samples through the texture unit's trilinear filter (build -DMT_HW_FILTER=1, ab_variants/lib_hwfilter.so)
against the default exact fp32 filter: device time of the 4K full-quality Cloud pass, and what the 9-bit filter weights do to the
frame (max relative error, PSNR, pixels beyond the 1e-3 bar, god-ray mask / alpha equality).
  python tools/ab_hwfilter.py            (on the GPU box; runs itself once per library)"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def render(out):
    import numpy as np
    from conftest import default_scene
    from meteoros_b200 import api, textures

    w, h = 3840, 2160
    cam, tm, _, tun = default_scene(w, h)
    with api.CloudRenderer(w, h) as r:
        r.upload_noise(textures.load_noise())
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        for _ in range(3):
            r.dispatch_cloud_full()
        ms = []
        for _ in range(10):
            r.flush_l2(0)
            r.event_record(0)
            r.dispatch_cloud_full()
            r.event_record(1)
            ms.append(r.event_elapsed_ms(0, 1))
        np.savez(out, hdr=r.read_image(api.IMAGE_CLOUD_CUR), mask=r.read_image(api.IMAGE_GODRAY_MASK), ms=np.array(ms))


def main():
    if len(sys.argv) > 1:
        render(sys.argv[1])
        return
    import numpy as np

    outs = {}
    for name, lib in (("exact", None), ("hw", ROOT / "ab_variants" / "lib_hwfilter.so")):
        env = dict(os.environ)
        if lib:
            env["METEOROS_B200_LIB"] = str(lib)
        out = f"/tmp/ab_hw_{name}.npz"
        subprocess.run([sys.executable, __file__, out], check=True, env=env)
        outs[name] = np.load(out)
    a, b = outs["exact"], outs["hw"]
    ref, got = a["hdr"][..., :3].astype(np.float64), b["hdr"][..., :3].astype(np.float64)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)
    mse = np.mean((got - ref) ** 2)
    res = {
        "ms_exact": round(float(np.mean(a["ms"])), 4), "ms_hw_filter": round(float(np.mean(b["ms"])), 4),
        "hdr_max_rel_err": float(rel.max()), "hdr_psnr_db": float(10 * np.log10(ref.max() ** 2 / mse)) if mse else float("inf"),
        "pixels_over_1e-3": int((rel.max(axis=-1) > 1e-3).sum()), "pixels": int(rel.shape[0] * rel.shape[1]),
        "pixels_differing": int((rel.max(axis=-1) > 0).sum()),
        "mask_equal": bool(np.array_equal(a["mask"], b["mask"])), "alpha_equal": bool(np.array_equal(a["hdr"][..., 3], b["hdr"][..., 3])),
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
