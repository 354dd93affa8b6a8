#!/usr/bin/env python
"""How many bits of filter position would the light-cone samples need?  Renders three views of the config-5 sweep at 1920x1080 with the
exact path, with the texture-unit mode (MT_FLAG_HW_CONE_FILTER) and with probe builds of the exact kernel whose filter fractions are
rounded to 8 / 9 / 10 bits, and compares each with the exact frame (and the 8-bit emulation with the texture unit's frame).
  for b in 8 9 10; do make -C meteoros_b200/csrc OUT=../../ab_variants/lib_q$b.so EXTRA=-DMT_EMU_WBITS=$b; done
  python tools/probes/emu_weight_bits.py          (on the GPU box; record: profiles/r2c_emu_weight_bits.txt)"""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    from bench import scene_for_view
    from meteoros_b200 import api, textures
    w, h = 1920, 1080
    flags = int(sys.argv[2])
    out = {}
    with api.CloudRenderer(w, h, flags=flags) as r:
        r.upload_noise(textures.load_noise())
        for v in (0, 96, 128):
            cam, tm, sky, tun = scene_for_view(v, w, h, sweep=True)
            r.set_camera(cam); r.set_camera_old(cam); r.set_time(tm); r.set_sun_and_sky(sky); r.set_tuning(tun)
            r.dispatch_cloud_full()
            out[str(v)] = r.read_image(api.IMAGE_CLOUD_CUR)[..., :3]
    np.savez(sys.argv[1], **out)
    sys.exit(0)
runs = {"exact": (None, 0), "hw": (None, 128), "q8": ("ab_variants/lib_q8.so", 0), "q9": ("ab_variants/lib_q9.so", 0), "q10": ("ab_variants/lib_q10.so", 0)}
imgs = {}
for k, (lib, fl) in runs.items():
    env = dict(os.environ)
    if lib: env["METEOROS_B200_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, __file__, f"/tmp/emu_{k}.npz", str(fl)], check=True, env=env)
    imgs[k] = np.load(f"/tmp/emu_{k}.npz")
def rel(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return (np.abs(a - b) / np.maximum(np.abs(b), 1e-6)).max(axis=-1)
for v in ("0", "96", "128"):
    for k in ("hw", "q8", "q9", "q10"):
        e = rel(imgs[k][v], imgs["exact"][v])
        print(v, k, "vs exact: max %.3e p99.9 %.3e over1e-3 %d" % (e.max(), np.quantile(e, 0.999), int((e > 1e-3).sum())))
    e = rel(imgs["q8"][v], imgs["hw"][v])
    print(v, "q8 vs hw: max %.3e p99.9 %.3e" % (e.max(), np.quantile(e, 0.999)))
