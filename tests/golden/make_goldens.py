#!/usr/bin/env python
"""Mint the frame fixtures under tests/golden/ from the REFERENCE'S OWN SHADERS.

Runs only where /root/reference exists.  oracle/refshaders.py compiles the five GLSL files from their own text (g++,
the reference's vendored glm, canonical built-ins -- see oracle/glsl_rt.h) and dispatches them as Renderer.cpp does;
this script drives them with the reference's uniform producers' values and stores what they write.  The oracle is run
beside them and must agree bit for bit before anything is saved (the per-ray debug records, which only the oracle
has, are stored too).  The fixtures travel to the GPU box; the reference does not.

  cloud_64x36.npz          one full-quality cloud frame (all 16 pixel ids): HDR, god-ray mask, per-ray records
  sequence_96x54.npz       4 frames of REPROJ, CLOUD, GODRAYS, TONEMAP with a 0.25 degree pan (main.cpp:172-194)
  live_sequence_96x54.npz  16 frames of all five shaders chained: REPROJ, CLOUD, GODRAYS, TONEMAP, TXAA
  cloud_fullsize_digests.npz  the full-quality cloud frame at BASELINE's own sizes (1920x1080 and 3840x2160) as digests small enough
                           to commit: CRC-32 of every pixel row of the god-ray mask and of alpha (bit-exact quantities), every
                           8th / 16th pixel of the HDR colour, and float64 row sums of the HDR colour
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from conftest import default_scene  # noqa: E402
from meteoros_b200 import scene, textures  # noqa: E402
from oracle import refshaders  # noqa: E402

SOURCE = "reference shaders compiled from their own text (oracle/refshaders.py, canonical built-ins); oracle bit-identical"


def same(a, b, what):
    if not np.array_equal(a, b):
        raise SystemExit(f"oracle and reference shader disagree on {what}: refusing to write fixtures")


def frame_loop(n_frames, w, h, noise, with_txaa):
    """The reference's frame loop (main.cpp:172-194) on the reference shaders, the oracle run beside it."""
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    mask = np.zeros((h, w, 4), np.float32)
    hist = np.zeros((h, w, 4), np.uint8)     # TXAA history (previous presented frame)
    cur, cam_old = 0, cam.ubo()
    out = {"hdr": [], "ldr": [], "txaa": []}
    for k in range(n_frames):
        cam.rotate_about_up(0.25)
        sc.update_time(1 / 60)
        c, t, s = cam.ubo(), sc.ubo(), sky.ubo()
        o_img = oracle.reproject(c, cam_old, t, img[cur ^ 1])
        img[cur] = refshaders.reproject(c, cam_old, t, img[cur ^ 1])
        same(img[cur], o_img, f"frame {k} reprojection")
        o_mask = mask.copy()
        oracle.cloud(c, t, tun, noise, w, h, full=False, hdr=o_img, mask=o_mask)
        refshaders.cloud(c, t, s, noise, w, h, hdr=img[cur], mask=mask)
        same(img[cur], o_img, f"frame {k} cloud HDR"); same(mask, o_mask, f"frame {k} mask")
        o_img = oracle.godrays(c, s, mask, img[cur])
        img[cur] = refshaders.godrays(c, s, mask, img[cur])
        same(img[cur], o_img, f"frame {k} god rays")
        ldr = refshaders.tonemap(t, img[cur])
        same(ldr, oracle.tonemap(t, img[cur]), f"frame {k} tone map")
        out["hdr"].append(img[cur].copy()); out["ldr"].append(ldr)
        if with_txaa:
            aa = refshaders.txaa(c, cam_old, t, ldr, hist)
            same(aa, oracle.txaa(c, cam_old, t, ldr, hist), f"frame {k} TXAA")
            out["txaa"].append(aa)
            hist = aa
        cur ^= 1
        cam_old = c
    out["mask"] = mask
    return out


def frame_digests(hdr, mask, sub):
    """What tests/test_gpu_fullsize.py::test_reference_shader_digests_at_full_size recomputes from the CUDA frame."""
    import zlib

    return {
        "mask_crc": np.array([zlib.crc32(np.ascontiguousarray(mask[y]).tobytes()) for y in range(mask.shape[0])], np.uint32),
        "alpha_crc": np.array([zlib.crc32(np.ascontiguousarray(hdr[y, :, 3]).tobytes()) for y in range(hdr.shape[0])], np.uint32),
        "hdr_sub": np.ascontiguousarray(hdr[::sub, ::sub, :3]),
        "hdr_rowsum": hdr[..., :3].astype(np.float64).sum(axis=1),
    }


def main():
    if not refshaders.available():
        raise SystemExit("the reference tree is not present: fixtures can only be minted where /root/reference exists")
    noise = textures.load_noise()
    gold = ROOT / "tests" / "golden"

    w, h = 64, 36
    frame_id, total_time, yaw = 3, 2.5, 1.5
    cam, tm, sky, tun = default_scene(w, h, frame_id=frame_id, total_time=total_time, yaw=yaw)
    ref = refshaders.cloud_full(cam, tm, sky, noise, w, h)
    r = oracle.cloud(cam, tm, tun, noise, w, h, full=True, debug=True)
    same(ref["hdr"], r["hdr"], "cloud_64x36 HDR"); same(ref["mask"], r["mask"], "cloud_64x36 mask")
    np.savez_compressed(gold / "cloud_64x36.npz", hdr=ref["hdr"], mask=ref["mask"], steps=r["debug"]["steps"],
                        jitter_hash=r["debug"]["jitter_hash"], accum=r["debug"]["accum"], frame_id=frame_id,
                        total_time=total_time, yaw=yaw, source=SOURCE)

    s = frame_loop(4, 96, 54, noise, with_txaa=False)
    np.savez_compressed(gold / "sequence_96x54.npz", ldr=np.stack(s["ldr"]), hdr=np.stack(s["hdr"]), source=SOURCE)

    s = frame_loop(16, 96, 54, noise, with_txaa=True)
    np.savez_compressed(gold / "live_sequence_96x54.npz", txaa=np.stack(s["txaa"]), ldr=np.stack(s["ldr"]),
                        hdr_last=s["hdr"][-1], hdr_first=s["hdr"][0], mask_last=s["mask"], source=SOURCE)
    out = {"source": SOURCE}
    for (w, h, sub) in ((1920, 1080, 8), (3840, 2160, 16)):
        cam, tm, sky, tun = default_scene(w, h)
        ref = refshaders.cloud_full(cam, tm, sky, noise, w, h)
        r = oracle.cloud(cam, tm, tun, noise, w, h, full=True)
        same(ref["hdr"], r["hdr"], f"{w}x{h} HDR"); same(ref["mask"], r["mask"], f"{w}x{h} mask")
        for k, v in frame_digests(ref["hdr"], ref["mask"], sub).items():
            out[f"{k}_{w}x{h}"] = v
        out[f"sub_{w}x{h}"] = sub
    np.savez_compressed(gold / "cloud_fullsize_digests.npz", **out)
    print("goldens written from the reference shaders")


if __name__ == "__main__":
    main()
