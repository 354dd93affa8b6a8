"""The reference's own shaders, executed on the CPU -- TEST INFRASTRUCTURE ONLY.

`make -C oracle refshaders` (only where /root/reference exists) rewrites the five GLSL files token-wise into C++
(oracle/glsl2cpp.py), compiles them with g++ against the reference's vendored glm and wraps them in the dispatch loops
of Renderer.cpp (oracle/refshader_*.cpp, oracle/glsl_rt.h).  This module is the ctypes face of the two resulting
libraries; its functions take and return what oracle.cloud / reproject / godrays / tonemap / txaa do, so a test can run
the restatement and the reference text side by side.

  variant "canonical": built-ins as DESIGN.md section 2 fixes them -> the oracle must match bit for bit
  variant "glm":       glm's own built-ins (another reading of the GLSL spec) -> the oracle matches to rounding
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE = Path("/root/reference/src/CloudScapes/shaders/cloudRayMarch.comp")
_LIBS: dict = {}


def _path(variant: str) -> Path:
    return HERE / "_ref" / ("libmeteoros_refshaders.so" if variant == "canonical" else "libmeteoros_refshaders_glm.so")


def available() -> bool:
    """The reference tree is here: the libraries can be (re)built from the shader sources."""
    return REFERENCE.exists()


def built(variant: str = "canonical") -> bool:
    """A prebuilt library travelled here (the GPU box has oracle/_ref/ but no /root/reference)."""
    return _path(variant).exists()


def lib(variant: str = "canonical") -> C.CDLL:
    if variant not in _LIBS:
        if available():
            subprocess.run(["make", "-s", "-C", str(HERE), "refshaders"], check=True, capture_output=True)
        elif not built(variant):
            raise RuntimeError("neither the reference tree nor a prebuilt oracle/_ref library is present on this machine")
        else:
            from . import lib as oracle_lib

            oracle_lib()   # the shader library links the oracle's sampler: make sure libmeteoros_oracle.so exists here
        _LIBS[variant] = C.CDLL(str(_path(variant)))
    return _LIBS[variant]


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def cloud(cam, tm, sky, noise, W, H, hdr=None, mask=None, variant="canonical", group_stride=1, weather_scale=None):
    """One dispatch of cloudRayMarch.comp (1 of 16 pixels, id = tm.frameCountMod16).  hdr / mask are modified in place.
    group_stride > 1 runs only every group_stride-th row of invocations (bench.py's bounded CPU sample).
    weather_scale not None runs the shader with its dead weather-map block revived (canonical variant only)."""
    hdr = np.zeros((H, W, 4), np.float32) if hdr is None else hdr
    mask = np.zeros((H, W, 4), np.float32) if mask is None else mask
    prev = np.zeros((H, W, 4), np.float32)   # bound (set 0, binding 1), never read by the shader
    cam, tm, sky = _c(cam), _c(tm), _c(sky)
    lo, hi, cu, we = (_c(noise[k], np.uint8) for k in ("low", "high", "curl", "weather"))
    entry = lib(variant).mtrefsh_cloud if weather_scale is None else lib("canonical").mtrefsh_cloud_weather
    rc = entry(_p(cam), _p(tm), _p(sky), _p(lo), lo.shape[2], lo.shape[1], lo.shape[0], _p(hi), hi.shape[2], hi.shape[1], hi.shape[0],
               _p(cu), cu.shape[1], cu.shape[0], _p(we), we.shape[1], we.shape[0], W, H, _p(prev), _p(hdr), _p(mask), int(group_stride),
               C.c_float(1.0 if weather_scale is None else weather_scale))
    assert rc == 0
    return {"hdr": hdr, "mask": mask}


def cloud_full(cam, tm, sky, noise, W, H, variant="canonical", hdr=None, mask=None, group_stride=1, weather_scale=None):
    """All 16 pixel ids with the same camera / time: what mtDispatchCloudFull computes."""
    hdr = np.zeros((H, W, 4), np.float32) if hdr is None else hdr
    mask = np.zeros((H, W, 4), np.float32) if mask is None else mask
    t = _c(tm).copy()
    for fid in range(16):
        t["frameCountMod16"] = fid
        cloud(cam, t, sky, noise, W, H, hdr=hdr, mask=mask, variant=variant, group_stride=group_stride, weather_scale=weather_scale)
    return {"hdr": hdr, "mask": mask}


def reproject(cam, cam_old, tm, prev, variant="canonical"):
    H, W, _ = prev.shape
    prev = _c(prev, np.float32)
    cur = np.zeros_like(prev)
    cam, cam_old, tm = _c(cam), _c(cam_old), _c(tm)
    assert lib(variant).mtrefsh_reproject(_p(cam), _p(cam_old), _p(tm), W, H, _p(prev), _p(cur)) == 0
    return cur


def godrays(cam, sky, mask, hdr, variant="canonical"):
    H, W, _ = hdr.shape
    out = _c(hdr, np.float32).copy()
    mask = _c(mask, np.float32)
    cam, sky = _c(cam), _c(sky)
    assert lib(variant).mtrefsh_godrays(_p(cam), _p(sky), W, H, _p(mask), _p(out)) == 0
    return out


def tonemap(tm, hdr, want_f32=False, variant="canonical"):
    H, W, _ = hdr.shape
    hdr = _c(hdr, np.float32)
    ldr = np.zeros((H, W, 4), np.uint8)
    f = np.zeros((H, W, 4), np.float32) if want_f32 else None
    tm = _c(tm)
    assert lib(variant).mtrefsh_tonemap(_p(tm), W, H, _p(hdr), _p(ldr), _p(f)) == 0
    return (ldr, f) if want_f32 else ldr


def txaa(cam, cam_old, tm, cur, prev, want_f32=False, variant="canonical"):
    H, W, _ = cur.shape
    cur, prev = _c(cur, np.uint8), _c(prev, np.uint8)
    out = np.zeros((H, W, 4), np.uint8)
    f = np.zeros((H, W, 4), np.float32) if want_f32 else None
    cam, cam_old, tm = _c(cam), _c(cam_old), _c(tm)
    assert lib(variant).mtrefsh_txaa(_p(cam), _p(cam_old), _p(tm), W, H, _p(cur), _p(prev), _p(out), _p(f)) == 0
    return (out, f) if want_f32 else out
