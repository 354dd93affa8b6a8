"""CPU parity oracle for meteoros_b200 -- TEST INFRASTRUCTURE ONLY (see oracle/meteoros_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this package.
Pinned to the reference's own shader text where /root/reference exists (oracle/refshaders.py,
tests/test_reference_shaders.py); what a Vulkan driver would add on top stays unpinned.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libmeteoros_oracle.so"


class MtoTextures(C.Structure):
    _fields_ = [
        ("low", C.c_void_p), ("low_w", C.c_int32), ("low_h", C.c_int32), ("low_d", C.c_int32),
        ("high", C.c_void_p), ("high_w", C.c_int32), ("high_h", C.c_int32), ("high_d", C.c_int32),
        ("curl", C.c_void_p), ("curl_w", C.c_int32), ("curl_h", C.c_int32),
        ("weather", C.c_void_p), ("weather_w", C.c_int32), ("weather_h", C.c_int32),
    ]


class MtCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("rays", "rays_marched", "steps", "steps_incloud", "cone_hits", "early_exits")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


RAY_DEBUG_DTYPE = np.dtype(
    [("dir", "<f4", (3,)), ("t_in", "<f4"), ("t_out", "<f4"), ("step_size", "<f4"), ("branch", "<i4"), ("steps", "<i4"),
     ("jitter_hash", "<u4"), ("accum", "<f4")]
)

_lib = None


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < (HERE / "meteoros_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(str(LIB))
        except OSError:  # built on another machine with an incompatible toolchain: rebuild here
            build(force=True)
            _lib = C.CDLL(str(LIB))
        _lib.mto_cloud.restype = C.c_int
        _lib.mto_reproject.restype = C.c_int
        _lib.mto_godrays.restype = C.c_int
        _lib.mto_tonemap.restype = C.c_int
        _lib.mto_wang_hash.restype = C.c_uint32
        _lib.mto_wang_hash.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _textures(noise):
    keep = {k: np.ascontiguousarray(noise[k], dtype=np.uint8) for k in ("low", "high", "curl")}
    t = MtoTextures()
    t.low, (t.low_d, t.low_h, t.low_w) = keep["low"].ctypes.data, keep["low"].shape[:3]
    t.high, (t.high_d, t.high_h, t.high_w) = keep["high"].ctypes.data, keep["high"].shape[:3]
    t.curl, (t.curl_h, t.curl_w) = keep["curl"].ctypes.data, keep["curl"].shape[:2]
    if "weather" in noise and noise["weather"] is not None:
        keep["weather"] = np.ascontiguousarray(noise["weather"], dtype=np.uint8)
        t.weather, (t.weather_h, t.weather_w) = keep["weather"].ctypes.data, keep["weather"].shape[:2]
    return t, keep


def cloud(cam, tm, tuning, noise, W, H, full=False, rows=None, hdr=None, mask=None, counters=False, debug=False,
          group_stride=1):
    """cloudRayMarch.comp over the reference grid.  Returns dict(hdr, mask[, counters][, debug]).
    hdr / mask may be passed in (float32 HxWx4, modified in place: unwritten pixels keep their value)."""
    cam = np.ascontiguousarray(cam); tm = np.ascontiguousarray(tm); tuning = np.ascontiguousarray(tuning)
    hdr = np.zeros((H, W, 4), np.float32) if hdr is None else hdr
    mask = np.zeros((H, W, 4), np.float32) if mask is None else mask
    assert hdr.dtype == np.float32 and hdr.flags.c_contiguous and mask.dtype == np.float32 and mask.flags.c_contiguous
    tex, keep = _textures(noise)
    cnt = MtCounters()
    dbg = np.zeros((H, W), RAY_DEBUG_DTYPE) if debug else None
    r0, r1 = rows if rows is not None else (0, H)
    rc = lib().mto_cloud(_p(cam), _p(tm), _p(tuning), C.byref(tex), C.c_int(W), C.c_int(H), C.c_int(int(bool(full))),
                         C.c_int(r0), C.c_int(r1), C.c_int(group_stride), _p(hdr), _p(mask), C.byref(cnt) if counters else None, _p(dbg))
    if rc != 0:
        raise ValueError("mto_cloud: invalid arguments")
    del keep
    out = {"hdr": hdr, "mask": mask}
    if counters:
        out["counters"] = cnt.as_dict()
    if debug:
        out["debug"] = dbg
    return out


def reproject(cam, cam_old, tm, prev, taps=False):
    H, W, _ = prev.shape
    prev = np.ascontiguousarray(prev, dtype=np.float32)
    cur = np.zeros_like(prev)
    t = np.zeros((H, W, 10), np.int32) if taps else None
    cam = np.ascontiguousarray(cam); cam_old = np.ascontiguousarray(cam_old); tm = np.ascontiguousarray(tm)
    rc = lib().mto_reproject(_p(cam), _p(cam_old), _p(tm), C.c_int(W), C.c_int(H), _p(prev), _p(cur), _p(t))
    if rc != 0:
        raise ValueError("mto_reproject: invalid arguments")
    return (cur, t) if taps else cur


def godrays(cam, sky, mask, hdr):
    """Returns a new HDR image = hdr + god rays."""
    H, W, _ = hdr.shape
    out = np.ascontiguousarray(hdr, dtype=np.float32).copy()
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    cam = np.ascontiguousarray(cam); sky = np.ascontiguousarray(sky)
    rc = lib().mto_godrays(_p(cam), _p(sky), C.c_int(W), C.c_int(H), _p(mask), _p(out))
    if rc != 0:
        raise ValueError("mto_godrays: invalid arguments")
    return out


def tonemap(tm, hdr, want_f32=False):
    H, W, _ = hdr.shape
    hdr = np.ascontiguousarray(hdr, dtype=np.float32)
    ldr = np.zeros((H, W, 4), np.uint8)
    f = np.zeros((H, W, 4), np.float32) if want_f32 else None
    tm = np.ascontiguousarray(tm)
    rc = lib().mto_tonemap(_p(tm), C.c_int(W), C.c_int(H), _p(hdr), _p(ldr), _p(f))
    if rc != 0:
        raise ValueError("mto_tonemap: invalid arguments")
    return (ldr, f) if want_f32 else ldr


def txaa(cam, cam_old, tm, cur, prev, want_f32=False):
    H, W, _ = cur.shape
    cur = np.ascontiguousarray(cur, dtype=np.uint8)
    prev = np.ascontiguousarray(prev, dtype=np.uint8)
    out = np.zeros((H, W, 4), np.uint8)
    f = np.zeros((H, W, 4), np.float32) if want_f32 else None
    cam = np.ascontiguousarray(cam); cam_old = np.ascontiguousarray(cam_old); tm = np.ascontiguousarray(tm)
    lib().mto_txaa.restype = C.c_int
    rc = lib().mto_txaa(_p(cam), _p(cam_old), _p(tm), C.c_int(W), C.c_int(H), _p(cur), _p(prev), _p(out), _p(f))
    if rc != 0:
        raise ValueError("mto_txaa: invalid arguments")
    return (out, f) if want_f32 else out


def sample3d(vol, s, t, r):
    vol = np.ascontiguousarray(vol, dtype=np.uint8)
    d, h, w, _ = vol.shape
    out = (C.c_float * 4)()
    lib().mto_sample3d(_p(vol), C.c_int(w), C.c_int(h), C.c_int(d), C.c_float(s), C.c_float(t), C.c_float(r), out)
    return np.array(out[:], np.float32)


def sample2d(img, s, t):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    out = (C.c_float * 4)()
    lib().mto_sample2d(_p(img), C.c_int(w), C.c_int(h), C.c_float(s), C.c_float(t), out)
    return np.array(out[:], np.float32)


def wang_hash(u, v, s):
    return int(lib().mto_wang_hash(u & 0xFFFFFFFF, v & 0xFFFFFFFF, s & 0xFFFFFFFF))


def encode_float_rgba(v):
    out = (C.c_float * 4)()
    lib().mto_encode_float_rgba(C.c_float(v), out)
    return np.array(out[:], np.float32)


def ray_sphere(ro, rd, c, radius):
    f3 = C.c_float * 3
    pt, t, valid = f3(), C.c_float(), C.c_int()
    lib().mto_ray_sphere(f3(*ro), f3(*rd), f3(*c), C.c_float(radius), pt, C.byref(t), C.byref(valid))
    return np.array(pt[:], np.float32), np.float32(t.value), bool(valid.value)


def density_height_gradient(relative_height, cloud_type):
    f = lib().mto_density_height_gradient
    f.restype = C.c_float
    return float(f(C.c_float(relative_height), C.c_float(cloud_type)))


def cloud_grid(W, H):
    tx, ty = C.c_int(), C.c_int()
    lib().mto_cloud_grid(C.c_int(W), C.c_int(H), C.byref(tx), C.byref(ty))
    return tx.value, ty.value


def atmosphere_color(direction, sun_minus_origin, sun_intensity, sky_sun):
    f3 = C.c_float * 3
    out = f3()
    lib().mto_atmosphere_color(f3(*direction), f3(*sun_minus_origin), C.c_float(sun_intensity), f3(*sky_sun), out)
    return np.array(out[:], np.float32)
