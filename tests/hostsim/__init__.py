"""Host build of the kernels' per-pixel cores (TEST TOOL; see hostsim.cpp)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libhostsim.so"
CSRC = HERE.parents[1] / "meteoros_b200" / "csrc"


def build():
    srcs = [HERE / "hostsim.cpp"] + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h"))
    if LIB.exists() and all(LIB.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return LIB
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-x", "c++",
           str(HERE / "hostsim.cpp"), "-o", str(LIB)]
    if "fma" in Path("/proc/cpuinfo").read_text().split():
        cmd.insert(1, "-mfma")
    subprocess.run(cmd, check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def cloud(cam, tm, tun, noise, W, H, full, debug_dtype):
    hdr = np.zeros((H, W, 4), np.float32)
    mask = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(6, np.uint64)
    dbg = np.zeros((H, W), debug_dtype)
    cam, tm, tun = (np.ascontiguousarray(x) for x in (cam, tm, tun))
    lo, hi, cu, we = noise["low"], noise["high"], noise["curl"], noise["weather"]
    lib().hs_cloud(_p(cam), _p(tm), _p(tun), _p(lo), lo.shape[2], lo.shape[1], lo.shape[0], _p(hi), hi.shape[2], hi.shape[1],
                   hi.shape[0], _p(cu), cu.shape[1], cu.shape[0], _p(we), we.shape[1], we.shape[0], W, H, int(full), _p(hdr),
                   _p(mask), _p(cnt), _p(dbg))
    keys = ("rays", "rays_marched", "steps", "steps_incloud", "cone_hits", "early_exits")
    return hdr, mask, dict(zip(keys, (int(v) for v in cnt))), dbg


def reproject(cam, cam_old, tm, prev):
    H, W, _ = prev.shape
    cur = np.zeros_like(prev)
    taps = np.zeros((H, W, 10), np.int32)
    cam, cam_old, tm = (np.ascontiguousarray(x) for x in (cam, cam_old, tm))
    lib().hs_reproject(_p(cam), _p(cam_old), _p(tm), W, H, _p(prev), _p(cur), _p(taps))
    return cur, taps


def godrays(cam, light_color, mask, hdr):
    H, W, _ = hdr.shape
    out = hdr.copy()
    lc = np.ascontiguousarray(light_color, dtype=np.float32)
    cam = np.ascontiguousarray(cam)
    lib().hs_godrays(_p(cam), _p(lc), W, H, _p(mask), _p(out))
    return out


def tonemap(tm, hdr):
    H, W, _ = hdr.shape
    ldr = np.zeros((H, W), np.uint32)
    tm = np.ascontiguousarray(tm)
    lib().hs_tonemap(_p(tm), W, H, _p(hdr), _p(ldr))
    return ldr.view(np.uint8).reshape(H, W, 4)


def txaa(cam, cam_old, tm, cur, prev):
    H, W, _ = cur.shape
    out = np.zeros((H, W), np.uint32)
    cam, cam_old, tm = (np.ascontiguousarray(x) for x in (cam, cam_old, tm))
    cur = np.ascontiguousarray(cur); prev = np.ascontiguousarray(prev)
    lib().hs_txaa(_p(cam), _p(cam_old), _p(tm), W, H, _p(cur), _p(prev), _p(out))
    return out.view(np.uint8).reshape(H, W, 4)
