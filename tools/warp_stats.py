#!/usr/bin/env python
"""Lane occupancy of the full-quality Cloud pass by phase, replayed on the host (tools/probes/warp_stats.cpp).
  python tools/warp_stats.py [W H [row_stride [tile_w tile_h]]]"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import default_scene  # noqa: E402
from meteoros_b200 import textures  # noqa: E402


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("warps", "warps_marching", "warp_iters", "lane_iters", "warp_iters_hit", "lane_hits")] + [
        (k, C.c_uint64 * 6) for k in ("cone_warp_iters", "cone_lanes", "cone_nonempty_warp", "cone_nonempty_lanes", "cone_hit_lanes")] + [
        ("hist_hits", C.c_uint64 * 33), ("early_lane_idle", C.c_uint64), ("tail_lane_idle", C.c_uint64), ("nonempty_total_hist", C.c_uint64 * 193)]


def main():
    a = [int(x) for x in sys.argv[1:]]
    W, H = (a + [3840, 2160])[:2] if len(a) >= 2 else (3840, 2160)
    stride = a[2] if len(a) > 2 else 8
    tw, th = (a[3], a[4]) if len(a) > 4 else (16, 2)
    so = Path("/tmp/libwarp_stats.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off", "-x", "c++",
                    str(ROOT / "tools" / "probes" / "warp_stats.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    noise = textures.load_noise()
    cam, tm, _, tun = default_scene(W, H)
    cam, tm, tun = (np.ascontiguousarray(x) for x in (cam, tm, tun))
    lo, hi, cu = noise["low"], noise["high"], noise["curl"]
    p = lambda x: C.c_void_p(x.ctypes.data)
    S = Stats()
    lib.ws_run(p(cam), p(tm), p(tun), p(lo), lo.shape[2], lo.shape[1], lo.shape[0], p(hi), hi.shape[2], hi.shape[1], hi.shape[0],
               p(cu), cu.shape[1], cu.shape[0], W, H, tw, th, stride, C.byref(S))
    print(f"{W}x{H}, {tw}x{th} ray tiles, every {stride}th tile row")
    print(f"warps {S.warps}, marching {S.warps_marching}")
    print(f"march iterations: {S.warp_iters} warp, {S.lane_iters} lane  -> {S.lane_iters / S.warp_iters:.2f} lanes / iteration")
    print(f"  idle lane-iterations: early exit {S.early_lane_idle} ({S.early_lane_idle / (32 * S.warp_iters):.3f}), tail/culled {S.tail_lane_idle} ({S.tail_lane_idle / (32 * S.warp_iters):.3f})")
    print(f"iterations with >=1 in-cloud lane: {S.warp_iters_hit} ({S.warp_iters_hit / S.warp_iters:.3f}); in-cloud lanes {S.lane_hits} -> {S.lane_hits / max(S.warp_iters_hit, 1):.2f} lanes / such iteration")
    for i in range(6):
        print(f"  cone {i}: executed {S.cone_warp_iters[i]} warp-iters, {S.cone_lanes[i]} lanes; non-empty cell: {S.cone_nonempty_warp[i]} warp-iters, "
              f"{S.cone_nonempty_lanes[i]} lanes ({S.cone_nonempty_lanes[i] / max(S.cone_lanes[i], 1):.3f}) -> {S.cone_nonempty_lanes[i] / max(S.cone_nonempty_warp[i], 1):.2f} lanes/filter; density>0: {S.cone_hit_lanes[i]}")
    h = np.array(S.hist_hits[:], dtype=np.float64)
    print("in-cloud lanes per iteration, histogram (0..32):", " ".join(f"{int(v)}" for v in h))
    ne = np.array(S.nonempty_total_hist[:], dtype=np.float64)
    tot_ne = float((ne * np.arange(193)).sum())
    batches = float((ne * np.ceil(np.arange(193) / 32.0)).sum())
    print(f"non-empty cone samples per in-cloud iteration: mean {tot_ne / max(ne.sum(), 1):.1f}; filter batches if redistributed per iteration: {batches:.0f} "
          f"vs {sum(S.cone_nonempty_warp[:])} executed today; ideal (queue across iterations) {tot_ne / 32:.0f}")


if __name__ == "__main__":
    main()
