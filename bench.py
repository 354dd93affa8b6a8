#!/usr/bin/env python
"""bench.py -- headline benchmark of the cloud hot path (BASELINE.json: raymarched Mrays/s at 3840x2160).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cloud4k|frame8k|seq1080p|views256]

A step = one full-quality Cloud pass (mtDispatchCloudFull: all sixteen pixel ids, max steps, full light cone, no
reprojection) over one synthetic 3840x2160 frame of the default cloudscape = 8 294 400 rays (BASELINE config 3).
  value  device-timed Mrays/s: CUDA events on the context's stream around each dispatch, inputs resident in HBM,
         L2 evicted between steps (mtFlushL2 writes 256 MiB), summed over exactly K steps, max over ranks.
  e2e    the same metric through the public C-ABI call sequence with host buffers: uniforms from host memory in; the
         RGBA32F HDR frame AND the god-ray image (one float per pixel, mtReadGodRayGreyAsync) read back into pinned host
         memory, every step, wall-clock between synchronisations.  e2e_f16: the same with RGBA16F images (MT_STORAGE_F16,
         the reference's own image format).
  N > 1  weak scaling, no data-path collective: every rank renders its own copy of the 4K frame (independent views,
         BASELINE config 5 style; --sweep varies sun elevation / coverage per rank); value = N * rays / max-over-ranks time.
         The same line carries `sharded_8k` (and `sharded_8k_f16`): ONE 7680x4320 frame cut into cyclic 8-row tiles over the
         ranks, HDR tiles stored straight into GPU 0's image over NVLink by the march kernel (BASELINE config 4), timed
         against the same frame on rank 0 alone in the same run, gathered frame compared bit for bit with the single-GPU
         frame.  `--workload frame8k` runs only that, with --gather / --tile-rows variants.
  seq1080p / views256 (sub-records of the default line)  BASELINE configs 2 and 5 beside the headline: ms per frame of the
         1920x1080 16-frame pan through mtFrame (reprojection + 1-of-16 Cloud dispatch + god rays + tone map) -- the second
         half of BASELINE.json's metric -- and the 256-view sun / coverage sweep.  --workload seq1080p / views256 run them alone.
  --impl reference   the reference's Cloud shader compiled for the CPU from its own text (oracle/_ref; else the oracle
         restatement), OpenMP over all host cores (team size set explicitly and reported as measured), on a bounded,
         evenly spread sample of the same frame: the reference's own path needs Vulkan + a window.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W4K, H4K = 3840, 2160
METRIC = "raymarched Mrays/s at 3840x2160 (device-timed)"
UNIT = "Mrays/s"

# Algorithmic FP32 operations per unit of work (DESIGN.md "Work model"; fma = 2, div/sqrt/exp/pow = 1), counted
# from the canonical arithmetic of oracle/meteoros_oracle.c.
FLOP_PER_RAY = 40.0            # castRay + horizon test (every ray)
FLOP_PER_MARCHED_RAY = 300.0   # Preetham sky, two shell intersections, phase function, composite, mask encode
FLOP_PER_STEP = 182.0          # jittered position, height, wind skew + 1 filtered RGBA low-frequency sample (112)
FLOP_PER_INCLOUD_STEP = 936.0  # curl + high-frequency erosion (127), 6 cone samples (6 x 127), light energy (47)
FLOP_PER_CONE_HIT = 4.0        # erosion remap of a cone sample with density > 0
HBM_BYTES_PER_RAY = 32.0       # RGBA32F colour + RGBA32F god-ray mask, written once


def algorithmic_flop(c: dict) -> float:
    return (FLOP_PER_RAY * c["rays"] + FLOP_PER_MARCHED_RAY * c["rays_marched"] + FLOP_PER_STEP * c["steps"]
            + FLOP_PER_INCLOUD_STEP * c["steps_incloud"] + FLOP_PER_CONE_HIT * c["cone_hits"])


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture."""
    f = ROOT / "profiles" / "cloud_raymarch_traffic.json"
    try:
        return int(json.loads(f.read_text())["dram_bytes_per_launch"])
    except (OSError, ValueError, KeyError):
        return None


def filtered_fetches(c: dict) -> int:
    return c["steps"] + 8 * c["steps_incloud"]  # S + 8C


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML is polled
    from a thread every 2 ms (a 4K step is 5 ms: `nvidia-smi -lms` is too slow to land a sample inside a short region);
    nvidia-smi is the fallback when the NVML binding is unusable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device: int):
        self.device, self.rows, self.proc, self.thread = device, [], None, None
        self.nvml, self.handle, self.stop_flag, self.sm_max, self.reason_mask = None, None, threading.Event(), None, 0

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001 -- any NVML problem: fall back to the CLI
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:  # noqa: BLE001
                    pw = None
                try:
                    self.reason_mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                except Exception:  # noqa: BLE001
                    pass
                self.rows.append((sm, pw))
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.002)

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = [r[0] for r in self.rows]
            pw = [r[1] for r in self.rows if r[1] is not None]
            reasons = [k for k, bit in self.REASON_BITS.items() if self.reason_mask & bit]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max, "power_w_max": round(max(pw), 2) if pw else None,
                    "samples": len(self.rows), "reasons": reasons, "source": "nvml, 2 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons, "source": "nvidia-smi -lms 25"}


def scene_for_view(view: int, w: int, h: int, sweep: bool = False):
    """View 0 = the default cloudscape; views > 0 sweep sun elevation (5..85 deg) x coverage (0.3..0.9), config 5."""
    from meteoros_b200 import scene

    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    sc.update_time(1.0 / 60.0)
    tun = scene.default_tuning()
    if view > 0 or sweep:
        tun["sun_location"] = scene.sun_on_elevation_circle(5.0 + 80.0 * ((view % 16) / 15.0))
        tun["coverage"] = 0.3 + 0.6 * (((view // 16) % 16) / 15.0)
    return cam.ubo(), sc.ubo(), sky.ubo(), tun


def cpu_reference_sample(noise, w, h, target_s=12.0, probe_stride=64):
    """Times the reference's Cloud pass on the host cores, on every `stride`-th 4-row group of the frame (evenly spread,
    so ocean / sky / cloud rows are sampled in proportion).  kind "reference": the reference's OWN shader, compiled from
    its text into oracle/_ref (built where /root/reference exists; the library travels to the GPU box); kind "port": the
    oracle restatement, when that library is absent.  Returns (Mrays/s, cores, description, seconds, kind)."""
    import oracle
    from oracle import refshaders

    cam, tm, sky, tun = scene_for_view(0, w, h)
    hdr = np.zeros((h, w, 4), np.float32)
    mask = np.zeros((h, w, 4), np.float32)
    # torchrun exports OMP_NUM_THREADS=1 to every rank: ask for every core this process may run on, and report the team
    # size OpenMP really forms (one libgomp per process, so this also governs oracle/_ref's loops)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    olib = oracle.lib()
    olib.mto_set_num_threads(int(avail))
    cores = int(olib.mto_num_threads())
    assert cores == avail or avail == 1, f"OpenMP formed a team of {cores} threads, {avail} cores are available"
    groups = (h + 3) // 4
    kind = "reference" if (refshaders.available() or refshaders.built()) else "port"

    def run(stride):
        t0 = time.perf_counter()
        if kind == "reference":   # 16 dispatches of cloudRayMarch.comp, as Renderer.cpp:701-716 issues each of them
            refshaders.cloud_full(cam, tm, sky, noise, w, h, hdr=hdr, mask=mask, group_stride=stride)
            rays = len(range(0, groups, stride)) * 4 * w
        else:
            rays = oracle.cloud(cam, tm, tun, noise, w, h, full=True, hdr=hdr, mask=mask, counters=True, group_stride=stride)["counters"]["rays"]
        return time.perf_counter() - t0, rays

    run(max(probe_stride * 4, 1))  # warm the library / page in the volumes
    dt, rays = run(probe_stride)
    est_full = dt * probe_stride
    stride = int(min(max(round(est_full / target_s), 1), groups // 8))
    dt, rays = run(stride)
    what = "reference cloudRayMarch.comp compiled from its own text (oracle/_ref, g++ + glm)" if kind == "reference" else "oracle port"
    desc = f"every {stride}th 4-row group of the {w}x{h} full-quality frame ({rays} rays), {what}, OpenMP x{cores}"
    return rays / dt / 1e6, cores, desc, dt, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's Cloud shader on the host cores (oracle/_ref, else the oracle port); rank 0 only."""
    if rank != 0:
        return
    from meteoros_b200 import textures

    noise = textures.load_noise()
    w, h = (W4K, H4K)
    vals, secs, desc, cores, kind = [], [], "", 1, "port"
    budget = 150.0 / max(args.steps + args.warmup, 1)
    for i in range(args.warmup + args.steps):
        v, cores, desc, dt, kind = cpu_reference_sample(noise, w, h, target_s=min(12.0, budget))
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * statistics.mean(secs), 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3840x2160 full-quality Cloud pass, default cloudscape (BASELINE config 3)", "sample": desc},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference application needs Vulkan + a window and ships no SPIR-V (SURVEY.md 8c); the timed arm is its Cloud "
                "shader compiled for the CPU from its own text (kind reference) or, without oracle/_ref, the oracle port (kind port)",
    }
    emit(line)  # the process's real stdout (quiet_stdout rerouted fd 1 to stderr for library chatter)


def views256_record(args, api, sharding, torch, dist, world, rank, local_rank, noise, batches=3):
    """BASELINE config 5 beside the headline: 256 full-quality 1920x1080 views -- 16 sun elevations x 16 coverages -- round-robin
    over the ranks, no communication.  Device-timed per batch (events on each rank's stream), max over ranks.  Every change of
    coverage rebuilds the empty-cell bitmap (and the cell flags inside the (r, F) bricks) on the device, inside the timed region."""
    w, h = 1920, 1080
    cam, tm, sky, _ = scene_for_view(0, w, h)
    mine = [scene_for_view(v, w, h, sweep=True)[3] for v in sharding.views_of_rank(256, world, rank)]
    r = api.CloudRenderer(w, h, device=local_rank, storage=args.storage)
    r.upload_noise(noise)
    r.set_camera(cam); r.set_camera_old(cam); r.set_time(tm); r.set_sun_and_sky(sky)

    def batch():
        for vt in mine:
            r.set_tuning(vt)
            r.dispatch_cloud_full()

    batch()
    r.synchronize()
    ms = []
    for _ in range(batches):
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
        r.event_record(4)
        batch()
        r.event_record(5)
        ms.append(r.event_elapsed_ms(4, 5))
    r.close()
    t_ms = float(statistics.mean(ms))
    if dist is not None:
        t = torch.tensor([t_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t[0])
    if rank != 0:
        return None
    return {"workload": "256 full-quality 1920x1080 views: 16 sun elevations (5..85 deg) x 16 coverages (0.3..0.9), round-robin over the ranks (BASELINE config 5)",
            "ms_per_batch": round(t_ms, 3), "ms_per_view": round(t_ms / (256 / world), 4), "mrays_per_s": round(256 * w * h / (t_ms * 1e-3) / 1e6, 1),
            "views_per_rank": 256 // world, "batches": batches, "scaling": "strong"}


def seq1080p_record(args, api, torch, dist, world, rank, local_rank, noise, frames=16, repeats=8):
    """BASELINE config 2 beside the headline ("ms/frame at 1080p" is the second half of BASELINE.json's metric): the reference's
    frame loop (main.cpp:172-194) at 1920x1080 -- a 16-frame camera pan, RotateAboutUp(0.25 deg) and time += 1/60 per frame, frame
    ids 1..15, 0 -- through mtFrame: reprojection + 1-of-16 Cloud dispatch + god rays + tone map.  Whole frames, device-timed with
    events around each 16-frame pass of the pan (no per-pass events, so the passes mtFrameEx runs side by side do); the first pass
    warms up, the mean of the next `repeats` passes is reported.  Every rank runs the same sequence (replicas); max over ranks."""
    from meteoros_b200 import scene as _scene
    w, h = 1920, 1080
    cam, sc, sky = _scene.Camera(w, h), _scene.Scene(), _scene.Sky()
    r = api.CloudRenderer(w, h, device=local_rank, storage=args.storage)
    r.upload_noise(noise)
    r.set_sun_and_sky(sky.ubo())
    old = cam.ubo()
    ms = []
    for rep in range(repeats + 1):
        ubos = []
        for _ in range(frames):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            ubos.append((cam.ubo(), old, sc.ubo()))
            old = cam.ubo()
        r.event_record(4)
        for c, o, t in ubos:
            r.set_camera(c); r.set_camera_old(o); r.set_time(t)
            r.frame(True, False)
        r.event_record(5)
        r.synchronize()
        if rep:
            ms.append(r.event_elapsed_ms(4, 5) / frames)
    launches = r.launch_count()
    r.close()
    t_ms = float(statistics.mean(ms))
    if dist is not None:
        t = torch.tensor([t_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t[0])
    if rank != 0:
        return None
    return {"workload": "1920x1080 16-frame camera pan through mtFrame: reprojection + 1-of-16 Cloud dispatch + god rays + tone map (BASELINE config 2)",
            "ms_per_frame": round(t_ms, 4), "frames_per_s": round(1e3 / t_ms, 1), "frames_timed": frames * repeats,
            "mrays_per_s": round((w // 4) * (h // 4) / (t_ms * 1e-3) / 1e6, 1),
            "launches_per_frame": round(launches / (frames * (repeats + 1)), 2),
            "timing": "CUDA events around each 16-frame pass of the pan, warm (L2 not flushed: consecutive frames of one sequence); per-pass times and their rooflines: --workload seq1080p"}


def hw_cone_filter_record(args, api, local_rank, noise, cam, tm, sky, steps=10):
    """The opt-in texture-unit mode beside the headline (MT_FLAG_HW_CONE_FILTER, meteoros_b200.h): the same 3840x2160 full-quality
    step with the six light-cone samples filtered by the texture unit (CUDA 3D texture object, 8-bit weights) instead of the exact
    fp32 filter, timed like the headline (events on the context's stream, L2 flushed between steps), and what it costs in radiance
    against the default path's frame.  NOT the headline and not a parity path: reported so that the texture-unit question
    (north_star: "where the L1/tex path wins") has a measured answer in every bench line.  Rank 0 only."""
    w, h = W4K, H4K
    frames = {}
    ms = None
    for flags in (0, api.FLAG_HW_CONE_FILTER):
        r = api.CloudRenderer(w, h, device=local_rank, storage=0, flags=flags)
        r.upload_noise(noise)
        r.set_camera(cam); r.set_camera_old(cam); r.set_time(tm); r.set_sun_and_sky(sky)
        r.dispatch_cloud_full()
        r.synchronize()
        if flags:
            t = []
            for _ in range(steps):
                r.flush_l2(0)
                r.event_record(4)
                r.dispatch_cloud_full()
                r.event_record(5)
                t.append(r.event_elapsed_ms(4, 5))
            ms = float(statistics.mean(t))
        frames[flags] = (r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK))
        r.close()
    (ex, exm), (hw, hwm) = frames[0], frames[api.FLAG_HW_CONE_FILTER]
    a, b = hw[..., :3].astype(np.float64), ex[..., :3].astype(np.float64)
    rel = (np.abs(a - b) / np.maximum(np.abs(b), 1e-6)).max(axis=-1)
    mse = float(((a - b) ** 2).mean())
    peak = float(b.max())
    return {"mode": "MT_FLAG_HW_CONE_FILTER (opt-in): light-cone samples through a CUDA 3D texture object, the texture unit's 8-bit filter weights",
            "ms_per_step": round(ms, 4), "mrays_per_s": round(w * h / (ms * 1e-3) / 1e6, 1), "steps": steps,
            "vs_exact_path": {"hdr_max_rel_err": float(f"{rel.max():.3e}"), "pixels_over_1e-3": int((rel > 1e-3).sum()), "pixels": w * h,
                              "pixels_differing": int((rel > 0).sum()),
                              "psnr_db": round(10.0 * np.log10(peak * peak / mse), 1) if mse > 0 else None,
                              "mask_equal": bool(np.array_equal(hwm, exm)), "alpha_equal": bool(np.array_equal(hw[..., 3], ex[..., 3]))},
            "note": "outside the 1e-3 parity bar on a few pixels: never the default, never the headline value"}


def sharded_8k_record(args, api, sharding, torch, dist, world, rank, local_rank, noise, steps, storage=None):
    """BASELINE config 4 beside the N>1 views line: ONE 7680x4320 full-quality frame cut into cyclic row tiles over the
    ranks and gathered on GPU 0 over NVLink, timed like the headline (events on each rank's stream, L2 flushed, max over
    ranks) against the same frame on rank 0 alone, measured in the same run; the gathered frame is compared bit for bit
    with the single-GPU frame.  Returns the record on rank 0 (None elsewhere)."""
    w, h = 7680, 4320
    storage = args.storage if storage is None else storage
    cam, tm, sky, tun = scene_for_view(0, w, h)
    r8 = api.CloudRenderer(w, h, device=local_rank, storage=storage)
    r8.upload_noise(noise)
    r8.set_camera(cam); r8.set_camera_old(cam); r8.set_time(tm); r8.set_sun_and_sky(sky); r8.set_tuning(tun)

    def sync_all():
        r8.synchronize()
        torch.cuda.synchronize()
        dist.barrier()

    single_ms, ref_frame = 0.0, None
    if rank == 0:  # the whole frame on one GPU: the denominator of the speed-up
        for _ in range(3):
            r8.dispatch_cloud_full()
        ms = []
        for _ in range(steps):
            r8.flush_l2(0)
            r8.event_record(2)
            r8.dispatch_cloud_full()
            r8.event_record(3)
            ms.append(r8.event_elapsed_ms(2, 3))
        single_ms = float(statistics.mean(ms))
        ref_frame = r8.read_image(api.IMAGE_CLOUD_CUR)
        r8.clear_images()  # a tile that never arrives must not pass the comparison below
    sync_all()
    shard = sharding.ShardedFrame(r8, dist, tile_rows=args.tile_rows, with_mask=args.gather_mask, mode=args.gather)
    for _ in range(3):
        shard.dispatch()
        shard.finish()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = []
    for _ in range(steps):
        r8.flush_l2(0)
        dist.barrier()  # every rank starts the frame together
        r8.event_record(2)
        shard.dispatch()
        if args.gather in ("copy", "forward"):
            r8.join_copies()  # the interval ends when this rank's tile pushes have landed
        r8.event_record(3)
        shard.finish()
        ms.append(r8.event_elapsed_ms(2, 3))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([float(statistics.mean(ms))], dtype=torch.float64, device=f"cuda:{local_rank}")
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(per_rank, t)
    rec = None
    if rank == 0:
        got = r8.read_image(api.IMAGE_CLOUD_CUR)
        frame_ms = max(float(x[0]) for x in per_rank)
        rec = {
            "workload": f"7680x4320 full-quality frame, cyclic {args.tile_rows}-row tiles over {world} GPUs, HDR gathered on GPU 0 over NVLink (BASELINE config 4)",
            "ms_per_frame": round(frame_ms, 4), "single_gpu_ms": round(single_ms, 4), "speedup": round(single_ms / frame_ms, 3),
            "mrays_per_s": round(w * h / (frame_ms * 1e-3) / 1e6, 1), "rank_ms": [round(float(x[0]), 4) for x in per_rank],
            "gather": args.gather, "gather_mask": bool(args.gather_mask), "tile_rows": args.tile_rows, "steps": steps, "scaling": "strong",
            "bit_identical_to_single_gpu": bool(np.array_equal(got, ref_frame)),
            "storage": {0: "RGBA32F", 1: "RGBA32F holding binary16 values", 2: "RGBA16F (the reference's own image format, Renderer.cpp:1431-1440)"}[storage],
            "gathered_bytes": int(w * h * (8 if storage == 2 else 16) * (world - 1) // world), "clocks": clocks,
        }
    shard.close()
    r8.close()
    return rec


def bind_to_gpu_numa_node(device: int):
    """One process per GPU on a two-socket host: run this rank on the CPUs next to its GPU (NVML's ideal CPU set) BEFORE any
    pinned host buffer is allocated, so that first touch places the buffers on the GPU's own NUMA node.  Eight ranks reading
    133-265 MB per step back into one node's memory are limited by that node (~90 GB/s in aggregate, round 1); spread over
    both nodes each GPU keeps its own PCIe link busy.  Best effort: returns a description or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs near GPU {device} ({min(cpus)}..{max(cpus)})"
    except Exception:  # noqa: BLE001 -- no NVML, no affinity support: keep the default placement
        return None


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's real stdout; everything libraries print (NCCL's version banner, torchrun
    notices) was rerouted to stderr by quiet_stdout()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cloud4k", choices=["cloud4k", "frame8k", "seq1080p", "views256"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ctx-flags", type=int, default=0, help="extra MT_FLAG_* bits for A/B runs (e.g. 8 = no quad layout)")
    ap.add_argument("--tile-rows", type=int, default=8, help="frame8k: pixel rows per cyclic tile (multiple of 8)")
    ap.add_argument("--gather", default="peer_store", choices=["peer_store", "bulk_store", "copy", "forward", "local"],
                    help="frame8k: kernel peer stores (default), copy-engine tile pushes, a tile-forwarding side kernel (correct, not yet "
                         "timed at 8 GPUs), or (diagnostic) no gather at all")
    ap.add_argument("--gather-mask", action="store_true", help="frame8k: also send the god-ray mask tiles to GPU 0 (needed only if god rays run)")
    ap.add_argument("--sweep", action="store_true", help="N>1: rank r renders view r of the sun/coverage sweep")
    ap.add_argument("--no-views256", action="store_true", help="default workload: skip the views256 sub-record (config 5)")
    ap.add_argument("--no-seq1080p", action="store_true", help="default workload: skip the seq1080p sub-record (config 2)")
    ap.add_argument("--no-hw-filter", action="store_true", help="default workload: skip the hw_cone_filter sub-record (opt-in texture-unit mode)")
    ap.add_argument("--no-sharded-8k", action="store_true", help="N>1 default workload: skip the sharded_8k sub-record (config 4)")
    ap.add_argument("--storage", type=int, default=0, help="image storage: 0 = RGBA32F (default), 1 = binary16-rounded values in RGBA32F")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    from meteoros_b200 import api, sharding, textures

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    noise = textures.load_noise()

    if args.workload == "frame8k":
        w, h, workload = 7680, 4320, f"7680x4320 full-quality frame, cyclic {args.tile_rows}-row tiles over the ranks, HDR tiles {'stored straight into GPU 0 by the march kernel' if args.gather == 'peer_store' else 'pushed to GPU 0 by the copy engine behind the next tile group'} over NVLink (BASELINE config 4)"
    elif args.workload == "seq1080p":
        w, h, workload = 1920, 1080, "1920x1080 16-frame pan: Reprojection + 1/16 Cloud + god rays + tone map (BASELINE config 2)"
    elif args.workload == "views256":
        w, h, workload = 1920, 1080, "256 full-quality 1920x1080 views: 16 sun elevations x 16 coverages, round-robin over the ranks (BASELINE config 5)"
    else:
        w, h, workload = W4K, H4K, "3840x2160 full-quality Cloud pass, all 16 pixel ids, no reprojection (BASELINE config 3)"

    # weak scaling wants identical work per GPU: every rank renders the default view (--sweep gives each rank its own
    # sun-elevation / coverage view as in BASELINE config 5, whose cost varies with coverage)
    view = rank if (args.workload == "cloud4k" and world > 1 and args.sweep) else 0
    cam, tm, sky, tun = scene_for_view(view, w, h)

    # ---- work accounting (untimed): exact counters of this rank's frame from the counting variant of the kernel
    with api.CloudRenderer(w, h, device=local_rank, flags=api.FLAG_COUNTERS) as rc:
        rc.upload_noise(noise)
        rc.set_camera(cam); rc.set_time(tm); rc.set_tuning(tun)
        rc.dispatch_cloud_full()
        counters = rc.counters()
        fp32_peak_gflops = rc.measure_fp32_peak_gflops()

    r = api.CloudRenderer(w, h, device=local_rank, storage=args.storage, flags=(api.FLAG_PASS_TIMING if args.workload == "seq1080p" else 0) | args.ctx_flags)
    r.upload_noise(noise)
    r.set_camera(cam); r.set_camera_old(cam); r.set_time(tm); r.set_sun_and_sky(sky); r.set_tuning(tun)
    from meteoros_b200 import scene as _scene
    pan_cam, pan_scene = _scene.Camera(w, h), _scene.Scene()   # seq1080p: the reference's frame loop (main.cpp:172-194)
    pan_old = [pan_cam.ubo()]
    pass_ms = {"reproject": [], "cloud_1of16": [], "godrays": [], "tonemap": []}

    shard = None
    if args.workload == "frame8k":
        class _Dist:  # single-process stand-in so N=1 runs the same code path
            def get_rank(self): return 0
            def get_world_size(self): return 1
        shard = sharding.ShardedFrame(r, dist if world > 1 else _Dist(), tile_rows=args.tile_rows, with_mask=args.gather_mask, mode=args.gather)

    seq_state = {"frame": 0}
    my_views = [scene_for_view(v, w, h, sweep=True)[3] for v in sharding.views_of_rank(256, world, rank)] if args.workload == "views256" else []

    def step():
        if args.workload == "cloud4k":
            r.dispatch_cloud_full()
        elif args.workload == "frame8k":
            shard.dispatch()
        elif args.workload == "views256":
            for vt in my_views:  # every view: new sun position + coverage (the empty-cell bitmap is rebuilt per coverage)
                r.set_tuning(vt)
                r.dispatch_cloud_full()
        else:  # one reference frame: 0.25 deg pan, dt = 1/60, ids 1..15,0, REPROJ + CLOUD + GODRAYS + TONEMAP + swap
            pan_cam.rotate_about_up(0.25)
            pan_scene.update_time(1.0 / 60.0)
            c = pan_cam.ubo()
            r.set_camera(c); r.set_camera_old(pan_old[0]); r.set_time(pan_scene.ubo())
            r.frame(with_godrays=True)
            pan_old[0] = c
            seq_state["frame"] += 1

    def barrier():
        r.synchronize()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    rays_per_step_rank = counters["rays"] if args.workload != "seq1080p" else counters["rays"] // 16
    if args.workload == "frame8k" and world > 1:
        rays_total_per_step = w * h
    elif args.workload == "views256":
        rays_total_per_step = 256 * w * h
    else:
        rays_total_per_step = rays_per_step_rank * world

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- device-timed region: exactly K steps, events on the launching stream, L2 evicted between steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = r.launch_count()
    step_ms = []
    barrier()
    for _ in range(args.steps):
        r.flush_l2(0)
        r.event_record(0)
        step()
        if args.workload == "frame8k" and args.gather in ("copy", "forward"):
            r.join_copies()  # the timed interval ends when this rank's tile pushes have landed, not when its kernels end
        r.event_record(1)
        if args.workload == "frame8k" and world > 1:
            shard.finish()  # frame boundary: all ranks' tiles have landed on GPU 0
        step_ms.append(r.event_elapsed_ms(0, 1))
        if args.workload == "seq1080p":
            for k, name in enumerate(pass_ms):
                pass_ms[name].append(r.last_pass_ms(k))
    barrier()
    launches = r.launch_count() - launches0
    clocks = sampler.stop()
    dev_ms_total = float(sum(step_ms))

    # ---- end-to-end region: host uniforms in, finished frame out to pinned host memory, every step.  The read-back of
    # frame k runs on the context's copy stream while frame k+1 renders into the other ping-pong image (mtReadImageAsync).
    seq = args.workload == "seq1080p"
    # after the swap: the frame just finished.  The sharded 8K frame does not swap: the peers store into the one image GPU 0
    # exported, so every frame is gathered in, and read back from, IMAGE_CLOUD_CUR (the next dispatch waits for the read).
    sharded = args.workload == "frame8k"
    out_which = api.IMAGE_LDR_PREV if seq else (api.IMAGE_CLOUD_CUR if sharded else api.IMAGE_CLOUD_PREV)
    nbytes = w * h * (4 if seq else (8 if args.storage == 2 else 16))
    pinned = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    h2d = int(cam.nbytes + tm.nbytes + tun.nbytes + sky.nbytes)
    e2e_read = (rank == 0) or args.workload != "frame8k"
    read_mask = args.workload in ("cloud4k", "views256") or (sharded and args.gather_mask)
    grey_bytes = w * h * 4  # one float per pixel: the decoded god-ray value (mtReadGodRayGreyAsync)
    pinned_mask = [torch.empty(grey_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)] if read_mask else None

    def e2e_step(i):
        if not seq:
            r.set_camera(cam); r.set_time(tm); r.set_tuning(tun); r.set_sun_and_sky(sky)
        step()
        if args.workload == "frame8k" and world > 1:
            shard.finish()
        if not seq and not sharded:
            r.swap_ping_pong()  # mtFrame swaps by itself
        if e2e_read:
            r.read_image_async(out_which, pinned[i & 1].data_ptr(), nbytes)
            if read_mask:  # the Cloud pass has two outputs (cloudRayMarch.comp:824-825): HDR colour and the grey-scale god-ray image
                r.read_godray_grey_async(pinned_mask[i & 1].data_ptr(), grey_bytes)

    for i in range(2):
        e2e_step(i)
    r.wait_reads()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    r.wait_reads()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- the same end-to-end loop with RGBA16F images (MT_STORAGE_F16: the reference's own image format, Renderer.cpp:1431-1440):
    # half the bytes over PCIe.  Reported beside the RGBA32F figure (north_star names RGBA32F as the output), cloud4k only.
    e2e16_s, nbytes16 = None, w * h * 8
    if args.workload == "cloud4k" and args.storage != 2:
        r16 = api.CloudRenderer(w, h, device=local_rank, storage=2, flags=args.ctx_flags)
        r16.upload_noise(noise)
        pin16 = [torch.empty(nbytes16, dtype=torch.uint8, pin_memory=True) for _ in range(2)] + [torch.empty(w * h * 4, dtype=torch.uint8, pin_memory=True) for _ in range(2)]

        def e2e16_step(i):
            r16.set_camera(cam); r16.set_time(tm); r16.set_tuning(tun); r16.set_sun_and_sky(sky)
            r16.dispatch_cloud_full()
            r16.swap_ping_pong()
            r16.read_image_async(api.IMAGE_CLOUD_PREV, pin16[i & 1].data_ptr(), nbytes16)
            r16.read_godray_grey_async(pin16[2 + (i & 1)].data_ptr(), w * h * 4)

        for i in range(2):
            e2e16_step(i)
        r16.wait_reads()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e16_step(i)
        r16.wait_reads()
        r16.synchronize()
        barrier()
        e2e16_s = time.perf_counter() - t0
        r16.close()

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([dev_ms_total, e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(per_rank, t)
        rank_ms = [round(float(x[0]) / args.steps, 4) for x in per_rank]   # device ms per step of every rank
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_total, e2e_s = float(t[0]), float(t[1])
        if e2e16_s is not None:
            t16 = torch.tensor([e2e16_s], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t16, op=dist.ReduceOp.MAX)
            e2e16_s = float(t16[0])
        cs = torch.tensor([counters[k] for k in ("rays", "rays_marched", "steps", "steps_incloud", "cone_hits", "early_exits")],
                          dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)

    ms_per_step = dev_ms_total / args.steps
    value = rays_total_per_step / (ms_per_step * 1e-3) / 1e6
    e2e_value = rays_total_per_step * args.steps / e2e_s / 1e6

    line = None
    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        # roofline of the dominant kernel (cloud_raymarch), per launch, rank 0's frame
        flop = algorithmic_flop(counters)
        kern_ms = statistics.mean(step_ms) if args.workload in ("cloud4k", "frame8k") and not (args.workload == "frame8k" and world > 1) else None
        roofline = None
        if kern_ms:
            ach_tflops = flop / (kern_ms * 1e-3) / 1e12
            hbm_ach = HBM_BYTES_PER_RAY * counters["rays"] / (kern_ms * 1e-3) / 1e9
            roofline = {
                "bound": "fp32-issue", "kernel": "cloud_raymarch_kernel", "achieved": round(ach_tflops, 3),
                "peak": round(fp32_peak_gflops / 1e3, 3), "unit": "TFLOP/s", "frac": round(ach_tflops / (fp32_peak_gflops / 1e3), 4),
                "peak_source": "measured FP32 FMA micro-benchmark (mtMeasureFp32Peak) on this GPU; no tensor work in this path",
                "peak_check": "filled below",
                "algorithmic_gflop_per_launch": round(flop / 1e9, 3), "filtered_fetches_per_launch": filtered_fetches(counters),
                "gfetch_per_s": round(filtered_fetches(counters) / (kern_ms * 1e-3) / 1e9, 3), "traffic": ncu_traffic_bytes(),
                "hbm": {"bound": "hbm", "achieved": round(hbm_ach, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(hbm_ach / hbm_peak, 5),
                        "algorithmic_bytes_per_launch": int(HBM_BYTES_PER_RAY * counters["rays"]), "peak_source": hbm_src},
            }
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "strong" if (args.workload == "frame8k") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "rays_per_step": int(rays_total_per_step), "l2": "flushed between timed steps (256 MiB memset)",
                       "noise": "reference noise volumes (tests/golden/noise_volumes.npz)", "parallelism": f"row-tiles x{world}" if args.workload == "frame8k" else f"views x{world}"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int((nbytes + (grey_bytes if read_mask else 0)) if e2e_read else 0),
                    "reads": ("HDR colour (RGBA) + grey-scale god-ray image (one float per pixel, decoded on the device)" if read_mask else ("LDR frame" if seq else "HDR colour (the god-ray image stays on its GPU unless --gather-mask)")),
                    "ms_per_step": round(1e3 * e2e_s / args.steps, 4)},
            "e2e_f16": None if e2e16_s is None else {
                "value": round(rays_total_per_step * args.steps / e2e16_s / 1e6, 2), "unit": UNIT, "ms_per_step": round(1e3 * e2e16_s / args.steps, 4),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(nbytes16 + w * h * 4), "storage": "MT_STORAGE_F16: RGBA16F images, the reference's own format (Renderer.cpp:1431-1440)",
                "reads": "HDR colour (RGBA16F) + grey-scale god-ray image (one float per pixel)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "work": {k: int(v) for k, v in counters.items()},
        }
        if world > 1:
            line["rank_ms"] = rank_ms
            line["e2e"]["host_placement"] = numa or "default (no NUMA binding)"
        if args.workload == "seq1080p":  # per-pass device time and HBM roofline of the bandwidth passes (SURVEY 8d bytes/pixel)
            px = w * h
            algo = {"reproject": 32 * px, "godrays": 48 * px, "tonemap": 20 * px}
            passes = {}
            for name, v in pass_ms.items():
                ms = statistics.median(v)
                e = {"ms": round(ms, 4)}
                if name == "tonemap" and algo["tonemap"] / (ms * 1e-3) / 1e9 > hbm_peak:
                    # mtFrame fuses the tone map into the god-ray kernel's store: what is timed here is an empty event pair, and
                    # the pass's 20 B/pixel are part of the god-ray kernel's traffic -- no roofline fraction of its own
                    e["note"] = "fused into the god-ray kernel (mtFrame); no kernel of its own, no roofline fraction"
                elif name in algo:
                    gbs = algo[name] / (ms * 1e-3) / 1e9
                    e.update({"bound": "hbm", "algorithmic_bytes": algo[name], "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                              "frac": round(gbs / hbm_peak, 4)})
                passes[name] = e
            line["passes"] = passes
            line["config"]["frame"] = "16-frame pan repeated; value = rays marched per frame (W*H/16) / frame time"

    if shard is not None and world > 1:
        shard.close()
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    r.close()

    if args.workload == "cloud4k" and not args.no_views256:
        rec = views256_record(args, api, sharding, torch, dist if world > 1 else None, world, rank, local_rank, noise)
        if rank == 0:
            line["views256"] = rec

    if args.workload == "cloud4k" and not args.no_seq1080p:
        rec = seq1080p_record(args, api, torch, dist if world > 1 else None, world, rank, local_rank, noise)
        if rank == 0:
            line["seq1080p"] = rec

    if rank == 0 and world == 1 and args.workload == "cloud4k" and not args.no_hw_filter:
        hcam, htm, hsky, _ = scene_for_view(0, W4K, H4K)
        line["hw_cone_filter"] = hw_cone_filter_record(args, api, local_rank, noise, hcam, htm, hsky)

    if world > 1 and args.workload == "cloud4k" and not args.no_sharded_8k:
        rec = sharded_8k_record(args, api, sharding, torch, dist, world, rank, local_rank, noise, steps=max(args.steps, 10))
        rec16 = sharded_8k_record(args, api, sharding, torch, dist, world, rank, local_rank, noise, steps=max(args.steps, 10), storage=2) if args.storage != 2 else None
        if rank == 0:
            line["sharded_8k"] = rec
            line["sharded_8k_f16"] = rec16

    if rank == 0 and line.get("roofline"):
        # cross-check of the probe: SMs x 128 FP32 lanes x 2 flop x the SM clock observed under load
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 0.0
        nominal = sm_count * 128 * 2 * float(mhz) * 1e6 / 1e12
        line["roofline"]["peak_check"] = {"sm_count": sm_count, "sm_mhz": mhz, "nominal_tflops": round(nominal, 2),
                                          "probe_over_nominal": round((fp32_peak_gflops / 1e3) / nominal, 4) if nominal else None,
                                          "frac_vs_nominal": round(line["roofline"]["achieved"] / nominal, 4) if nominal else None}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "cloud4k":
        v, cores, desc, _, kind = cpu_reference_sample(noise, W4K, H4K, target_s=12.0)
        line["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
