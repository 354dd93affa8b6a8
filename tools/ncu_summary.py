#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters the design is argued
from.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--blocks [kernel-regex]] [> profiles/rN_name.md]
--blocks adds, from the source page (`--page source --csv`, reports captured with --import-source on), the opcode mix weighted by
executions and the basic blocks by execution count (share of warp instructions, share of stall samples, active lanes)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (top pipe)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU (MUFU/conversions) %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "pipe TEX %"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "pipe tensor %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no_instruction (I-cache)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short_scoreboard (MUFU/LDS)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long_scoreboard (L1/L2/DRAM)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not_selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch_resolving"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
]


def blocks(rep, kernel):
    import collections
    import re
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kernel] if kernel else [])
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    groups, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            groups.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for g in groups[:1]:
        hdr = g["rows"][0]
        data = [r for r in g["rows"][1:] if len(r) == len(hdr)]
        ix = {h: i for i, h in enumerate(hdr)}
        ex = lambda r: int(r[ix["Instructions Executed"]])
        sm = lambda r: int(r[ix["# Samples"]])
        tot, ts = sum(map(ex, data)), max(1, sum(map(sm, data)))
        print(f"## Where the instructions go: `{g['name']}` ({tot} warp instructions, {len(data)} SASS instructions)\n")
        hist = collections.Counter()
        for r in data:
            src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
            hist[src.split()[0].rstrip(";").split(".")[0]] += ex(r)
        print("Opcode mix (share of executed warp instructions): " + ", ".join(f"{op} {100 * c / tot:.1f}" for op, c in hist.most_common(24)) + "\n")
        print("| SASS instructions | executions per instruction | share of warp instructions | share of stall samples | active lanes |")
        print("|---|---|---|---|---|")
        prev, start = None, 0
        for i, r in enumerate(data + [None]):
            e = ex(r) if r else -1
            if e != prev:
                if prev is not None and (i - start) >= 4 and prev * (i - start) / tot > 0.004:
                    smp = sum(sm(x) for x in data[start:i])
                    print(f"| {i - start} | {prev} | {100 * prev * (i - start) / tot:.1f} % | {100 * smp / ts:.1f} % | {float(data[start][ix['Avg. Threads Executed']]):.1f} |")
                start, prev = i, e
        print()


def main():
    rep = sys.argv[1]
    if "--blocks" in sys.argv:
        k = sys.argv.index("--blocks")
        blocks(rep, sys.argv[k + 1] if k + 1 < len(sys.argv) else "")
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"### launch {k}: `{d.get('Kernel Name', '?')}`  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
        print("| counter | value |")
        print("|---|---|")
        for key, label in KEYS:
            if key in d and d[key] != "":
                print(f"| {label} (`{key}`) | {d[key]} {u.get(key, '')} |")
        print()


if __name__ == "__main__":
    main()
