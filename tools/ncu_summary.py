#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics, stall mix and
the executed-instruction mix per member-step.  Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [warps] [steps] > profiles/xxx.txt
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
warps = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 7306
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"), " grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active"]
    for k in keys:
        if k in d:
            print(f"  {k:70s} {d[k]:>18s} {u[k]}")
    print("  stall reasons (warp-cycles per issued instruction):")
    for k in sorted(d):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v >= 0.01:
                print(f"    {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
byop, samp = collections.Counter(), collections.Counter()
tot = tots = n = 0
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        h = r
        ia, isamp, iexe = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        continue
    if h is None or len(r) <= iexe:
        continue
    try:
        e, s = int(r[iexe]), int(r[isamp])
    except ValueError:
        continue
    toks = r[ia].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    byop[op] += e
    samp[op] += s
    tot += e
    tots += s
    n += 1
ws = warps * steps
print(f"static SASS instructions {n}; executed warp-instructions {tot}; per warp-step {tot / ws:.1f}")
print(f"  {'opcode':10s} {'exec/step':>10s} {'stall-sample %':>15s}")
for op, c in byop.most_common(24):
    print(f"  {op:10s} {c / ws:10.1f} {100 * samp[op] / max(tots, 1):15.1f}")
