#!/usr/bin/env python
"""Kernel facts for bench.py's roofline section, read from one `ncu --set full --import-source on` capture of the
step kernel (CPU box):  python tools/ncu_facts.py gpurun_out/x.ncu-rep MEMBERS STEPS profiles/r02_kernel_facts.json

fp64_pipe_inst_per_member_step = executed FP64-pipe warp-instructions (DFMA, DADD, DMUL, DSETP, DMNMX) per
warp-step = per member-step per thread; ncu_* = the capture's own utilisation figures."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

rep, members, steps, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
u = dict(zip(rows[0], rows[1]))


def num(k):
    return float(d[k].replace(",", ""))


def to_bytes(k):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
    return num(k) * scale


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = next(r for r in srows if "Source" in r and "Instructions Executed" in r)
ia, ie = h.index("Source"), h.index("Instructions Executed")
byop = collections.Counter()
for r in srows:
    if len(r) <= ie:
        continue
    try:
        e = int(r[ie])
    except ValueError:
        continue
    toks = r[ia].split()
    byop[(toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]] += e
ws = (members // 32) * steps
fp64 = sum(byop[o] for o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
facts = {
    "source": os.path.join("profiles", os.path.basename(out).replace("_kernel_facts.json", "") + "_ncu_step_kernel.txt") + " (" + os.path.basename(rep) + ")",
    "kernel": d.get("Kernel Name"), "members": members, "steps": steps,
    "kernel_ms_under_ncu": num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u["gpu__time_duration.sum"], 1.0),
    "inst_per_member_step": sum(byop.values()) / ws,
    "fp64_pipe_inst_per_member_step": fp64 / ws,
    "fp64_mix_per_member_step": {o: byop[o] / ws for o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX") if byop[o]},
    "ncu_fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"),
    "ncu_issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "ncu_warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": int(num("launch__registers_per_thread")),
    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
}
json.dump(facts, open(out, "w"), indent=1)
print(json.dumps(facts, indent=1))
