#!/usr/bin/env python
"""Executed instructions per member-step BY SOURCE LINE of the step kernel (CPU box):

    python tools/ncu_lines.py gpurun_out/x.ncu-rep sipnet_b200/csrc/build/sip_run_fast_cropn_128.o MEMBERS STEPS [TOP]

Joins the per-SASS-instruction execution counts of an `ncu --set full --import-source on` capture with the line table
of the object the kernel was built from (`nvdisasm -g`): the capture lists the kernel's instructions in address order,
so does the disassembly.  The object must be the one inside the library the capture ran (same build)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, members, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernel = dict(zip(rows[0], rows[2]))["Kernel Name"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = next(r for r in srows if "Source" in r and "Instructions Executed" in r)
ia, ie = h.index("Source"), h.index("Instructions Executed")
isamp = h.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in h else None
executed = []
for r in srows:
    if len(r) <= ie:
        continue
    try:
        e = int(r[ie])
    except ValueError:
        continue
    toks = r[ia].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0])
    executed.append((op, e, int(r[isamp]) if isamp is not None and r[isamp].isdigit() else 0))

# the kernel's section in the object: match by the demangled name's distinguishing template arguments
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout.splitlines()
    names = subprocess.run(["cuobjdump", "-elf", os.path.join(td, cubin)], capture_output=True, text=True).stdout
sections = [(i, l[6:-1]) for i, l in enumerate(dis) if l.startswith(".text.")]
want = re.sub(r"\s+", "", kernel)


def demangle(n):
    return re.sub(r"\s+", "", subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip())


match = [(i, n) for i, n in sections if demangle(n).replace("(bool)", "").replace("(int)", "") == want.replace("(bool)", "").replace("(int)", "")]
if not match:  # fall back: same instruction count
    cand = []
    for j, (i, n) in enumerate(sections):
        end = sections[j + 1][0] if j + 1 < len(sections) else len(dis)
        cnt = sum(1 for l in dis[i:end] if re.match(r"\s+/\*[0-9a-f]+\*/\s", l))
        if cnt == len(executed):
            cand.append((i, n))
    match = cand
assert len(match) == 1, f"kernel section not found uniquely ({len(match)} candidates) for {kernel}"
start = match[0][0]
end = next((i for i, _ in sections if i > start), len(dis))
cur = None
ins = []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", l)
    if m:
        ins.append((m.group(1), cur))
assert len(ins) == len(executed), (len(ins), len(executed))
bad = sum(1 for (o1, _), (o2, _, _) in zip(ins, executed) if o1.split(".")[0] != o2.split(".")[0])
assert bad == 0, f"{bad} opcodes differ between the capture and the object: not the same build"
ws = (members // 32) * steps
by = collections.Counter()
byop = collections.defaultdict(collections.Counter)
samp = collections.Counter()
static = collections.Counter()
for (op, line), (_, e, s) in zip(ins, executed):
    by[line] += e
    byop[line][op.split(".")[0]] += e
    samp[line] += s
    static[line] += 1
tot = sum(by.values())
tots = max(1, sum(samp.values()))
print(f"kernel: {kernel}\nexecuted warp-instructions per warp-step: {tot / ws:.1f}; static {len(ins)}")
files = collections.Counter()
for (f, _), c in by.items():
    files[f] += c
print("by file:", {f: round(c / ws, 1) for f, c in files.most_common()})
print(f"{'exec/step':>9} {'stall %':>7} {'static':>6}  line")
for line, c in by.most_common(top):
    mix = ", ".join(f"{o} {v / ws:.1f}" for o, v in byop[line].most_common(5))
    print(f"{c / ws:9.1f} {100.0 * samp[line] / tots:7.1f} {static[line]:6d}  {line[0]}:{line[1]}  [{mix}]")
