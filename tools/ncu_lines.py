#!/usr/bin/env python
"""ncu_lines.py — attributes the executed warp instructions of one kernel to CUDA source lines.

usage: python tools/ncu_lines.py REPORT.ncu-rep OBJECT.o MANGLED_KERNEL_NAME STEP_ATTEMPTS [top]

Joins the per-SASS-instruction counts of `ncu -i REPORT --page source --csv` (same instruction order as the binary) with the
line table of `nvdisasm --print-line-info` run on the cubin inside OBJECT.o (compiled with -lineinfo).  Output: instructions
per warp-step by source line (FP64-pipe / other), inlined callees attributed to the innermost line.
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, obj, kname, attempts = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
lanes = float(dict(zip(raw[0], raw[2]))["smsp__thread_inst_executed_per_inst_executed.ratio"])
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(src[h])}
counts = [(r[col["Source"]].strip(), float(r[col["Instructions Executed"]] or 0)) for r in src[h + 1:] if len(r) > col["Instructions Executed"]]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
sec = dis[dis.index(".text." + kname + ":"):]
end = sec.find("//--------------------- .", 10)
sec = sec[:end] if end > 0 else sec
lines, cur = [], ("?", 0)
for l in sec.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
assert len(lines) == len(counts), (len(lines), len(counts))
ws = attempts / lanes
agg = {}
for (loc, sass), (s2, n) in zip(lines, counts):
    op = re.sub(r"^@!?U?P\d+\s+", "", sass).split()[0]
    a = agg.setdefault(loc, [0.0, 0.0])
    a[0 if (op.startswith("D") and not op.startswith("DEPBAR")) else 1] += n / ws
byfile = {}
for (f, _), (a, b) in agg.items():
    x = byfile.setdefault(f, [0.0, 0.0]); x[0] += a; x[1] += b
print("per warp-step by file (FP64, other):", {k: (round(v[0], 1), round(v[1], 1)) for k, v in byfile.items()})
for (f, ln), (a, b) in sorted(agg.items(), key=lambda kv: -(2 * kv[1][0] + kv[1][1]))[:top]:
    print(f"{a:7.1f} {b:7.1f}  {f}:{ln}")
