"""Experiment driver (not a test): time config 2 (parity + fast) with the product library and each
libdiffeq_b200_<variant>.so present.  Usage: python tests/gpu_variants.py"""
import glob, os, subprocess, sys
code = r'''
import sys
sys.path.insert(0, ".")
import diffeq_b200 as D
from diffeq_b200 import synth
y0 = synth.lorenz_ics(0, 1 << 20)
p = D.OdeProblem.builder().fun(D.Rhs.builtin(0)).params(synth.LORENZ_PARAMS).init(y0).tspan([0.0, 10.0]).build()
out = []
for mode in (0, 1):
    o = D.OdeOptionMap().insert(D.Reltol(1e-8), D.Abstol(1e-8), D.Mode(mode))
    best = 1e9
    for rep in range(4):
        s = p.solve(D.Ode.Ode45, o, fetch=False)
        best = min(best, s.kernel_ms)
    out.append("%s %.3f ms (steps %d)" % ("parity" if mode == 0 else "fast", best, s.totals["accepted"] + s.totals["rejected"]))
print(" | ".join(out))
'''
libs = [os.path.abspath("diffeq_b200/libdiffeq_b200.so")] + sorted(os.path.abspath(l) for l in glob.glob("diffeq_b200/libdiffeq_b200_*.so"))
for l in libs:
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DEQ_LIB=l), capture_output=True, text=True)
    print(os.path.basename(l), (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
