"""Probe (not a test): the four dense-output flavours of config 3 at 2^20 trajectories, one launch each after a warm-up
(per order: SoA row-synchronous kernel, staged trajectory-major, SoA lane-refill kernel; every second launch is the timed one)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import diffeq_b200 as D
from diffeq_b200 import synth
N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
y0, mu = synth.vdp_sweep(0, N, N)
ts = D.linspace(0.0, 20.0, 1000)
perm = np.random.default_rng(0).permutation(N)
for name, m in (("sorted", mu), ("permuted", mu[perm].copy())):
    p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(m).init(y0).tspan(ts).build()
    for lay, refill in ((0, 2), (0, 0), (0, 1), (1, 2)):   # SoA auto, SoA row-synchronous kernel, SoA lane-refill kernel, staged trajectory-major
        o = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), D.Output.Tspan, D.DenseLayout(lay), D.Refill(refill))
        for rep in range(2):
            s = p.solve(D.Ode.Ode45fe, o, fetch=False)
        print(json.dumps({"order": name, "layout": D.DenseLayout(lay).name, "refill": refill, "kernel_ms": round(s.kernel_ms, 3), "steps": s.totals["accepted"] + s.totals["rejected"],
                          "GBps": N * 1000 * 2 * 8 / (s.kernel_ms * 1e-3) / 1e9}), flush=True)
    del p
