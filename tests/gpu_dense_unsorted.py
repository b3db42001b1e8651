"""Probe (not a test): config 3 with the mu sweep in sorted vs randomly permuted trajectory order (dense output)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import diffeq_b200 as D
from diffeq_b200 import synth
N = 1 << 20
y0, mu = synth.vdp_sweep(0, N, N)
ts = D.linspace(0.0, 20.0, 1000)
perm = np.random.default_rng(0).permutation(N)
for name, m in (("sorted", mu), ("permuted", mu[perm].copy())):
    p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(m).init(y0).tspan(ts).build()
    for out, lay in ((D.Output.Final, 0), (D.Output.Tspan, 0), (D.Output.Tspan, 1), (D.Output.All, 0)):
        o = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), out, D.DenseLayout(lay))
        for rep in range(2):
            s = p.solve(D.Ode.Ode45fe, o, fetch=False)
        print(json.dumps({"order": name, "output": out.name, "layout": D.DenseLayout(lay).name, "kernel_ms": round(s.kernel_ms, 3), "steps": s.totals["accepted"] + s.totals["rejected"]}), flush=True)
    o = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), D.Refill(0))
    for rep in range(2):
        s = p.solve(D.Ode.Ode45fe, o, fetch=False)
    print(json.dumps({"order": name, "output": "Final", "refill": 0, "kernel_ms": round(s.kernel_ms, 3)}), flush=True)
    del p
