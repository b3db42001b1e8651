import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import diffeq_b200 as D
from diffeq_b200 import synth
N = 1 << 20
y0, mu = synth.vdp_sweep(0, N, N)
ts = D.linspace(0.0, 20.0, 1000)
perm = np.random.default_rng(0).permutation(N)
p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(mu[perm].copy()).init(y0).tspan(ts).build()
o = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), D.Output.Tspan, D.DenseLayout(0), D.Refill(2))
for rep in range(3):
    t0 = time.time()
    s = p.solve(D.Ode.Ode45fe, o, fetch=False)
    print("rep", rep, "kernel_ms", round(s.kernel_ms, 2), "wall_ms", round((time.time() - t0) * 1e3, 1), flush=True)
    del s
