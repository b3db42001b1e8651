"""Probe (not a test): ode23s throughput on a Van der Pol stiffness sweep.  usage: python tests/gpu_stiff_bench.py [ntraj] [T]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffeq_b200 as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
mu = np.geomspace(0.1, 1000.0, n)[:, None]
y0 = np.tile([2.0, 0.0], (n, 1))
prob = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(mu).init(y0).tspan([0.0, T]).build()
for name, tset, extra in (("reference", 0, ()), ("textbook", 1, ()), ("reference_llvm_folded", 0, (D.Libm.LlvmFolded,)),
                          ("textbook_fast_mode", 1, (D.Mode.Fast,)), ("reference_fast_mode", 0, (D.Mode.Fast,))):
    o = D.OdeOptionMap().insert(D.Reltol(1e-5), D.Abstol(1e-8), D.MaxSteps(20000), *extra)   # the reference variant loops forever on ~3 per million trajectories
    best = None
    for rep in range(3):
        s = prob.solve(D.Ode.Ode23s, o, tableau_set=tset, fetch=False)
        best = s.kernel_ms if best is None else min(best, s.kernel_ms)
    steps = s.totals["accepted"] + s.totals["rejected"]
    st = prob.solve(D.Ode.Ode23s, o, tableau_set=tset).status
    print(f"ode23s {name}: status counts {np.bincount(st).tolist()}")
    print(f"ode23s {name}: ntraj={n} T={T} kernel_ms={best:.2f} attempts={steps} steps/s={steps / best * 1e3:.3e} nfe={s.totals['nfe']}")
# CPU oracle on a sample for scale
try:
    from oracle import oracle as O
    m = 65536
    idx = np.linspace(0, n - 1, m).astype(int)
    t0 = time.perf_counter()
    r = O.ensemble_ode23s(O.VANDERPOL, mu[idx], y0[idx], [0.0, T], O.opts(1e-5, 1e-8, max_steps=20000))
    dt = time.perf_counter() - t0
    st = int(r["accepted"].sum() + r["rejected"].sum())
    print(f"cpu oracle sample: {m} trajectories, {st} attempts in {dt:.2f}s on {O.num_threads()} threads = {st / dt:.3e} steps/s")
except Exception as e:  # noqa: BLE001
    print("cpu oracle sample skipped:", e)
