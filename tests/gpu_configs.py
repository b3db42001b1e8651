"""Probe (not a test): BASELINE configs 3 and 4 at full size on one GPU, device-resident; prints one JSON line each."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import diffeq_b200 as D
from diffeq_b200 import synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # log2 reduction of the trajectory count
HBM_PEAK = 6554.6  # MEASURED_PEAKS.json hbm_gbs

if which in ("1", "all"):
    ts = D.linspace(0.0, 100.0, 100001)
    p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.Lorenz)).params(synth.LORENZ_PARAMS).init([0.1, 0.0, 0.0]).tspan(ts).build()
    for rep in range(2):
        s = p.solve(D.Ode.Ode4, D.OdeOptionMap().insert(D.Output.Tspan), fetch=False)
    print(json.dumps({"config": 1, "what": "single Lorenz, RK4, linspace(0,100,100001), all 100001 states stored", "kernel_ms": s.kernel_ms,
                      "steps_per_s": 100000 / (s.kernel_ms * 1e-3), "note": "one thread: latency-bound parity anchor, not a throughput config"}), flush=True)
    y0 = synth.lorenz_ics(0, 1 << 20)
    p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.Lorenz)).params(synth.LORENZ_PARAMS).init(y0).tspan(D.linspace(0.0, 1.0, 1001)).build()
    for rep in range(2):
        s = p.solve(D.Ode.Ode4, fetch=False)
    print(json.dumps({"config": "1-ensemble", "what": "2^20 Lorenz, RK4, 1000 fixed steps, final state only", "kernel_ms": s.kernel_ms,
                      "steps_per_s": (1 << 20) * 1000 / (s.kernel_ms * 1e-3), "TFLOPs_at_96_per_step": (1 << 20) * 1000 * 96 / (s.kernel_ms * 1e-3) / 1e12}), flush=True)
    del p

if which in ("3", "all"):
    N = (1 << 22) >> scale
    y0, mu = synth.vdp_sweep(0, N, N)
    ts = D.linspace(0.0, 20.0, 1000)
    p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(mu).init(y0).tspan(ts).build()
    for mode, refill in ((D.Mode.Parity, 2), (D.Mode.Parity, 1), (D.Mode.Fast, 2)):
        o = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), D.Output.Tspan, mode, D.Refill(refill))
        for rep in range(2):
            s = p.solve(D.Ode.Ode45fe, o, fetch=False)
        bytes_out = N * 1000 * 2 * 8
        steps = s.totals["accepted"] + s.totals["rejected"]
        print(json.dumps({"config": 3, "mode": mode.name, "refill": refill, "ntraj": N, "kernel_ms": s.kernel_ms, "steps": steps,
                          "steps_per_s": steps / (s.kernel_ms * 1e-3), "out_GB": bytes_out / 1e9,
                          "achieved_GBps": bytes_out / 1e9 / (s.kernel_ms * 1e-3), "frac_of_measured_hbm": bytes_out / 1e9 / (s.kernel_ms * 1e-3) / HBM_PEAK}), flush=True)
    del p

if which in ("4", "all"):
    N = (1 << 24) >> scale
    y0 = synth.pendulum_ics(0, N)
    p = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.Pendulum)).params(synth.PENDULUM_PARAMS).init(y0).tspan([0.0, 10.0]).build()
    for mode in (D.Mode.Parity, D.Mode.Fast):
        o = D.OdeOptionMap().insert(D.Reltol(1e-12), D.Abstol(1e-12), mode)
        for rep in range(2):
            s = p.solve(D.Ode.Ode78, o, tableau_set=D.TableauSet.Textbook, fetch=False)
        steps = s.totals["accepted"] + s.totals["rejected"]
        print(json.dumps({"config": 4, "mode": mode.name, "ntraj": N, "kernel_ms": s.kernel_ms, "steps": steps,
                          "steps_per_s": steps / (s.kernel_ms * 1e-3), "TFLOPs_at_495_per_step": steps * 495 / (s.kernel_ms * 1e-3) / 1e12}), flush=True)
