"""Quick GPU timing probe (not a test): config 2 at 1M trajectories, device-resident, plus the FP64 peak."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import diffeq_b200 as D
from diffeq_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
print("fp64 peak TFLOP/s, MHz:", D.measure_fp64_peak())
y0 = synth.lorenz_ics(0, N)
p = D.OdeProblem.builder().fun(D.Rhs.builtin(0)).params(synth.LORENZ_PARAMS).init(y0).tspan([0.0, T]).build()
for refill, mode in ((2, 0), (1, 0), (0, 0), (2, 0), (2, 1), (1, 1), (2, 1)):
    o = D.OdeOptionMap().insert(D.Reltol(1e-8), D.Abstol(1e-8), D.Refill(refill), D.Mode(mode))
    t = time.time(); s = p.solve(D.Ode.Ode45, o, fetch=False); w = time.time() - t
    steps = s.totals["accepted"] + s.totals["rejected"]
    print(f"mode={mode} refill={refill} kernel_ms={s.kernel_ms:.3f} wall_ms={w*1e3:.1f} steps={steps} steps/s={steps/(s.kernel_ms*1e-3):.4e} TFLOP/s(314/step)={steps*314/(s.kernel_ms*1e-3)/1e12:.2f}")

o = D.OdeOptionMap().insert(D.Reltol(1e-8), D.Abstol(1e-8), D.Libm.LlvmFolded)
for rep in range(2):
    s = p.solve(D.Ode.Ode45, o, fetch=False)
steps = s.totals["accepted"] + s.totals["rejected"]
print(f"parity, libm LLVM-folded (sqrt) kernel_ms={s.kernel_ms:.3f} steps={steps} steps/s={steps/(s.kernel_ms*1e-3):.4e}")
o = D.OdeOptionMap().insert(D.Reltol(1e-5), D.Abstol(1e-6), D.Precision.F32)
for rep in range(2):
    s = p.solve(D.Ode.Ode45, o, fetch=False)
steps = s.totals["accepted"] + s.totals["rejected"]
print(f"f32 (rtol 1e-5) kernel_ms={s.kernel_ms:.3f} steps={steps} steps/s={steps/(s.kernel_ms*1e-3):.4e}")
o = D.OdeOptionMap().insert(D.Reltol(1e-5), D.Abstol(1e-6), D.Mode.Fast)
for rep in range(2):
    s = p.solve(D.Ode.Ode45, o, fetch=False)
steps = s.totals["accepted"] + s.totals["rejected"]
print(f"f64 fast (rtol 1e-5) kernel_ms={s.kernel_ms:.3f} steps={steps} steps/s={steps/(s.kernel_ms*1e-3):.4e}")

# small ensembles (strong-scaling regime): 2^17 trajectories = 1.7 waves of the resident lanes
y1 = synth.lorenz_ics(0, 1 << 17)
p1 = D.OdeProblem.builder().fun(D.Rhs.builtin(0)).params(synth.LORENZ_PARAMS).init(y1).tspan([0.0, T]).build()
for refill, mode in ((2, 0), (1, 0), (2, 1), (1, 1)):
    o = D.OdeOptionMap().insert(D.Reltol(1e-8), D.Abstol(1e-8), D.Refill(refill), D.Mode(mode))
    for rep in range(2):
        s = p1.solve(D.Ode.Ode45, o, fetch=False)
    steps = s.totals["accepted"] + s.totals["rejected"]
    print(f"2^17 traj mode={mode} refill={refill} kernel_ms={s.kernel_ms:.3f} steps/s={steps/(s.kernel_ms*1e-3):.4e}")
