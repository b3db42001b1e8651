"""Probe (not a test): BASELINE config 4 (pendulum, Fehlberg 7(8) TEXTBOOK, rtol = atol = 1e-12) at a reduced trajectory count.
usage: python tests/gpu_cfg4_probe.py [log2_ntraj] [parity|fast]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffeq_b200 as D
from diffeq_b200 import synth

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = D.Mode.Fast if (len(sys.argv) > 2 and sys.argv[2] == "fast") else D.Mode.Parity
n = 1 << lg
prob = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.Pendulum)).params(synth.PENDULUM_PARAMS).init(synth.pendulum_ics(0, n)).tspan([0.0, 10.0]).build()
o = D.OdeOptionMap().insert(D.Reltol(1e-12), D.Abstol(1e-12), mode)
for rep in range(3):
    s = prob.solve(D.Ode.Ode78, o, tableau_set=D.TableauSet.Textbook, fetch=False)
    st = s.totals["accepted"] + s.totals["rejected"]
    print(f"cfg4 {mode.name} ntraj=2^{lg} kernel_ms={s.kernel_ms:.2f} steps={st} steps/s={st / s.kernel_ms * 1e3:.3e}")
