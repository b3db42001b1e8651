"""CPU: numerical sanity of the restated methods (SURVEY.md §4 iii) — independent of bit parity.  Fixed-step methods show
their textbook convergence order, the TEXTBOOK adaptive tableaus agree with scipy's integrators on a smooth and on a
driven problem, and the REFERENCE tableaus show the defects the survey documents (step-size collapse)."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

LOR = [10.0, 28.0, 8.0 / 3.0]
PEND = [0.5, 1.2, 2.0 / 3.0]


def _vdp_exact(T, mu):
    s = solve_ivp(lambda t, y: [y[1], mu * (1 - y[0] ** 2) * y[1] - y[0]], [0, T], [2.0, 0.0], method="DOP853", rtol=1e-13, atol=1e-13)
    return s.y[:, -1]


@pytest.mark.parametrize("method,order", [("FEULER", 1), ("HEUN", 2), ("MIDPOINT", 2), ("ODE4", 4)])
def test_fixed_step_convergence_order(oracle, method, order):
    exact = _vdp_exact(1.0, 1.0)
    errs = []
    for n in (200, 400, 800):
        y = oracle.solve_fixed(getattr(oracle, method), oracle.VANDERPOL, [1.0], [2.0, 0.0], oracle.linspace(0.0, 1.0, n + 1), want_all=False)["yfinal"]
        errs.append(np.linalg.norm(y - exact))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(abs(r - order) < 0.25 for r in rates), (errs, rates)


@pytest.mark.parametrize("method", ["ODE45", "ODE45FE", "ODE23", "ODE78"])
def test_textbook_adaptive_tableaus_match_scipy(oracle, method):
    exact = _vdp_exact(5.0, 2.0)
    tol = 1e-10 if method in ("ODE45", "ODE45FE", "ODE78") else 1e-7
    r = oracle.solve_adaptive(getattr(oracle, method), oracle.VANDERPOL, [2.0], [2.0, 0.0], [0.0, 5.0], oracle.opts(tol, tol, max_steps=2_000_000),
                              tset=oracle.TEXTBOOK, want_all=False)
    assert r["status"] == 0 and np.linalg.norm(r["yfinal"] - exact) < 3e3 * tol
    # driven pendulum (time-dependent RHS: exercises the c nodes, e.g. the dopri5 c[3] fix)
    s = solve_ivp(lambda t, y: [y[1], -PEND[0] * y[1] - np.sin(y[0]) + PEND[1] * np.cos(PEND[2] * t)], [0, 10], [0.3, 0.1],
                  method="DOP853", rtol=1e-13, atol=1e-13)
    r = oracle.solve_adaptive(getattr(oracle, method), oracle.PENDULUM, PEND, [0.3, 0.1], [0.0, 10.0], oracle.opts(tol, tol, max_steps=2_000_000),
                              tset=oracle.TEXTBOOK, want_all=False)
    assert np.linalg.norm(r["yfinal"] - s.y[:, -1]) < 3e4 * tol


def test_reference_tableau_defects_show_up(oracle):
    """the reference dopri5 (c[3] = 0.75) needs ~40x the steps of the textbook one on a time-dependent RHS;
    reference rk23 / feh78 collapse (they exhaust any step budget on a short horizon)."""
    o = oracle.opts(1e-8, 1e-8, max_steps=100000)
    ref = oracle.solve_adaptive(oracle.ODE45, oracle.PENDULUM, PEND, [0.3, 0.1], [0.0, 10.0], o, tset=oracle.REFERENCE, want_all=False)
    txt = oracle.solve_adaptive(oracle.ODE45, oracle.PENDULUM, PEND, [0.3, 0.1], [0.0, 10.0], o, tset=oracle.TEXTBOOK, want_all=False)
    assert ref["accepted"] > 20 * txt["accepted"]
    for m in (oracle.ODE23, oracle.ODE78):
        r = oracle.solve_adaptive(m, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], [0.0, 5.0], oracle.opts(1e-8, 1e-8, max_steps=20000), want_all=False)
        assert r["status"] == oracle.ST_MAX_STEPS and r["t_reached"] < 1.0


def test_dense_output_is_cubic_hermite_accurate(oracle):
    """Points::All rows at tspan times interpolate the solution to the order of the cubic Hermite formula (problem.rs:783-798)."""
    ts = oracle.linspace(0.0, 5.0, 51)
    r = oracle.solve_adaptive(oracle.ODE45, oracle.VANDERPOL, [1.0], [2.0, 0.0], ts, oracle.opts(1e-9, 1e-9), tset=oracle.TEXTBOOK)
    s = solve_ivp(lambda t, y: [y[1], (1 - y[0] ** 2) * y[1] - y[0]], [0, 5], [2.0, 0.0], method="DOP853", rtol=1e-13, atol=1e-13, t_eval=ts[1:-1])
    lut = dict(zip(r["tout"].tolist(), r["yout"]))
    err = max(np.linalg.norm(lut[t] - s.y[:, i]) for i, t in enumerate(ts[1:-1].tolist()) if t in lut)
    assert sum(t in lut for t in ts[1:-1].tolist()) >= 45 and err < 1e-5
