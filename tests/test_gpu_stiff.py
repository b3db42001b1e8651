"""GPU parity of the stiff solvers: ode23s (problem.rs:409-557) and oderosenbrock kr4 / s4
(problem.rs:560-650) through the C ABI against the CPU oracle and the committed goldens — every assert BIT-EXACT.
"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
STIFF = json.load(open(os.path.join(HERE, "golden", "stiff_cases.json")))
LOR = [10.0, 28.0, 8.0 / 3.0]


def fh(v):
    return [fh(x) for x in v] if isinstance(v, list) else float.fromhex(v)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<f8").tobytes()).hexdigest()


def _problem(D, system, params, y0, tspan):
    return D.OdeProblem.builder().fun(D.Rhs.builtin(system)).params(params).init(y0).tspan(tspan).build()


def _tspan(deq, desc):
    return [float(x) for x in desc[1:]] if desc[0] == "list" else deq.linspace(desc[1], desc[2], desc[3])


def _options(deq, opt, *extra):
    o = deq.OdeOptionMap().insert(*extra)
    for k, cls in (("reltol", deq.Reltol), ("abstol", deq.Abstol), ("minstep", deq.Minstep), ("maxstep", deq.Maxstep), ("initstep", deq.Initstep)):
        if k in opt:
            o.insert(cls(float.fromhex(opt[k])))
    if "max_steps" in opt:
        o.insert(deq.MaxSteps(opt["max_steps"]))
    if opt.get("libm_folded"):
        o.insert(deq.Libm.LlvmFolded)
    if opt.get("points_all") is False:
        o.insert(deq.Points.Specified)
    return o


def test_device_pow_special_operands(deq, oracle):
    """The step-size update pow(delta/err, 1/3) and the norms pow(s, 0.5) meet 0, inf and NaN on the stiff path."""
    x = np.array([0.0, np.inf, np.nan, 1.0, 5e-324, 1e-320, 1.7e308, 0.0, np.inf, np.nan, 2.0, 1e-300])
    y = np.array([1 / 3] * 7 + [0.5] * 5)
    got = deq.device_pow(x, y)
    L = oracle.lib()
    ref = np.array([L.oracle_pow(float(a), float(b)) for a, b in zip(x, y)])
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    rng = np.random.default_rng(5)
    x = np.exp(rng.uniform(-60, 60, 200000))
    got = deq.device_pow(x, np.full_like(x, 1 / 3))
    ref = np.array([L.oracle_pow(float(a), 1 / 3) for a in x[:40000]])
    assert np.array_equal(got[:40000].view(np.uint64), ref.view(np.uint64))


def test_rosenbrock_coefficients(deq):
    for which, m in (("kr4", deq.Ode.Ode4skr), ("s4", deq.Ode.Ode4ss)):
        g = STIFF["rosenbrock_coeffs"][which]
        co = deq.rosenbrock_coeffs(m)
        assert co["gamma"] == float.fromhex(g["gamma"])
        assert co["a"].tolist() == fh(g["a"]) and co["b"].tolist() == fh(g["b"]) and co["c"].tolist() == fh(g["c"])


def test_stiff_goldens_through_the_cuda_path(deq):
    """Every committed stiff golden (known answers of the independent Python restatement) reproduced by the CUDA path
    directly: the complete (tout, yout) streams by SHA-256, the counters, the exit status."""
    ran = 0
    for c in STIFF["cases"]:
        ts = _tspan(deq, c["tspan"])
        if c["system"].startswith("Cvdp"):   # dimension 4 / 6 / 12 / 24: user source through NVRTC (the built-in systems have n <= 3)
            prob = deq.OdeProblem.builder().fun(deq.Rhs.from_source(USER_CVDP, 2 * int(c["system"][4:]), 2)).params(fh(c["params"])) \
                .init(fh(c["y0"])).tspan(ts).build()
        else:
            prob = _problem(deq, getattr(deq.System, c["system"]), fh(c["params"]), fh(c["y0"]), ts)
        if c["kind"] == "ode23s":
            sol = prob.solve(deq.Ode.Ode23s, _options(deq, c["options"], deq.Output.All), tableau_set=0 if c["tableau_set"] == "reference" else 1)
            tout, yout = sol.trajectory(0)
            assert (int(sol.accepted[0]), int(sol.rejected[0]), int(sol.status[0]), len(tout)) == (c["accepted"], c["rejected"], c["status"], c["nout"]), c["name"]
            assert sol.final[0].tolist() == fh(c["final"]) and sol.t_reached[0] == float.fromhex(c["t_reached"]), c["name"]
            assert digest(tout) == c["sha256_tout"] and digest(yout) == c["sha256_yout"], c["name"]
        else:
            sol = prob.solve(deq.Ode.Ode4skr if c["which"] == "kr4" else deq.Ode.Ode4ss, deq.OdeOptionMap().insert(deq.Output.Tspan))
            nv = int(sol.nvalid[0])
            assert (int(sol.status[0]), nv) == (c["status"], c["nout"]), c["name"]
            assert digest(sol.yout[0][:nv]) == c["sha256_yout"] and digest(sol.tout) == c["sha256_tout"], c["name"]
            assert sol.final[0].tolist() == fh(c["final"]), c["name"]
        ran += 1
    assert ran == len(STIFF["cases"]) >= 30


@pytest.mark.parametrize("tset", [0, 1])
@pytest.mark.parametrize("folded", [0, 1])
def test_ode23s_vdp_sweep_final(deq, oracle, tset, folded):
    """Van der Pol stiffness sweep mu = 0.1 .. 1000 (the natural ode23s workload), final state + every counter."""
    n = 1003
    mu = np.geomspace(0.1, 1000.0, n)[:, None]
    y0 = np.tile([2.0, 0.0], (n, 1))
    ts = [0.0, 20.0]
    ref = oracle.ensemble_ode23s(oracle.VANDERPOL, mu, y0, ts, oracle.opts(1e-5, 1e-8, libm_folded=folded), tset=tset)
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-5), deq.Abstol(1e-8))
    if folded:
        o.insert(deq.Libm.LlvmFolded)
    sol = _problem(deq, deq.System.VanDerPol, mu, y0, ts).solve(deq.Ode.Ode23s, o, tableau_set=tset)
    for k in ("accepted", "rejected", "nfe", "status", "t_reached"):
        assert np.array_equal(getattr(sol, k), ref[k]), k
    assert np.array_equal(sol.final, ref["yfinal"])
    assert sol.totals == dict(accepted=int(ref["accepted"].sum()), rejected=int(ref["rejected"].sum()), nfe=int(ref["nfe"].sum()))
    assert (sol.status == 0).all()


@pytest.mark.parametrize("system", ["Lorenz", "Pendulum"])
def test_ode23s_three_and_two_dimensional_systems(deq, oracle, system):
    from diffeq_b200 import synth
    n = 300
    if system == "Lorenz":
        y0, params, ts = synth.lorenz_ics(0, n), np.array(LOR), [0.0, 1.5]
    else:
        y0, params, ts = synth.pendulum_ics(0, n), synth.PENDULUM_PARAMS, [0.0, 4.0]
    for tset in (0, 1):
        ref = oracle.ensemble_ode23s(getattr(oracle, system.upper()), params, y0, ts, oracle.opts(1e-6, 1e-9), tset=tset)
        sol = _problem(deq, getattr(deq.System, system), params, y0, ts).solve(
            deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-9)), tableau_set=tset)
        assert np.array_equal(sol.final, ref["yfinal"]) and np.array_equal(sol.accepted, ref["accepted"])
        assert np.array_equal(sol.rejected, ref["rejected"]) and np.array_equal(sol.nfe, ref["nfe"])


def test_ode23s_dense_rows_and_ragged_stream(deq, oracle):
    """Output.Tspan = the rows the reference pushes for the tspan times; Output.All = its complete (tout, yout), with
    Points::All and Points::Specified (problem.rs:519-542)."""
    n = 130
    mu = np.geomspace(0.5, 200.0, n)[:, None]
    y0 = np.tile([2.0, 0.0], (n, 1))
    ts = deq.linspace(0.0, 30.0, 77)
    oo = oracle.opts(1e-5, 1e-8)
    ref = oracle.ensemble_ode23s(oracle.VANDERPOL, mu, y0, ts, oo, dense=True)
    prob = _problem(deq, deq.System.VanDerPol, mu, y0, ts)
    base = (deq.Reltol(1e-5), deq.Abstol(1e-8))
    sol = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(*base, deq.Output.Tspan))
    assert np.array_equal(sol.nvalid, ref["nvalid"]) and (sol.nvalid == len(ts)).all()
    assert np.array_equal(sol.yout, ref["out"].transpose(2, 1, 0))
    assert np.array_equal(sol.final, ref["yfinal"])
    for pts, opt in ((0, deq.Points.All), (1, deq.Points.Specified)):
        sa = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(*base, deq.Output.All, opt))
        oo.points = pts
        for i in (0, 1, 57, n - 1):
            one = oracle.solve_ode23s(oracle.VANDERPOL, mu[i], y0[i], ts, oo)
            tout, yout = sa.trajectory(i)
            assert np.array_equal(tout, one["tout"]) and np.array_equal(yout, one["yout"]), (pts, i)
        if pts == 1:
            assert all(np.array_equal(sa.trajectory(i)[0], ts) for i in range(n))
    oo.points = 0


def test_ode23s_scheduling_and_partition_invariance(deq):
    n = 2000
    mu = np.geomspace(0.1, 300.0, n)[:, None]
    y0 = np.tile([2.0, 0.0], (n, 1))
    a = _problem(deq, deq.System.VanDerPol, mu, y0, [0.0, 10.0]).solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Refill(1)))
    b = _problem(deq, deq.System.VanDerPol, mu, y0, [0.0, 10.0]).solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Refill(0)))
    c = _problem(deq, deq.System.VanDerPol, mu[700:1500], y0[700:1500], [0.0, 10.0]).solve(deq.Ode.Ode23s)
    assert np.array_equal(a.final, b.final) and np.array_equal(a.accepted, b.accepted)
    assert np.array_equal(a.final[700:1500], c.final) and np.array_equal(a.rejected[700:1500], c.rejected)


def test_ode23s_fast_mode_within_tolerance(deq):
    """Mode.Fast for ode23s (FMA, reciprocal instead of n^2 divisions, sqrt norms, table pow): same formulas, held to the
    north-star floating-point bar against PARITY — matching step counts and final states within 10 x (rtol |y| + atol) —
    on the stiffness sweep away from the relaxation jumps, where any rounding difference is amplified."""
    n = 4096
    mu = np.geomspace(0.1, 1000.0, n)[:, None]
    y0 = np.tile([2.0, 0.0], (n, 1))
    rtol, atol = 1e-6, 1e-9
    base = (deq.Reltol(rtol), deq.Abstol(atol))
    prob = _problem(deq, deq.System.VanDerPol, mu, y0, [0.0, 20.0])
    for tset in (0, 1):
        a = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(*base, deq.MaxSteps(200000)), tableau_set=tset)
        f = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(*base, deq.MaxSteps(200000), deq.Mode.Fast), tableau_set=tset)
        ok = (a.status == 0) & (f.status == 0)
        assert ok.mean() > 0.999
        dev = np.max(np.abs(f.final - a.final) / (np.abs(a.final) * rtol + atol), axis=1)[ok]
        same = (a.accepted == f.accepted)[ok]
        print("ode23s fast vs parity set", tset, "same accepted", same.mean(), "dev median/p99/max", np.median(dev), np.quantile(dev, 0.99), dev.max(),
              "attempts", int(a.accepted.sum() + a.rejected.sum()), int(f.accepted.sum() + f.rejected.sum()))
        assert abs(int(f.accepted[ok].sum()) - int(a.accepted[ok].sum())) <= 2e-3 * int(a.accepted[ok].sum())
        if tset == 1:
            assert same.mean() >= 0.9 and np.quantile(dev, 0.99) <= 10.0
    # the user-source (NVRTC) flavour of the fast kernel
    rhs = deq.Rhs.from_source(USER_VDP, 2, 1)
    pu = deq.OdeProblem.builder().fun(rhs).params(mu[::8]).init(y0[::8]).tspan([0.0, 20.0]).build()
    fu = pu.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(*base, deq.Mode.Fast), tableau_set=1)
    au = pu.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(*base), tableau_set=1)
    du = np.max(np.abs(fu.final - au.final) / (np.abs(au.final) * rtol + atol), axis=1)
    assert (fu.accepted == au.accepted).mean() >= 0.9 and np.quantile(du, 0.99) <= 10.0


@pytest.mark.parametrize("which", ["kr4", "s4"])
@pytest.mark.parametrize("system", ["Lorenz", "VanDerPol", "Pendulum"])
def test_rosenbrock_ensembles(deq, oracle, which, system):
    """oderosenbrock reproduced bug-for-bug (it diverges on ordinary problems — inf/NaN rows must match too)."""
    from diffeq_b200 import synth
    n = 257
    if system == "Lorenz":
        y0, params = synth.lorenz_ics(0, n), np.array(LOR)
    elif system == "VanDerPol":
        y0, params = synth.vdp_sweep(0, n, n)
    else:
        y0, params = synth.pendulum_ics(0, n), synth.PENDULUM_PARAMS
    ts = deq.linspace(0.0, 0.3, 61)
    w = oracle.KR4 if which == "kr4" else oracle.S4
    m = deq.Ode.Ode4skr if which == "kr4" else deq.Ode.Ode4ss
    ref = oracle.ensemble_rosenbrock(w, getattr(oracle, system.upper()), params, y0, ts, want_all=True)
    prob = _problem(deq, getattr(deq.System, system), params, y0, ts)
    sol = prob.solve(m)
    assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True) and np.array_equal(sol.status, ref["status"])
    assert (sol.accepted == len(ts) - 1).all() and (sol.t_reached == ts[-1]).all()
    dense = prob.solve(m, deq.OdeOptionMap().insert(deq.Output.Tspan))
    assert np.array_equal(dense.yout, ref["yall"], equal_nan=True)
    assert len(dense.tout) == len(ts) - 1 and np.array_equal(dense.tout, np.diff(ts))   # problem.rs:572,639 (sic)
    assert dense.totals["accepted"] == n * (len(ts) - 1)


def test_invalid_matrix_is_a_per_trajectory_status(deq, oracle):
    """OdeError::InvalidMatrix (problem.rs:481,588) stops the failing trajectory only."""
    y0 = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 20.0], [0.0, 0.0, 0.0]])
    prm = np.array([[10.0, 0.0, -4.0], [10.0, 28.0, 8 / 3], [10.0, 0.0, -4.0]])
    h = float.fromhex("0x1.b504f333f9de6p-1")
    o = deq.OdeOptionMap().insert(deq.Initstep(h), deq.Maxstep(5.0))
    sol = _problem(deq, deq.System.Lorenz, prm, y0, [0.0, 10.0]).solve(deq.Ode.Ode23s, o)
    ref = oracle.ensemble_ode23s(oracle.LORENZ, prm, y0, [0.0, 10.0], oracle.opts(initstep=h, maxstep=5.0))
    assert sol.status.tolist() == [3, 0, 3] == ref["status"].tolist()
    assert np.array_equal(sol.final, ref["yfinal"]) and np.array_equal(sol.nfe, ref["nfe"])
    ts = [0.0, 0.25, 0.75, 9.0]
    sr = _problem(deq, deq.System.Lorenz, prm, y0, ts).solve(deq.Ode.Ode4ss, deq.OdeOptionMap().insert(deq.Output.Tspan))
    rr = oracle.ensemble_rosenbrock(oracle.S4, oracle.LORENZ, prm, y0, ts, want_all=True)
    assert sr.status.tolist() == [3, 0, 3] == rr["status"].tolist() and sr.nvalid.tolist() == [2, 4, 2]
    assert np.array_equal(sr.final, rr["yfinal"], equal_nan=True)
    assert np.array_equal(sr.yout[1], rr["yall"][1], equal_nan=True) and np.array_equal(sr.yout[0][:2], rr["yall"][0][:2])


USER_VDP = r'''
__device__ void rhs(double t, const double* v, double* d, const double* p) {
    d[0] = v[1];
    d[1] = p[0] * (1.0 - v[0] * v[0]) * v[1] - v[0];
}
'''
USER_ROBER_LIKE = r'''
// one-dimensional stiff decay towards cos(t): y' = -p0 (y - cos t)
__device__ void rhs(double t, const double* y, double* d, const double* p) { d[0] = -p[0] * (y[0] - deq_cos(t)); }
'''


def test_stiff_nvrtc_user_rhs(deq, oracle):
    """User CUDA source runs through the same stiff templates: bit-exact vs the built-in system and the oracle."""
    n = 200
    mu = np.geomspace(0.2, 500.0, n)[:, None]
    y0 = np.tile([2.0, 0.0], (n, 1))
    rhs = deq.Rhs.from_source(USER_VDP, 2, 1)
    ts = deq.linspace(0.0, 15.0, 31)
    ref = oracle.ensemble_ode23s(oracle.VANDERPOL, mu, y0, ts, oracle.opts(), dense=True)
    sol = deq.OdeProblem.builder().fun(rhs).params(mu).init(y0).tspan(ts).build().solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Output.Tspan))
    assert np.array_equal(sol.final, ref["yfinal"]) and np.array_equal(sol.accepted, ref["accepted"])
    assert np.array_equal(sol.yout, ref["out"].transpose(2, 1, 0))
    tr = np.linspace(0.0, 0.2, 21)
    rr = oracle.ensemble_rosenbrock(oracle.KR4, oracle.VANDERPOL, mu, y0, tr)
    sr = deq.OdeProblem.builder().fun(rhs).params(mu).init(y0).tspan(tr).build().solve(deq.Ode.Ode4skr)
    assert np.array_equal(sr.final, rr["yfinal"], equal_nan=True)
    # n = 1 (the reference skips the LU for a 1x1 Jacobian, problem.rs:467): TEXTBOOK and REFERENCE coincide and converge
    r1 = deq.Rhs.from_source(USER_ROBER_LIKE, 1, 1)
    p1 = deq.OdeProblem.builder().fun(r1).params([1000.0]).init([[0.0]]).tspan([0.0, 3.0]).build()
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-7), deq.Abstol(1e-10))
    a, b = p1.solve(deq.Ode.Ode23s, o, tableau_set=0), p1.solve(deq.Ode.Ode23s, o, tableau_set=1)
    assert np.array_equal(a.final, b.final) and int(a.status[0]) == 0
    lam, t = 1000.0, 3.0   # exact: y = (lam^2 cos t + lam sin t) / (1 + lam^2) + C e^{-lam t}
    exact = (lam * lam * np.cos(t) + lam * np.sin(t)) / (1 + lam * lam)
    assert abs(a.final[0, 0] - exact) < 1e-5


USER_CVDP = r'''
// DEQ_USER_N / 2 Van der Pol oscillators in a ring (the oracle's CoupledVdpF): same expression, same order
__device__ void rhs(double t, const double* v, double* d, const double* p) {
    const int M = DEQ_USER_N / 2;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const double x = v[2 * i], y = v[2 * i + 1], xn = v[2 * ((i + 1) % M)];
        d[2 * i] = y;
        d[2 * i + 1] = p[0] * (1.0 - x * x) * y - x + p[1] * (xn - x);
    }
}
'''


@pytest.mark.parametrize("n", [4, 6, 12, 24])
def test_stiff_beyond_closed_form_inverses(deq, oracle, n):
    """n = 4 goes through nalgebra's 4x4 cofactor inverse (`do_inverse4`), n >= 5 through its LU-based `try_inverse`
    (problem.rs:481,588 on a DMatrix of any size; n = 12 and 24 with the matrices in thread-local memory on the device): ode23s in
    both W-inverse variants (final states, counters, dense rows) and both Rosenbrock coefficient sets against the oracle, bit for
    bit; the committed pyref goldens pin the oracle."""
    osys = {4: oracle.CVDP2, 6: oracle.CVDP3, 12: oracle.CVDP6, 24: oracle.CVDP12}[n]
    rng = np.random.default_rng(n)
    m = 150 if n <= 6 else 60
    y0 = rng.uniform(-2.0, 2.0, (m, n))
    prm = np.column_stack([np.geomspace(0.5, 300.0, m), rng.uniform(0.0, 2.0, m)])
    rhs = deq.Rhs.from_source(USER_CVDP, n, 2)
    ts = deq.linspace(0.0, 12.0, 25)
    prob = deq.OdeProblem.builder().fun(rhs).params(prm).init(y0).tspan(ts).build()
    for tset in (0, 1):
        oo = oracle.opts(1e-5, 1e-8, max_steps=20000)
        ref = oracle.ensemble_ode23s(osys, prm, y0, ts, oo, tset=tset, dense=True)
        sol = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.MaxSteps(20000), deq.Output.Tspan), tableau_set=tset)
        for k in ("accepted", "rejected", "nfe", "status", "t_reached"):
            assert np.array_equal(getattr(sol, k), ref[k]), (tset, k)
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True), tset
        assert np.array_equal(sol.nvalid, ref["nvalid"]) and np.array_equal(sol.yout, ref["out"].transpose(2, 1, 0), equal_nan=True), tset
    assert (ref["status"] == 0).mean() > 0.9      # the textbook variant integrates the stiff end of the sweep
    tr = np.linspace(0.0, 0.2, 21)
    pr = deq.OdeProblem.builder().fun(rhs).params(prm).init(y0).tspan(tr).build()
    for which, ode in ((oracle.KR4, deq.Ode.Ode4skr), (oracle.S4, deq.Ode.Ode4ss)):
        rr = oracle.ensemble_rosenbrock(which, osys, prm, y0, tr, want_all=True)
        sr = pr.solve(ode, deq.OdeOptionMap().insert(deq.Output.Tspan))
        assert np.array_equal(sr.final, rr["yfinal"], equal_nan=True) and np.array_equal(sr.status, rr["status"]), which
        assert np.array_equal(sr.yout[::37], rr["yall"][::37], equal_nan=True)


def test_stiff_argument_checks(deq):
    src = "__device__ void rhs(double t, const double* y, double* dy, const double* p) { for (int i = 0; i < 33; ++i) dy[i] = -y[i]; }"
    p4 = deq.OdeProblem.builder().fun(deq.Rhs.from_source(src, 33, 0)).init([[1.0] * 33]).tspan([0.0, 1.0]).build()   # the stiff path carries n <= 32
    for m in (deq.Ode.Ode23s, deq.Ode.Ode4skr, deq.Ode.Ode4ss):
        with pytest.raises(deq.OdeError) as e:
            p4.solve(m)
        assert e.value.kind == "Unsupported"
    p = _problem(deq, deq.System.VanDerPol, [1.0], [2.0, 0.0], [0.0, 2.0, 1.0, 3.0])
    with pytest.raises(deq.OdeError):
        p.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Output.Tspan))     # non-monotone tspan with row output
    assert p.solve(deq.Ode.Ode23s).status[0] == 0                                   # final-state output does not care
    # empty tspan: nothing to solve (problem.rs:413-416, 567-570)
    e0 = _problem(deq, deq.System.VanDerPol, [1.0], [2.0, 0.0], [])
    assert e0.solve(deq.Ode.Ode23s).totals["accepted"] == 0 and e0.solve(deq.Ode.Ode4ss).totals["accepted"] == 0
    with pytest.raises(deq.OdeError):
        p.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Output.Tspan, deq.DenseLayout.TrajMajor))


def test_stiff_one_shot_host_call(deq, oracle):
    """deq_solve_host (the e2e entry point) carries the stiff methods, counters included."""
    n = 500
    mu = np.geomspace(0.1, 100.0, n)[:, None]
    y0 = np.tile([2.0, 0.0], (n, 1))
    ref = oracle.ensemble_ode23s(oracle.VANDERPOL, mu, y0, [0.0, 5.0], oracle.opts())
    out = deq.solve_host_multi(deq.Rhs.builtin(deq.System.VanDerPol), y0, mu, [0.0, 5.0], deq.Ode.Ode23s, deq.OdeOptionMap(), devices=[0, 0])
    assert np.array_equal(out["final"], ref["yfinal"]) and np.array_equal(out["accepted"], ref["accepted"])
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["t_reached"], ref["t_reached"])


def test_ode23s_textbook_accuracy_on_device(deq):
    """TEXTBOOK ode23s on the GPU against scipy Radau on a stiff Van der Pol sample."""
    from scipy.integrate import solve_ivp
    mus = [1.0, 30.0, 300.0]
    y0 = np.tile([2.0, 0.0], (3, 1))
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-9))
    sol = _problem(deq, deq.System.VanDerPol, np.array(mus)[:, None], y0, [0.0, 40.0]).solve(deq.Ode.Ode23s, o, tableau_set=1)
    for i, mu in enumerate(mus):
        ref = solve_ivp(lambda t, v: [v[1], mu * (1 - v[0] ** 2) * v[1] - v[0]], (0.0, 40.0), [2.0, 0.0], method="Radau", rtol=1e-10, atol=1e-12)
        assert abs(sol.final[i, 0] - ref.y[0, -1]) < 5e-3 * max(1.0, abs(ref.y[0, -1])), (mu, sol.final[i], ref.y[:, -1])


def test_ode23s_randomised_option_sweep_bit_exact(deq, oracle):
    """Seeded fuzz over systems, both W-inverse variants, libm flavours, tolerances, horizons, directions, step bounds, Points
    kinds and tspan grids (with repeated points): final states, counters and sampled (tout, yout) streams bit-identical."""
    from diffeq_b200 import synth
    rng = np.random.default_rng(20261021)
    for case in range(30):
        system = ["Lorenz", "VanDerPol", "Pendulum"][case % 3]
        n = int(rng.integers(1, 120))
        if system == "Lorenz":
            y0, params = synth.lorenz_ics(case * 1000, case * 1000 + n), np.array(LOR)
        elif system == "VanDerPol":
            y0, params = np.tile([2.0, 0.0], (n, 1)), 10.0 ** rng.uniform(-1, 2.5, (n, 1))
        else:
            y0, params = synth.pendulum_ics(case * 1000, case * 1000 + n), synth.PENDULUM_PARAMS
        tset, folded = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        rtol, atol = 10.0 ** rng.uniform(-7, -3), 10.0 ** rng.uniform(-10, -5)
        t0, span = float(rng.uniform(-1, 1)), float(10.0 ** rng.uniform(-1.5, 0.8))
        npts = int(rng.integers(2, 40))
        grid = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, npts - 2)]))
        if npts > 4 and rng.random() < 0.3:
            grid[2] = grid[1]                                   # a repeated tspan point
        ts = t0 + span * grid
        if rng.random() < 0.2:
            ts = ts[::-1].copy()                                # backward: no tspan row is ever emitted (problem.rs:520)
        pts = int(rng.random() < 0.4)
        kw = dict(max_steps=4000, points=pts, libm_folded=folded)
        o = deq.OdeOptionMap().insert(deq.Reltol(rtol), deq.Abstol(atol), deq.MaxSteps(4000), deq.Points.Specified if pts else deq.Points.All)
        if folded:
            o.insert(deq.Libm.LlvmFolded)
        if rng.random() < 0.3:
            kw["maxstep"] = span / float(rng.integers(3, 30)); o.insert(deq.Maxstep(kw["maxstep"]))
        if rng.random() < 0.3:
            kw["initstep"] = float(rng.choice([-1.0, 1.0]) * span * 10.0 ** rng.uniform(-5, -1)); o.insert(deq.Initstep(kw["initstep"]))
        if rng.random() < 0.2:
            kw["minstep"] = span * 10.0 ** rng.uniform(-6, -3); o.insert(deq.Minstep(kw["minstep"]))
        oo = oracle.opts(rtol, atol, **kw)
        ref = oracle.ensemble_ode23s(getattr(oracle, system.upper()), params, y0, ts, oo, tset=tset, dense=True)
        prob = _problem(deq, getattr(deq.System, system), params, y0, ts)
        tag = (case, system, n, tset, folded, rtol, atol, ts[0], ts[-1], kw)
        sol = prob.solve(deq.Ode.Ode23s, o, tableau_set=tset)
        for k in ("accepted", "rejected", "nfe", "status", "t_reached"):
            assert np.array_equal(getattr(sol, k), ref[k]), (k, tag)
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True), tag
        sa = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap(o).insert(deq.Output.All), tableau_set=tset)
        assert np.array_equal(np.diff(sa.offsets).astype(np.uint32), ref["nrows"]), tag
        for i in {0, n // 2, n - 1}:
            one = oracle.solve_ode23s(getattr(oracle, system.upper()), params[i] if np.ndim(params) == 2 else params, y0[i], ts, oo, tset=tset)
            tout, yout = sa.trajectory(i)
            assert np.array_equal(tout, one["tout"]) and np.array_equal(yout, one["yout"], equal_nan=True), (i, tag)
        sd = prob.solve(deq.Ode.Ode23s, deq.OdeOptionMap(o).insert(deq.Output.Tspan), tableau_set=tset)
        assert np.array_equal(sd.nvalid, ref["nvalid"]), tag
        want = ref["out"].transpose(2, 1, 0)
        m = np.arange(len(ts))[None, :] < ref["nvalid"][:, None]
        m &= ~np.isnan(want[:, :, 0])   # rows the reference never pushed (a tspan point equal to t0, a backward run) stay unset
        assert np.array_equal(sd.yout[m], want[m]), tag


def test_rosenbrock_randomised_grids_bit_exact(deq, oracle):
    """oderosenbrock on irregular, repeated (hs = 0 -> 1/(gamma hs) = inf) and descending tspan grids: every row bit-identical,
    inf / NaN included."""
    from diffeq_b200 import synth
    rng = np.random.default_rng(77)
    for case in range(16):
        system = ["Lorenz", "VanDerPol", "Pendulum"][case % 3]
        which = case % 2
        n = int(rng.integers(1, 90))
        if system == "Lorenz":
            y0, params = synth.lorenz_ics(case * 500, case * 500 + n), np.array(LOR)
        elif system == "VanDerPol":
            y0, params = np.tile([2.0, 0.0], (n, 1)), 10.0 ** rng.uniform(-1, 2, (n, 1))
        else:
            y0, params = synth.pendulum_ics(case * 500, case * 500 + n), synth.PENDULUM_PARAMS
        npts = int(rng.integers(1, 30))
        ts = np.cumsum(10.0 ** rng.uniform(-4, -1.5, npts)) + float(rng.uniform(-1, 1))
        if npts > 3 and rng.random() < 0.4:
            ts[2] = ts[1]
        if rng.random() < 0.25:
            ts = ts[::-1].copy()
        ref = oracle.ensemble_rosenbrock(which, getattr(oracle, system.upper()), params, y0, ts, want_all=True)
        sol = _problem(deq, getattr(deq.System, system), params, y0, ts).solve(
            deq.Ode.Ode4skr if which == 0 else deq.Ode.Ode4ss, deq.OdeOptionMap().insert(deq.Output.Tspan))
        tag = (case, system, which, n, npts)
        assert np.array_equal(sol.status, ref["status"]), tag
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True), tag
        m = np.arange(npts)[None, :] < sol.nvalid[:, None]
        assert np.array_equal(sol.yout[m], ref["yall"][m], equal_nan=True), tag
