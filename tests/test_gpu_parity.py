"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): fixed-step final states within 1e-12 relative, adaptive runs with identical
accepted-step counts and final states within 10 x rtol.  The parity kernels are built to do better — every assert
below is BIT-EXACT (np.array_equal on float64) in the default PARITY mode, including the pendulum (its sin/cos are
glibc's algorithm on the device); only the opt-in FAST mode is held to the north-star tolerance instead.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOR = [10.0, 28.0, 8.0 / 3.0]


def _problem(D, system, params, y0, tspan):
    return D.OdeProblem.builder().fun(D.Rhs.builtin(system)).params(params).init(y0).tspan(tspan).build()


def test_device_pow_log10_match_libm(deq, oracle):
    """deq_math.cuh on the device == glibc on the host (what the reference's f64::powf / log10 call)."""
    rng = np.random.default_rng(1)
    n = 400000
    x = np.exp(rng.uniform(-46, 46, n))
    x[::7] = 1.0 + rng.uniform(-0.5, 0.5, len(x[::7])) * 2.0 ** -rng.integers(0, 50, len(x[::7]))
    x[:6] = [0.0, np.inf, 1.0, 5e-324, 1e-310, 1e308]
    y = rng.choice([0.5, 0.2, 0.125, 1.0 / 3.0, 0.25], n)
    got = deq.device_pow(x, y)
    L = oracle.lib()
    ref = np.array([L.oracle_pow(float(a), float(b)) for a, b in zip(x[:50000], y[:50000])])
    assert np.array_equal(got[:50000].view(np.uint64), ref.view(np.uint64))
    p = rng.uniform(-30, 30, 50000)
    got = deq.device_pow(np.full_like(p, 10.0), p)
    ref = np.array([L.oracle_pow(10.0, float(b)) for b in p])
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    got = deq.device_log10(x[:50000])
    ref = np.array([L.oracle_log10(float(a)) for a in x[:50000]])
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))


def test_cfg1_rk4_lorenz_bit_exact(deq, oracle):
    """BASELINE config 1: single Lorenz, RK4 on linspace(0,100,100001) — the reference's own fixture
    (problem.rs:978-1010).  Every one of the 100001 stored states must equal the oracle's."""
    ts = deq.linspace(0.0, 100.0, 100001)
    assert np.array_equal(ts, oracle.linspace(0.0, 100.0, 100001))
    ref = oracle.solve_fixed(oracle.ODE4, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], ts)
    o = deq.OdeOptionMap().insert(deq.Output.Tspan)
    sol = _problem(deq, deq.System.Lorenz, LOR, [0.1, 0.0, 0.0], ts).solve(deq.Ode.Ode4, o)
    assert np.array_equal(sol.final[0], ref["yfinal"])
    assert np.array_equal(sol.yout[0], ref["yout"])
    assert sol.final[0].tolist() == [float.fromhex("0x1.12621c0124822p+1"), float.fromhex("-0x1.2b36df19c01edp+2"),
                                     float.fromhex("0x1.dd8c99d38f21ep+4")]


@pytest.mark.parametrize("method", ["Feuler", "Heun", "Midpoint", "Ode4"])
@pytest.mark.parametrize("system", ["Lorenz", "VanDerPol", "Pendulum"])
def test_fixed_methods_ensemble(deq, oracle, method, system):
    from diffeq_b200 import synth
    n = 257  # ragged: not a multiple of the block size
    if system == "Lorenz":
        y0, params = synth.lorenz_ics(0, n), np.array(LOR)
    elif system == "VanDerPol":
        y0, params = synth.vdp_sweep(0, n, n)
    else:
        y0, params = synth.pendulum_ics(0, n), synth.PENDULUM_PARAMS
    ts = deq.linspace(0.0, 2.0, 401)
    ref = oracle.ensemble_fixed(getattr(oracle, method.upper()), getattr(oracle, system.upper()), params, y0, ts)
    sol = _problem(deq, getattr(deq.System, system), params, y0, ts).solve(getattr(deq.Ode, method))
    assert np.array_equal(sol.final, ref)   # Pendulum included: sin/cos are glibc-exact on the device


def _adaptive_case(deq, oracle, method, system, y0, params, ts, rtol, atol, tset=0, exact=True, **kw):
    oo = oracle.opts(rtol, atol, **kw)
    ref = oracle.ensemble_adaptive(getattr(oracle, method.upper()), getattr(oracle, system.upper()), params, y0, ts, oo, tset)
    o = deq.OdeOptionMap().insert(deq.Reltol(rtol), deq.Abstol(atol))
    if "max_steps" in kw:
        o.insert(deq.MaxSteps(kw["max_steps"]))
    sol = _problem(deq, getattr(deq.System, system), params, y0, ts).solve(getattr(deq.Ode, method), o, tableau_set=tset)
    if exact:
        assert np.array_equal(sol.accepted, ref["accepted"])
        assert np.array_equal(sol.rejected, ref["rejected"])
        assert np.array_equal(sol.nfe, ref["nfe"])
        assert np.array_equal(sol.status, ref["status"])
        assert np.array_equal(sol.final, ref["yfinal"])
        assert np.array_equal(sol.t_reached, ref["t_reached"])
    else:
        same = (sol.accepted == ref["accepted"]) & (sol.rejected == ref["rejected"])
        assert same.mean() >= 0.9
        np.testing.assert_allclose(sol.final[same], ref["yfinal"][same], rtol=10 * rtol, atol=10 * atol)
    return sol, ref


def test_cfg2_dopri5_lorenz_ensemble_bit_exact(deq, oracle):
    """BASELINE config 2 at oracle-sized N: Lorenz ICs from the counter RNG, dopri5, rtol=atol=1e-8, tspan=[0,10]."""
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 4096 + 37)
    sol, ref = _adaptive_case(deq, oracle, "Ode45", "Lorenz", y0, LOR, [0.0, 10.0], 1e-8, 1e-8)
    assert sol.totals["accepted"] == int(ref["accepted"].sum())
    assert sol.totals["rejected"] == int(ref["rejected"].sum())
    assert sol.totals["nfe"] == int(ref["nfe"].sum())


def test_controller_division_sequences_match_ieee(deq):
    """div_main / rcp_main (the division and reciprocal of the step-size controller as straight-line instruction sequences,
    one shared out-of-range flag) against the IEEE operators: 2^30 hashed operand pairs — the full double range and the
    controller's magnitudes — bit for bit wherever the flag stays clear, and the flag stays clear on most of them."""
    mism, checked = deq.test_div(0xD1F, 1 << 30)
    assert mism == 0
    assert checked > 1.2 * (1 << 30)


def test_refill_and_static_scheduling_agree(deq, oracle):
    """Lane refill only changes which lane runs a trajectory, never its arithmetic (shard/schedule invariance)."""
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(100, 100 + 3000)
    p = _problem(deq, deq.System.Lorenz, LOR, y0, [0.0, 5.0])
    a = p.solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8), deq.Refill(1)))
    b = p.solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8), deq.Refill(0)))
    assert np.array_equal(a.final, b.final) and np.array_equal(a.accepted, b.accepted)
    # shard invariance: trajectory i gives the same result when solved in a different partition
    c = _problem(deq, deq.System.Lorenz, LOR, y0[1000:2000], [0.0, 5.0]).solve(
        deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8)))
    assert np.array_equal(a.final[1000:2000], c.final) and np.array_equal(a.accepted[1000:2000], c.accepted)


def test_single_lorenz_default_options_t100(deq, oracle):
    """Chaotic T=100 with default options (the reference's ode45_test fixture): 3357 accepted / 112 rejected."""
    sol, ref = _adaptive_case(deq, oracle, "Ode45", "Lorenz", np.array([[0.1, 0.0, 0.0]]), LOR, [0.0, 100.0], 1e-5, 1e-8)
    assert int(sol.accepted[0]) == 3357 and int(sol.rejected[0]) == 112


@pytest.mark.parametrize("method,tset", [("Ode45fe", 0), ("Ode45", 1), ("Ode23", 1), ("Ode21", 1), ("Ode78", 1)])
def test_adaptive_methods_lorenz(deq, oracle, method, tset):
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 515)
    tol = {"Ode21": 1e-3, "Ode23": 1e-5}.get(method, 1e-8)
    _adaptive_case(deq, oracle, method, "Lorenz", y0, LOR, [0.0, 2.0], tol, tol, tset=tset)


@pytest.mark.parametrize("method", ["Ode21", "Ode23", "Ode78"])
def test_broken_reference_tableaus_short_horizon(deq, oracle, method):
    """The reference's rk21 / rk23 / feh78 are mis-transcribed: the step size collapses.  Parity is
    checked bug-for-bug on a short horizon with the max_steps guard (status MAX_STEPS where the collapse hits)."""
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 130)
    _adaptive_case(deq, oracle, method, "Lorenz", y0, LOR, [0.0, 0.01], 1e-3, 1e-3, tset=0, max_steps=3000)


def test_vdp_sweep_rk45_per_trajectory_params(deq, oracle):
    from diffeq_b200 import synth
    y0, mu = synth.vdp_sweep(0, 1024, 1024)
    _adaptive_case(deq, oracle, "Ode45fe", "VanDerPol", y0, mu, [0.0, 20.0], 1e-6, 1e-8)


def test_vdp_dense_output_matches_oracle(deq, oracle):
    """BASELINE config 3 at oracle-sized N: Hermite output at linspace(0,20,1000), SoA, bit-exact incl. nvalid."""
    from diffeq_b200 import synth
    n = 300
    y0, mu = synth.vdp_sweep(0, n, n)
    ts = deq.linspace(0.0, 20.0, 1000)
    ref = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu, y0, ts, oracle.opts(1e-6, 1e-8))
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan)
    sol = _problem(deq, deq.System.VanDerPol, mu, y0, ts).solve(deq.Ode.Ode45fe, o)
    assert np.array_equal(sol.nvalid, ref["nvalid"])
    want = ref["out"].transpose(2, 1, 0)
    assert np.array_equal(np.isnan(sol.yout), np.isnan(want))
    m = ~np.isnan(want)
    assert np.array_equal(sol.yout[m], want[m])
    assert np.array_equal(sol.accepted, ref["accepted"])


@pytest.mark.parametrize("order", ["sorted", "permuted"])
def test_soa_dense_rows_any_order(deq, oracle, order):
    """SoA dense rows from the row-synchronous kernel (default: a warp owns 32 consecutive trajectories and emits each row
    together) and from the lane-refill kernel (Refill(1): every lane emits inside its own steps), on a parameter sweep in
    sorted and in randomly permuted order: both equal the oracle bit for bit — rows, nvalid, counters, final states — for the
    forward run, a backward run, a tspan whose points are hit exactly by step ends (the reference stops emitting there),
    an early stop (max_steps) and the FSAL / non-FSAL tableaus."""
    from diffeq_b200 import synth
    n = 1037
    y0, mu = synth.vdp_sweep(0, n, n)
    if order == "permuted":
        perm = np.random.default_rng(7).permutation(n)
        mu = mu[perm].copy()
    cases = [("Ode45fe", 0, deq.linspace(0.0, 20.0, 200), 1e-6, dict()),
             ("Ode45", 1, deq.linspace(0.0, 6.0, 61), 1e-6, dict()),                       # FSAL
             ("Ode45fe", 0, deq.linspace(3.0, -2.0, 41), 1e-6, dict(max_steps=400)),       # backward (the reference runs away after a rejection)
             ("Ode23", 1, np.arange(9) * 0.25, 0.5, dict(initstep=0.25, maxstep=0.25)),    # step ends fall exactly on tspan points
             ("Ode45fe", 0, deq.linspace(0.0, 20.0, 50), 1e-6, dict(max_steps=60))]        # trajectories stop early
    quirk = False
    for method, tset, ts, tol, kw in cases:
        atol = 1e-8 if tol < 0.1 else tol
        oo = oracle.opts(tol, atol, **kw)
        ref = oracle.ensemble_dense(getattr(oracle, method.upper()), oracle.VANDERPOL, mu, y0, ts, oo, tset)
        want = ref["out"].transpose(2, 1, 0)
        for refill in (2, 1, 0):
            o = deq.OdeOptionMap().insert(deq.Reltol(tol), deq.Abstol(atol), deq.Output.Tspan, deq.Refill(refill))
            if "max_steps" in kw: o.insert(deq.MaxSteps(kw["max_steps"]))
            if "initstep" in kw: o.insert(deq.Initstep(kw["initstep"]), deq.Maxstep(kw["maxstep"]))
            sol = _problem(deq, deq.System.VanDerPol, mu, y0, ts).solve(getattr(deq.Ode, method), o, tableau_set=tset)
            assert np.array_equal(sol.nvalid, ref["nvalid"]), (method, refill)
            assert np.array_equal(sol.yout, want, equal_nan=True), (method, refill)
            assert np.array_equal(sol.accepted, ref["accepted"]) and np.array_equal(sol.rejected, ref["rejected"]), (method, refill)
            assert np.array_equal(sol.status, ref["status"]), (method, refill)
        quirk = quirk or bool(((ref["nvalid"] < len(ts) - 1) & (ref["status"] == 0)).any())
    assert quirk                        # some complete run stopped emitting because a step end hit a tspan point exactly
    assert (ref["nvalid"] < 50).any()   # the early-stop case really leaves rows unwritten


def test_soa_dense_rows_sorted_by_first_step_large(deq, oracle):
    """A sweep handed over in random order, 70 001 trajectories: the sorted SoA path of deq_solve (work items ordered by first
    step size, lane-refill kernel through the order table, columns put back in place in several batches) — rows, nvalid, counters
    and final states equal the oracle's, for a ragged trajectory count, forward and with an early stop (rows left unwritten)."""
    from diffeq_b200 import synth
    n = 70001
    y0, mu = synth.vdp_sweep(0, n, n)
    perm = np.random.default_rng(11).permutation(n)
    mu = mu[perm].copy()
    for ts, kw in ((deq.linspace(0.0, 2.0, 14), dict()), (deq.linspace(0.0, 6.0, 9), dict(max_steps=40))):
        oo = oracle.opts(1e-6, 1e-8, **kw)
        ref = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu, y0, ts, oo, 0)
        want = ref["out"].transpose(2, 1, 0)
        for lay in (deq.DenseLayout.SoA, deq.DenseLayout.TrajMajor):   # (trajectory-major: same dispatch, rows written straight to their streams)
            o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan, lay)
            if "max_steps" in kw: o.insert(deq.MaxSteps(kw["max_steps"]))
            sol = _problem(deq, deq.System.VanDerPol, mu, y0, ts).solve(deq.Ode.Ode45fe, o)
            assert np.array_equal(sol.nvalid, ref["nvalid"]), int(lay)
            msk = np.arange(len(ts))[None, :] < np.maximum(ref["nvalid"], 1)[:, None]
            msk[:, -1] = True
            assert np.array_equal(sol.yout[msk], want[msk]), int(lay)
            assert np.array_equal(sol.accepted, ref["accepted"]) and np.array_equal(sol.rejected, ref["rejected"]), int(lay)
            assert np.array_equal(sol.status, ref["status"]), int(lay)
            assert np.array_equal(sol.final, want[:, -1, :]), int(lay)
    assert (ref["nvalid"] < 8).any()


def test_pendulum_feh78_textbook_bit_exact(deq, oracle):
    """BASELINE config 4 at oracle-sized N.  The pendulum calls sin/cos; the device evaluates glibc's algorithm
    (deq_sincos.cuh), so even at rtol = atol = 1e-12 every count and every final state equals the oracle's."""
    from diffeq_b200 import synth
    y0 = synth.pendulum_ics(0, 2048)
    _adaptive_case(deq, oracle, "Ode78", "Pendulum", y0, synth.PENDULUM_PARAMS, [0.0, 10.0], 1e-12, 1e-12, tset=1)
    _adaptive_case(deq, oracle, "Ode45", "Pendulum", y0[:512], synth.PENDULUM_PARAMS, [0.0, 30.0], 1e-8, 1e-8, tset=1)
    _adaptive_case(deq, oracle, "Ode45", "Pendulum", y0[:512], synth.PENDULUM_PARAMS, [0.0, 3.0], 1e-6, 1e-8, tset=0)


def test_backward_integration_and_options(deq, oracle):
    """Backward integration: the reference's timeout clamp `new_dt.min(dt)` (problem.rs:835-838) flips the sign of dt after
    the first rejection and the run never ends; reproduced bug-for-bug under the max_steps guard."""
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 200)
    for ts, kw in (([1.0, 0.0], dict()), ([1.0, 0.0], dict(initstep=-1e-3)), ([1.0, 0.0], dict(maxstep=0.01)),
                   ([1.0, 0.0], dict(minstep=1e-3)), ([0.02, 0.0], dict()), ([0.0, 1.0], dict(initstep=1e-4, maxstep=0.02))):
        oo = oracle.opts(1e-6, 1e-8, max_steps=2500, **kw)
        ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0, ts, oo)
        o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.MaxSteps(2500))
        if "initstep" in kw: o.insert(deq.Initstep(kw["initstep"]))
        if "maxstep" in kw: o.insert(deq.Maxstep(kw["maxstep"]))
        if "minstep" in kw: o.insert(deq.Minstep(kw["minstep"]))
        sol = _problem(deq, deq.System.Lorenz, LOR, y0, ts).solve(deq.Ode.Ode45, o)
        assert np.array_equal(sol.final, ref["yfinal"]), (ts, kw)
        assert np.array_equal(sol.accepted, ref["accepted"]) and np.array_equal(sol.status, ref["status"]), (ts, kw)
        assert np.array_equal(sol.t_reached, ref["t_reached"]), (ts, kw)


def test_error_behaviour(deq):
    p = _problem(deq, deq.System.Lorenz, LOR, [0.1, 0.0, 0.0], [0.0, 1.0])
    with pytest.raises(deq.OdeError) as e:
        p.solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Initstep(-0.1)))   # problem.rs:232-237
    assert e.value.kind == "InvalidInitstep"
    with pytest.raises(deq.OdeError) as e:
        p.solve(deq.Ode.Ode23s, deq.OdeOptionMap().insert(deq.Precision.F32))   # the stiff solvers are f64 only
    assert e.value.kind == "Unsupported"
    with pytest.raises(deq.OdeError) as e:
        deq.OdeProblem.builder().fun(deq.Rhs.builtin(0)).params(LOR).init([0.1, 0, 0]).build()  # problem.rs:99-101
    assert e.value.kind == "Uninitialized"
    # empty tspan: adaptive returns an empty solution (problem.rs:211-214)
    sol = _problem(deq, deq.System.Lorenz, LOR, [0.1, 0.0, 0.0], []).solve(deq.Ode.Ode45)
    assert sol.totals["accepted"] == 0
    # zero time span: one accepted zero-length step
    sol = _problem(deq, deq.System.Lorenz, LOR, [0.1, 0.0, 0.0], [1.0, 1.0]).solve(deq.Ode.Ode45)
    assert int(sol.accepted[0]) == 1


USER_LORENZ = r'''
// same operation order as the reference fixture (problem.rs:989-1000)
__device__ void rhs(double t, const double* v, double* d, const double* p) {
    double x = v[0], y = v[1], z = v[2];
    d[0] = p[0] * (y - x);
    d[1] = x * (p[1] - z) - y;
    d[2] = x * y - p[2] * z;
}
'''
USER_VDP = r'''
__device__ void rhs(double t, const double* v, double* d, const double* p) {
    d[0] = v[1];
    d[1] = p[0] * (1.0 - v[0] * v[0]) * v[1] - v[0];
}
'''


def test_nvrtc_user_rhs_matches_oracle(deq, oracle):
    """User CUDA source through NVRTC runs the same stepper templates: bit-exact vs the oracle, all output modes."""
    from diffeq_b200 import synth
    rhs = deq.Rhs.from_source(USER_LORENZ, 3, 3)
    assert (rhs.dim, rhs.nparams) == (3, 3)
    y0 = synth.lorenz_ics(0, 700)
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8))
    sol = deq.OdeProblem.builder().fun(rhs).params(LOR).init(y0).tspan([0.0, 3.0]).build().solve(deq.Ode.Ode45, o)
    ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0, [0.0, 3.0], oracle.opts(1e-8, 1e-8))
    assert np.array_equal(sol.final, ref["yfinal"]) and np.array_equal(sol.accepted, ref["accepted"])
    # fixed step, full solution of one trajectory
    ts = deq.linspace(0.0, 1.0, 1001)
    fx = deq.OdeProblem.builder().fun(rhs).params(LOR).init([0.1, 0.0, 0.0]).tspan(ts).build().solve(
        deq.Ode.Ode4, deq.OdeOptionMap().insert(deq.Output.Tspan))
    rf = oracle.solve_fixed(oracle.ODE4, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], ts)
    assert np.array_equal(fx.yout[0], rf["yout"])
    # per-trajectory parameters + dense output
    vd = deq.Rhs.from_source(USER_VDP, 2, 1)
    y0v, mu = synth.vdp_sweep(0, 200, 200)
    tsd = deq.linspace(0.0, 20.0, 300)
    od = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan)
    sd = deq.OdeProblem.builder().fun(vd).params(mu).init(y0v).tspan(tsd).build().solve(deq.Ode.Ode45fe, od)
    rd = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu, y0v, tsd, oracle.opts(1e-6, 1e-8))
    want = rd["out"].transpose(2, 1, 0)
    m = ~np.isnan(want)
    assert np.array_equal(sd.nvalid, rd["nvalid"]) and np.array_equal(sd.yout[m], want[m])


USER_PENDULUM_FORCING = """
#define DEQ_USER_NG 1
__device__ void forcing(double t, double* g, const double* p) { g[0] = deq_cos(p[2] * t); }
__device__ void rhs_g(double t, const double* v, double* d, const double* p, const double* g) {
    d[0] = v[1]; d[1] = -p[0] * v[1] - deq_sin(v[0]) + p[1] * g[0]; }
__device__ void rhs(double t, const double* v, double* d, const double* p) {
    d[0] = v[1]; d[1] = -p[0] * v[1] - deq_sin(v[0]) + p[1] * deq_cos(p[2] * t); }
"""


def test_nvrtc_user_rhs_with_time_only_forcing(deq, oracle):
    """A user right-hand side may declare its time-only terms (DEQ_USER_NG, forcing, rhs_g): the stepper evaluates them once per
    distinct stage time (rows of equal c, the end of an accepted step, the carried value at c = 0).  Same bits as the oracle,
    which evaluates cos(Omega t) at every stage: Fehlberg 7(8) (both coefficient sets: duplicate c rows, a c = 0 row), RK4
    (c1 = c2), dopri5 (FSAL), dense rows, backward, and a -0.0 start time (the carried value must not be used there)."""
    from diffeq_b200 import synth
    rhs = deq.Rhs.from_source(USER_PENDULUM_FORCING, 2, 3)
    P = list(synth.PENDULUM_PARAMS)
    y0 = synth.pendulum_ics(0, 333)
    for method, tset, ts, tol in (("Ode78", 1, [0.0, 10.0], 1e-12), ("Ode78", 0, [0.0, 3.0], 1e-7), ("Ode45", 0, [0.0, 10.0], 1e-9),
                                  ("Ode78", 1, [-0.0, 4.0], 1e-10), ("Ode78", 1, [2.0, -1.0], 1e-9)):
        # (bounded: the reference's own Fehlberg 7(8) coefficients collapse the step size, and a backward run never ends — DESIGN.md §6)
        kw = dict(max_steps=300 if (ts[1] < ts[0] or tset == 0 and method == "Ode78") else 5000)
        o = deq.OdeOptionMap().insert(deq.Reltol(tol), deq.Abstol(tol), deq.MaxSteps(kw["max_steps"]))
        sol = deq.OdeProblem.builder().fun(rhs).params(P).init(y0).tspan(ts).build().solve(getattr(deq.Ode, method), o, tableau_set=tset)
        ref = oracle.ensemble_adaptive(getattr(oracle, method.upper()), oracle.PENDULUM, P, y0, ts, oracle.opts(tol, tol, **kw), tset)
        for k in ("accepted", "rejected", "nfe", "status"):
            assert np.array_equal(getattr(sol, k), ref[k]), (method, tset, ts, k)
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True), (method, tset, ts)
    ts = deq.linspace(0.0, 5.0, 41)
    fx = deq.OdeProblem.builder().fun(rhs).params(P).init(y0).tspan(ts).build().solve(deq.Ode.Ode4, deq.OdeOptionMap().insert(deq.Output.Tspan))
    assert np.array_equal(fx.final, oracle.ensemble_fixed(oracle.ODE4, oracle.PENDULUM, P, y0, ts))
    tsd = deq.linspace(0.0, 6.0, 25)
    for lay in (deq.DenseLayout.SoA, deq.DenseLayout.TrajMajor):
        sd = deq.OdeProblem.builder().fun(rhs).params(P).init(y0).tspan(tsd).build().solve(
            deq.Ode.Ode78, deq.OdeOptionMap().insert(deq.Reltol(1e-9), deq.Abstol(1e-9), deq.Output.Tspan, lay), tableau_set=1)
        rd = oracle.ensemble_dense(oracle.ODE78, oracle.PENDULUM, P, y0, tsd, oracle.opts(1e-9, 1e-9), 1)
        want = rd["out"].transpose(2, 1, 0)
        m = ~np.isnan(want) & ~np.isnan(sd.yout)
        assert np.array_equal(sd.nvalid, rd["nvalid"]) and np.array_equal(sd.yout[m], want[m])


def test_nvrtc_same_source_reuses_compiled_kernel(deq):
    """A second deq_rhs from the same source (the first one already destroyed) hits the process-wide kernel cache:
    same results, and the second solve is not slowed by a recompile."""
    import time
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(3, 256)
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-7), deq.Abstol(1e-9))

    def run():
        rhs = deq.Rhs.from_source(USER_LORENZ + "// cache-test\n", 3, 3)
        t0 = time.perf_counter()
        s = deq.OdeProblem.builder().fun(rhs).params(LOR).init(y0).tspan([0.0, 2.0]).build().solve(deq.Ode.Ode45fe, o)
        dt = time.perf_counter() - t0
        fin = s.final.copy()
        del s, rhs
        return fin, dt

    a, ta = run()
    b, tb = run()
    assert np.array_equal(a, b)
    assert tb < 0.5 * ta, (ta, tb)      # first call pays the NVRTC compile (hundreds of ms), the second does not


def test_nvrtc_compile_error_is_reported(deq):
    with pytest.raises(deq.OdeError) as e:
        deq.Rhs.from_source("__device__ void rhs(double t, const double* y, double* dy, const double* p) { dy[0] = undefined_symbol; }", 1, 0)
    assert e.value.kind == "Nvrtc" and "undefined_symbol" in str(e.value)


def test_nvrtc_generic_dimension_converges(deq):
    """A 4-dimensional user system with no parameters (two uncoupled harmonic oscillators): exact solution known."""
    src = "__device__ void rhs(double t, const double* y, double* dy, const double* p) { dy[0] = y[1]; dy[1] = -y[0]; dy[2] = 2.0 * y[3]; dy[3] = -2.0 * y[2]; }"
    rhs = deq.Rhs.from_source(src, 4, 0)
    y0 = np.tile(np.array([1.0, 0.0, 0.0, 1.0]), (33, 1))
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-10), deq.Abstol(1e-10))
    sol = deq.OdeProblem.builder().fun(rhs).init(y0).tspan([0.0, 2.0]).build().solve(deq.Ode.Ode45, o, tableau_set=deq.TableauSet.Textbook)
    want = np.array([np.cos(2.0), -np.sin(2.0), np.sin(4.0), np.cos(4.0)])
    np.testing.assert_allclose(sol.final, np.tile(want, (33, 1)), rtol=0, atol=1e-8)


@pytest.mark.parametrize("method,system,T,tol,tset", [("Ode45", "Lorenz", 10.0, 1e-8, 0), ("Ode45fe", "VanDerPol", 20.0, 1e-6, 0),
                                                       ("Ode78", "Lorenz", 2.0, 1e-10, 1)])
def test_fast_mode_within_north_star_tolerance(deq, oracle, method, system, T, tol, tset):
    """Mode.Fast (FMA contraction, zero skipping, one pow per step) is NOT bit-exact; it is held to BASELINE.json's
    floating-point bar against the oracle: identical accepted-step counts and final states within 10 x rtol."""
    from diffeq_b200 import synth
    n = 8192
    if system == "Lorenz":
        y0, params = synth.lorenz_ics(0, n), np.array(LOR)
    else:
        y0, params = synth.vdp_sweep(0, n, n)
    ref = oracle.ensemble_adaptive(getattr(oracle, method.upper()), getattr(oracle, system.upper()), params, y0, [0.0, T],
                                   oracle.opts(tol, tol), tset)
    o = deq.OdeOptionMap().insert(deq.Reltol(tol), deq.Abstol(tol), deq.Mode.Fast)
    sol = _problem(deq, getattr(deq.System, system), params, y0, [0.0, T]).solve(getattr(deq.Ode, method), o, tableau_set=tset)
    same = (sol.accepted == ref["accepted"]) & (sol.rejected == ref["rejected"])
    err = np.abs(sol.final - ref["yfinal"]) / (np.abs(ref["yfinal"]) * tol + tol)   # in units of the step tolerance
    print(f"fast-mode {method}/{system}: identical step counts {same.mean():.6f}, max |dy|/(rtol|y|+atol) = {err.max():.3e}")
    assert same.mean() >= 0.999
    assert err.max() <= 10.0
    assert np.array_equal(sol.status, ref["status"])


def test_fast_mode_accuracy_against_truth(deq, oracle):
    """Mode.Fast cannot be held to "10 x rtol of PARITY" on every chaotic trajectory (Lorenz at T = 10 amplifies ANY rounding
    difference by up to ~1e10), so it is held to the truth instead: against a solution that is five orders of magnitude more
    accurate (Fehlberg 7(8), textbook coefficients, rtol = atol = 1e-13 — anchored on scipy's DOP853 for the first trajectories)
    the FAST result of every trajectory is as accurate as the bit-exact PARITY result (whose error, amplified by the chaotic
    dynamics, is ~6e3 x (rtol |y| + atol) in the median), and the step counts agree."""
    from scipy.integrate import solve_ivp
    from diffeq_b200 import synth
    n, tol = 256, 1e-8
    y0, params = synth.lorenz_ics(0, n), np.array(LOR)
    truth = oracle.ensemble_adaptive(oracle.ODE78, oracle.LORENZ, params, y0, [0.0, 10.0], oracle.opts(1e-13, 1e-13, max_steps=2_000_000),
                                     oracle.TEXTBOOK)["yfinal"]
    f = lambda t, y: [LOR[0] * (y[1] - y[0]), y[0] * (LOR[1] - y[2]) - y[1], y[0] * y[1] - LOR[2] * y[2]]
    for i in range(4):   # the truth itself: an independent integrator agrees far below the errors compared here
        s = solve_ivp(f, [0.0, 10.0], y0[i], method="DOP853", rtol=1e-13, atol=1e-13)
        assert np.max(np.abs(s.y[:, -1] - truth[i]) / (np.abs(truth[i]) * tol + tol)) < 50.0   # PARITY's own error: median ~6e3 of these units
    prob = _problem(deq, deq.System.Lorenz, params, y0, [0.0, 10.0])
    par = prob.solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(tol), deq.Abstol(tol), deq.Mode.Parity))
    fast = prob.solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(tol), deq.Abstol(tol), deq.Mode.Fast))
    unit = np.abs(truth) * tol + tol
    e_par = np.max(np.abs(par.final - truth) / unit, axis=1)
    e_fast = np.max(np.abs(fast.final - truth) / unit, axis=1)
    print(f"error vs truth in units of rtol|y|+atol: PARITY median {np.median(e_par):.3g} max {e_par.max():.3g}; "
          f"FAST median {np.median(e_fast):.3g} max {e_fast.max():.3g}; max ratio {np.max(e_fast / np.maximum(e_par, 1e-3)):.4f}")
    assert np.all(e_fast <= 1.5 * e_par + 50.0)   # 50 units: the uncertainty of the truth itself (see above)
    assert np.median(e_fast) <= 1.05 * np.median(e_par)
    assert np.mean((fast.accepted == par.accepted) & (fast.rejected == par.rejected)) >= 0.99


def test_cpp_host_mirror_on_gpu(deq, oracle):
    """The header-only C++ mirror (include/diffeq_b200.hpp) drives the same C ABI: result equals the oracle's."""
    import subprocess, os
    from tests.test_host_cpu import _build_hpp_smoke
    out = subprocess.run([_build_hpp_smoke()], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    tok = out.stdout.split()
    ref = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], [0.0, 1.0], oracle.opts(1e-8, 1e-8))
    assert tok[0] == "HPP_OK" and int(tok[1]) == 1000 and int(tok[2]) == ref["accepted"] and float(tok[3]) == ref["yfinal"][0]
    assert "HPP_ERR 1" in out.stdout   # Uninitialized, like OdeBuilder::build (problem.rs:92-104)
    # solve_trajectories: the reference-shaped (tout, yout) of each problem
    tr = [l for l in out.stdout.splitlines() if l.startswith("HPP_TRAJ")][0].replace("|", " ").split()[1:]
    ts3 = [0.0, 0.5, 1.0]
    a = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], ts3, oracle.opts(1e-8, 1e-8))
    f = oracle.solve_fixed(oracle.ODE4, oracle.LORENZ, LOR, [1.0, 2.0, 20.0], ts3)
    s3 = oracle.solve_ode23s(oracle.LORENZ, LOR, [0.1, 0.0, 0.0], ts3, oracle.opts())
    k = oracle.solve_rosenbrock(oracle.S4, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], ts3)
    assert (int(tr[0]), int(tr[1]), int(tr[2]), float(tr[3])) == (a["nout"], a["nout"], a["accepted"], a["yout"][-1][0])
    assert (int(tr[4]), float(tr[5])) == (3, f["yout"][-1][2])
    assert (int(tr[6]), int(tr[7]), float(tr[8])) == (s3["nout"], s3["accepted"], s3["yout"][-1][0])
    assert (int(tr[9]), int(tr[10])) == (len(k["tout"]), len(k["yout"])) == (2, 3)


def test_ragged_points_all_output_matches_reference_shape(deq, oracle):
    """Output.All = the reference's complete Points::All result per trajectory (y0, Hermite values at tspan points inside
    each step, every step end point; problem.rs:301-323), ragged, bit-exact vs the oracle's (tout, yout)."""
    from diffeq_b200 import synth
    n = 257
    y0, mu = synth.vdp_sweep(0, n, n)
    ts = deq.linspace(0.0, 20.0, 100)
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.All)
    sol = _problem(deq, deq.System.VanDerPol, mu, y0, ts).solve(deq.Ode.Ode45fe, o)
    for i in (0, 1, 100, 256):
        ref = oracle.solve_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu[i], y0[i], ts, oracle.opts(1e-6, 1e-8))
        tout, yout = sol.trajectory(i)
        assert len(tout) == ref["nout"] == 1 + int(sol.accepted[i]) + (int(sol.offsets[i + 1] - sol.offsets[i]) - 1 - int(sol.accepted[i]))
        assert np.array_equal(tout, ref["tout"]) and np.array_equal(yout, ref["yout"])
    # single Lorenz trajectory with a 2-point tspan: tout.len() == 1 + accepted (SURVEY §5 "observable proxy")
    s1 = _problem(deq, deq.System.Lorenz, LOR, [0.1, 0.0, 0.0], [0.0, 10.0]).solve(
        deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8), deq.Output.All))
    r1 = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, LOR, [0.1, 0.0, 0.0], [0.0, 10.0], oracle.opts(1e-8, 1e-8))
    t1, y1 = s1.trajectory(0)
    assert len(t1) == 947 and np.array_equal(t1, r1["tout"]) and np.array_equal(y1, r1["yout"])


def test_demo_json_entry_point(deq, oracle):
    """diffeq-example-wasm's `solve_lorenz_attractor` (lib.rs:38-65) on the GPU backend: same JSON in, serde-shaped
    OdeSolution out, values equal to the oracle's for an adaptive and a fixed-step method."""
    import json
    from diffeq_b200 import demo
    cfg = {"ode": "ode45", "init": [0.1, 0.0, 0.0], "start_time": 0.0, "end_time": 5.0, "num_times": 101}
    out = json.loads(demo.solve_lorenz_attractor(json.dumps(cfg)))
    ts = oracle.linspace(0.0, 5.0, 101)
    ref = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, LOR, cfg["init"], ts)
    assert list(out) == ["tout", "yout"] and out["tout"] == ref["tout"].tolist() and out["yout"] == ref["yout"].tolist()
    cfg["ode"] = "ode4"
    out = json.loads(demo.solve_lorenz_attractor(json.dumps(cfg)))
    rf = oracle.solve_fixed(oracle.ODE4, oracle.LORENZ, LOR, cfg["init"], ts)
    assert out["tout"] == ts.tolist() and out["yout"] == rf["yout"].tolist()
    with pytest.raises(ValueError, match="not a valid Ode identifier"):
        demo.solve_lorenz_attractor(json.dumps(dict(cfg, ode="rk99")))
    with pytest.raises(ValueError, match="Can't solve for 2 dimensional type"):
        demo.solve_lorenz_attractor(json.dumps(dict(cfg, init=[1.0, 2.0])))


def test_f32_mode_tracks_f64(deq, oracle):
    """Optional f32 mode (no reference counterpart: the crate cannot instantiate f32): final states track the f64 oracle
    at single-precision tolerances on a non-chaotic horizon; dense rows likewise."""
    from diffeq_b200 import synth
    y0, mu = synth.vdp_sweep(0, 2048, 2048)
    ref = oracle.ensemble_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu, y0, [0.0, 5.0], oracle.opts(1e-5, 1e-6))
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-5), deq.Abstol(1e-6), deq.Precision.F32)
    sol = _problem(deq, deq.System.VanDerPol, mu, y0, [0.0, 5.0]).solve(deq.Ode.Ode45fe, o)
    assert np.all(sol.status == 0)
    np.testing.assert_allclose(sol.final, ref["yfinal"], rtol=2e-3, atol=2e-3)
    assert abs(int(sol.accepted.sum()) - int(ref["accepted"].sum())) < 0.05 * ref["accepted"].sum()
    yl = synth.lorenz_ics(0, 1024)
    rl = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, yl, [0.0, 0.5], oracle.opts(1e-5, 1e-6))
    sl = _problem(deq, deq.System.Lorenz, LOR, yl, [0.0, 0.5]).solve(deq.Ode.Ode45, o)
    np.testing.assert_allclose(sl.final, rl["yfinal"], rtol=5e-3, atol=5e-3)
    with pytest.raises(deq.OdeError):
        _problem(deq, deq.System.Lorenz, LOR, yl, [0.0, 0.5]).solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Precision.F32, deq.Output.All))


def test_nan_and_overflow_paths_match_oracle(deq, oracle):
    """Trial states that overflow to inf/NaN take the reference's `err = 10, dt *= 0.2, timeout = 5` branch
    (problem.rs:819-825) or its natural inf arithmetic; every counter and state must still equal the oracle's
    (NaN-aware comparison)."""
    y0 = np.array([[1e200, -1e200, 1e150], [1e308, 1e308, 1e308], [0.0, 0.0, 0.0], [-0.0, 1e-320, 5e-324],
                   [1e154, 1e154, 1e154], [3.0, 4.0, np.inf], [np.nan, 1.0, 1.0], [1e100, 1.0, -1e100]])
    for ts, kw in (([0.0, 1.0], dict()), ([0.0, 1e-3], dict(initstep=1e-4)), ([0.0, 10.0], dict(maxstep=5.0))):
        ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0, ts, oracle.opts(1e-6, 1e-8, max_steps=400, **kw))
        o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.MaxSteps(400))
        if "initstep" in kw: o.insert(deq.Initstep(kw["initstep"]))
        if "maxstep" in kw: o.insert(deq.Maxstep(kw["maxstep"]))
        sol = _problem(deq, deq.System.Lorenz, LOR, y0, ts).solve(deq.Ode.Ode45, o)
        assert np.array_equal(sol.accepted, ref["accepted"]) and np.array_equal(sol.rejected, ref["rejected"]), (ts, kw)
        assert np.array_equal(sol.status, ref["status"]), (ts, kw)
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True), (ts, kw)
        assert np.array_equal(sol.t_reached, ref["t_reached"], equal_nan=True), (ts, kw)


def test_degenerate_sizes(deq, oracle):
    """Empty and single-element inputs: zero trajectories, one trajectory, one-point tspan."""
    e = _problem(deq, deq.System.Lorenz, LOR, np.zeros((0, 3)), [0.0, 1.0]).solve(deq.Ode.Ode45)
    assert e.final.shape == (0, 3) and e.totals["accepted"] == 0
    e = _problem(deq, deq.System.Lorenz, LOR, np.zeros((0, 3)), [0.0, 1.0]).solve(deq.Ode.Ode4)
    assert e.final.shape == (0, 3)
    # one-point tspan: fixed step takes no step (yout = [y0]); adaptive does one zero-length step (t0 == tend)
    f = _problem(deq, deq.System.Lorenz, LOR, [1.0, 2.0, 3.0], [0.5]).solve(deq.Ode.Ode4, deq.OdeOptionMap().insert(deq.Output.Tspan))
    assert f.final.tolist() == [[1.0, 2.0, 3.0]] and f.yout.shape == (1, 1, 3)
    a = _problem(deq, deq.System.Lorenz, LOR, [1.0, 2.0, 3.0], [0.5]).solve(deq.Ode.Ode45)
    r = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, LOR, [1.0, 2.0, 3.0], [0.5])
    assert int(a.accepted[0]) == r["accepted"] == 1 and np.array_equal(a.final[0], r["yfinal"])


def test_randomised_option_sweep_bit_exact(deq, oracle):
    """Seeded fuzz over methods, tableau sets, systems, tolerances, horizons, directions and step bounds: every case must
    be bit-identical to the oracle (guarded by max_steps so the reference's runaway cases terminate)."""
    from diffeq_b200 import synth
    rng = np.random.default_rng(20261020)
    methods = [("Ode45", 0), ("Ode45", 1), ("Ode45fe", 0), ("Ode23", 1), ("Ode23", 0), ("Ode21", 1), ("Ode21", 0), ("Ode78", 1), ("Ode78", 0)]
    for case in range(24):
        method, tset = methods[case % len(methods)]
        system = ["Lorenz", "VanDerPol", "Pendulum"][int(rng.integers(0, 3))]
        n = int(rng.integers(1, 200))
        if system == "Lorenz":
            y0, params = synth.lorenz_ics(case * 1000, case * 1000 + n), np.array(LOR)
        elif system == "VanDerPol":
            y0, params = synth.vdp_sweep(0, n, max(n, 2))
        else:
            y0, params = synth.pendulum_ics(case * 1000, case * 1000 + n), synth.PENDULUM_PARAMS
        rtol = 10.0 ** rng.uniform(-9, -3)
        atol = 10.0 ** rng.uniform(-10, -4)
        t0 = float(rng.uniform(-1, 1))
        span = float(10.0 ** rng.uniform(-2, 0.7))
        ts = [t0, t0 + span] if rng.random() < 0.8 else [t0 + span, t0]
        kw = dict(max_steps=1500)
        o = deq.OdeOptionMap().insert(deq.Reltol(rtol), deq.Abstol(atol), deq.MaxSteps(1500))
        if rng.random() < 0.3:
            kw["maxstep"] = span / float(rng.integers(3, 30)); o.insert(deq.Maxstep(kw["maxstep"]))
        if rng.random() < 0.3:
            kw["initstep"] = float(np.sign(ts[1] - ts[0]) * span * 10.0 ** rng.uniform(-5, -1)); o.insert(deq.Initstep(kw["initstep"]))
        if rng.random() < 0.2:
            kw["minstep"] = span * 1e-4; o.insert(deq.Minstep(kw["minstep"]))
        ref = oracle.ensemble_adaptive(getattr(oracle, method.upper()), getattr(oracle, system.upper()), params, y0, ts,
                                       oracle.opts(rtol, atol, **kw), tset)
        sol = _problem(deq, getattr(deq.System, system), params, y0, ts).solve(getattr(deq.Ode, method), o, tableau_set=tset)
        tag = (case, method, tset, system, n, rtol, atol, ts, kw)
        assert np.array_equal(sol.accepted, ref["accepted"]) and np.array_equal(sol.rejected, ref["rejected"]), tag
        assert np.array_equal(sol.status, ref["status"]) and np.array_equal(sol.nfe, ref["nfe"]), tag
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True) and np.array_equal(sol.t_reached, ref["t_reached"]), tag


def test_llvm_folded_libm_flavour_bit_exact(deq, oracle):
    """Libm.LlvmFolded (sqrt for the error norm, exp10 in hinit — what an optimised rustc build evaluates) is bit-identical
    to the oracle run in the same flavour, and differs from the as-written flavour on some trajectories."""
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 4096)
    ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0, [0.0, 10.0], oracle.opts(1e-8, 1e-8, libm_folded=1))
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8), deq.Libm.LlvmFolded)
    p = _problem(deq, deq.System.Lorenz, LOR, y0, [0.0, 10.0])
    sol = p.solve(deq.Ode.Ode45, o)
    assert np.array_equal(sol.final, ref["yfinal"]) and np.array_equal(sol.accepted, ref["accepted"])
    assert np.array_equal(sol.rejected, ref["rejected"]) and np.array_equal(sol.t_reached, ref["t_reached"])
    written = p.solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8)))
    frac = float((written.final != sol.final).any(axis=1).mean())
    print(f"trajectories whose final state differs between the two libm flavours: {frac:.3f}")
    assert np.array_equal(written.accepted, sol.accepted)          # same step counts either way on this config
    y0v, mu = synth.vdp_sweep(0, 300, 300)
    ts = deq.linspace(0.0, 20.0, 400)
    rd = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu, y0v, ts, oracle.opts(1e-6, 1e-8, libm_folded=1))
    sd = _problem(deq, deq.System.VanDerPol, mu, y0v, ts).solve(
        deq.Ode.Ode45fe, deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan, deq.Libm.LlvmFolded))
    want = rd["out"].transpose(2, 1, 0)
    m = ~np.isnan(want)
    assert np.array_equal(sd.nvalid, rd["nvalid"]) and np.array_equal(sd.yout[m], want[m])


def test_out_of_memory_fails_cleanly_and_library_stays_usable(deq, oracle):
    """A dense-output request far beyond the 180 GB of HBM returns a CUDA error through the ABI (no crash, no fallback) and
    the next solve still works and is still bit-exact."""
    from diffeq_b200 import synth
    y0, mu = synth.vdp_sweep(0, 1 << 16, 1 << 16)
    ts = deq.linspace(0.0, 20.0, 400000)            # 2^16 x 4e5 x 2 x 8 B = 419 GB
    p = _problem(deq, deq.System.VanDerPol, mu, y0, ts)
    with pytest.raises(deq.OdeError) as e:
        p.solve(deq.Ode.Ode45fe, deq.OdeOptionMap().insert(deq.Output.Tspan))
    assert e.value.kind == "Cuda"
    y1 = synth.lorenz_ics(0, 300)
    sol = _problem(deq, deq.System.Lorenz, LOR, y1, [0.0, 1.0]).solve(deq.Ode.Ode45)
    ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y1, [0.0, 1.0])
    assert np.array_equal(sol.final, ref["yfinal"])


@pytest.mark.parametrize("order", ["sorted", "permuted"])
def test_trajectory_major_staged_dense_output(deq, oracle, order):
    """DenseLayout.TrajMajor (per-lane shared-memory staging, cooperative 128-byte stores) returns exactly the rows of the
    SoA path / the oracle, for a sorted and for a randomly permuted parameter sweep, ragged ensemble size, and a tspan so
    dense that one step emits more rows than the staging ring holds."""
    from diffeq_b200 import synth
    n = 1000 + 37
    y0, mu = synth.vdp_sweep(0, n, n)
    if order == "permuted":
        mu = mu[np.random.default_rng(3).permutation(n)].copy()
    for npts in (1000, 23, 9001):
        ts = deq.linspace(0.0, 20.0, npts)
        ref = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu, y0, ts, oracle.opts(1e-6, 1e-8))
        o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan, deq.DenseLayout.TrajMajor)
        sol = _problem(deq, deq.System.VanDerPol, mu, y0, ts).solve(deq.Ode.Ode45fe, o)
        want = ref["out"].transpose(2, 1, 0)
        m = ~np.isnan(want)
        assert np.array_equal(sol.nvalid, ref["nvalid"]), npts
        assert np.array_equal(np.isnan(sol.yout), np.isnan(want)), npts
        assert np.array_equal(sol.yout[m], want[m]), npts
    # Lorenz (n = 3: rows straddle the 16-double lines) in FAST mode agrees with its own SoA output bit for bit
    yl = synth.lorenz_ics(0, 333)
    tl = deq.linspace(0.0, 2.0, 301)
    a = _problem(deq, deq.System.Lorenz, LOR, yl, tl).solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Output.Tspan, deq.Mode.Fast))
    b = _problem(deq, deq.System.Lorenz, LOR, yl, tl).solve(
        deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Output.Tspan, deq.Mode.Fast, deq.DenseLayout.TrajMajor))
    assert np.array_equal(a.yout, b.yout, equal_nan=True)


def test_nvrtc_max_dimension_high_order(deq):
    """n = 16 (the largest state the register-resident steppers accept) with the 13-stage Fehlberg 7(8): 8 decoupled
    oscillators with frequencies 1..8 through NVRTC, against the exact solution; also the fixed-step RK4 path."""
    src = ("__device__ void rhs(double t, const double* y, double* dy, const double* p) {\n"
           "  for (int i = 0; i < 8; ++i) { double w = p[0] * (i + 1); dy[2 * i] = w * y[2 * i + 1]; dy[2 * i + 1] = -w * y[2 * i]; }\n}\n")
    rhs = deq.Rhs.from_source(src, 16, 1)
    y0 = np.tile(np.array([1.0, 0.0] * 8), (70, 1))
    T = 1.5
    want = np.array([f(k * T) for k in range(1, 9) for f in (np.cos, lambda x: -np.sin(x))])
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-11), deq.Abstol(1e-11))
    sol = deq.OdeProblem.builder().fun(rhs).params([1.0]).init(y0).tspan([0.0, T]).build().solve(
        deq.Ode.Ode78, o, tableau_set=deq.TableauSet.Textbook)
    assert np.all(sol.status == 0)
    np.testing.assert_allclose(sol.final, np.tile(want, (70, 1)), rtol=0, atol=2e-9)
    fx = deq.OdeProblem.builder().fun(rhs).params([1.0]).init(y0[:3]).tspan(deq.linspace(0.0, T, 3001)).build().solve(deq.Ode.Ode4)
    np.testing.assert_allclose(fx.final, np.tile(want, (3, 1)), rtol=0, atol=1e-9)
    with pytest.raises(deq.OdeError):
        deq.Rhs.from_source(src, 65, 1)       # beyond the supported state dimension


def test_in_library_multi_gpu_solve_host(deq, oracle):
    """deq_solve_host shards an ensemble over the GPUs of the box with one host thread per device; trajectory i's result
    does not depend on the device count (runs with however many GPUs are visible; 1 is a valid count)."""
    import torch
    from diffeq_b200 import synth
    ndev = torch.cuda.device_count()
    if ndev < 2:
        print("NOTE: only one GPU visible; devices = list(range(ndev)) degenerates to [0] (see test_multi_gpu_devices_list)")
    y0 = synth.lorenz_ics(0, 5003)
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8))
    ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0, [0.0, 2.0], oracle.opts(1e-8, 1e-8))
    rhs = deq.Rhs.builtin(deq.System.Lorenz)
    for devices in ([0], list(range(ndev)), [0] * 3):   # the last one: three shards on one GPU (partition invariance)
        out = deq.solve_host_multi(rhs, y0, LOR, [0.0, 2.0], deq.Ode.Ode45, o, devices=devices)
        assert np.array_equal(out["final"], ref["yfinal"]), devices
        assert np.array_equal(out["accepted"], ref["accepted"]) and np.array_equal(out["t_reached"], ref["t_reached"]), devices
    fx = deq.solve_host_multi(rhs, y0[:100], LOR, deq.linspace(0.0, 1.0, 101), deq.Ode.Ode4, devices=list(range(ndev)))
    assert np.array_equal(fx["final"], oracle.ensemble_fixed(oracle.ODE4, oracle.LORENZ, LOR, y0[:100], oracle.linspace(0.0, 1.0, 101)))
    with pytest.raises(deq.OdeError):
        deq.solve_host_multi(rhs, y0, LOR, [0.0, 2.0], deq.Ode.Ode45, o, devices=[ndev + 7])
    # block-cyclic sharding (deq_options.shard_chunk) for parameter-sorted sweeps: same results, trajectory by trajectory
    from diffeq_b200 import synth as _s
    yv, mu = _s.vdp_sweep(0, 1037, 1037)
    rv = oracle.ensemble_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu, yv, [0.0, 10.0], oracle.opts(1e-6, 1e-8))
    vd = deq.Rhs.builtin(deq.System.VanDerPol)
    for chunk in (0, 1, 64, 5000):
        oc = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.ShardChunk(chunk))
        out = deq.solve_host_multi(vd, yv, mu, [0.0, 10.0], deq.Ode.Ode45fe, oc, devices=[0] * 8 if ndev == 1 else list(range(ndev)))   # chunk 5000: seven empty shards
        assert np.array_equal(out["final"], rv["yfinal"]) and np.array_equal(out["accepted"], rv["accepted"]), chunk
        assert np.array_equal(out["nfe"], rv["nfe"]) and np.array_equal(out["t_reached"], rv["t_reached"]), chunk


def test_solve_host_dense_rows_across_shards(deq, oracle, monkeypatch):
    """deq_solve_host_dense: the tspan rows of an ensemble on host buffers, integrated chunk by chunk (several chunks per
    device, several device threads — all visible GPUs, and three threads on GPU 0 for the partition invariance) while earlier
    chunks travel to the host.  Slices of the result are compared with the oracle's dense output, the whole of it with the
    handle API, in both layouts; a multi-GPU box exercises devices = 0..ndev-1."""
    import torch
    from diffeq_b200 import synth
    ndev = torch.cuda.device_count()
    if ndev < 2:
        print("NOTE: only one GPU visible; the multi-device path is exercised with several device threads on GPU 0")
    n = 5003
    y0, mu = synth.vdp_sweep(0, n, n)
    ts = deq.linspace(0.0, 10.0, 101)
    monkeypatch.setenv("DEQ_HOST_CHUNK", "1024")   # five chunks
    vd = deq.Rhs.builtin(deq.System.VanDerPol)
    slices = [slice(0, 40), slice(2030, 2070), slice(n - 40, n)]
    refs = [oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu[sl], y0[sl], ts, oracle.opts(1e-6, 1e-8)) for sl in slices]
    for layout in (deq.DenseLayout.SoA, deq.DenseLayout.TrajMajor):
        o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan, layout)
        whole = deq.OdeProblem.builder().fun(vd).params(mu).init(y0).tspan(ts).build().solve(deq.Ode.Ode45fe, o)
        for devices in ([0], list(range(ndev)), [0] * 3):
            out = deq.solve_host_multi(vd, y0, mu, ts, deq.Ode.Ode45fe, o, devices=devices)
            dense = out["dense"] if layout == deq.DenseLayout.TrajMajor else np.ascontiguousarray(out["dense"].transpose(2, 1, 0))   # [traj][row][comp]
            assert np.array_equal(out["nvalid"], whole.nvalid) and np.array_equal(out["final"], whole.final), (layout, devices)
            assert np.array_equal(out["accepted"], whole.accepted) and np.array_equal(out["t_reached"], whole.t_reached)
            for sl, ref in zip(slices, refs):
                want = ref["out"].transpose(2, 1, 0)   # oracle: SoA [n][nt][traj], NaN beyond nvalid (except the last row)
                got = dense[sl].copy()
                idx = np.arange(len(ts))[None, :]
                invalid = (idx >= ref["nvalid"][:, None]) & (idx != len(ts) - 1)
                got[invalid] = np.nan
                assert np.array_equal(got, want, equal_nan=True), (layout, devices, sl)
                assert np.array_equal(out["nvalid"][sl], ref["nvalid"])
    # fixed-step rows and the stiff solver go through the same call (SoA rows)
    ys = synth.lorenz_ics(0, 2500)
    of = deq.OdeOptionMap().insert(deq.Output.Tspan)
    fx = deq.solve_host_multi(deq.Rhs.builtin(deq.System.Lorenz), ys, LOR, deq.linspace(0.0, 0.5, 51), deq.Ode.Ode4, of, devices=[0, 0])
    wf = deq.OdeProblem.builder().fun(deq.Rhs.builtin(deq.System.Lorenz)).params(LOR).init(ys).tspan(deq.linspace(0.0, 0.5, 51)).build().solve(deq.Ode.Ode4, of)
    assert np.array_equal(np.ascontiguousarray(fx["dense"].transpose(2, 1, 0)), wf.yout) and np.array_equal(fx["final"], wf.final)
    with pytest.raises(deq.OdeError):   # Output.All has a data-dependent size: handle API only
        deq.solve_host_multi(vd, y0, mu, ts, deq.Ode.Ode45fe, deq.OdeOptionMap().insert(deq.Output.All))


def test_multi_gpu_devices_list(deq, oracle):
    """deq_solve_host(devices = 0..ndev-1) on a box with at least two GPUs: one process, one host thread and stream per device,
    results identical to the oracle trajectory by trajectory.  Skipped (and recorded as such) on a single-GPU box."""
    import torch
    from diffeq_b200 import synth
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip(f"needs >= 2 GPUs, this box has {ndev}")
    y0 = synth.lorenz_ics(0, 20011)
    ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0, [0.0, 2.0], oracle.opts(1e-8, 1e-8))
    out = deq.solve_host_multi(deq.Rhs.builtin(deq.System.Lorenz), y0, LOR, [0.0, 2.0], deq.Ode.Ode45,
                               deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8)), devices=list(range(ndev)))
    assert np.array_equal(out["final"], ref["yfinal"]) and np.array_equal(out["accepted"], ref["accepted"])
    assert np.array_equal(out["rejected"], ref["rejected"]) and np.array_equal(out["status"], ref["status"])


def test_golden_fixtures_through_the_cuda_path(deq):
    """Every committed golden case (tests/golden/solver_cases.json: known answers of the independent Python restatement,
    reference tableaus parsed from runge_kutta.rs) reproduced by the CUDA path directly — no oracle in between."""
    import hashlib, json, os
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solver_cases.json")))
    fh = lambda v: [fh(x) for x in v] if isinstance(v, list) else float.fromhex(v)  # noqa: E731
    ran = 0
    for c in cases:
        ts = [float(x) for x in c["tspan"][1:]] if c["tspan"][0] == "list" else deq.linspace(c["tspan"][1], c["tspan"][2], c["tspan"][3])
        prob = _problem(deq, getattr(deq.System, c["system"]), fh(c["params"]), fh(c["y0"]), ts)
        if c["kind"] == "fixed":
            sol = prob.solve(getattr(deq.Ode, c["method"]), deq.OdeOptionMap().insert(deq.Output.Tspan))
            assert sol.final[0].tolist() == fh(c["final"]), c["name"]
            assert hashlib.sha256(np.ascontiguousarray(sol.yout[0], dtype="<f8").tobytes()).hexdigest() == c["sha256_yout"], c["name"]
        else:
            opt = c["options"]
            o = deq.OdeOptionMap().insert(deq.Output.All)
            if opt.get("points_all") is False:
                o.insert(deq.Points.Specified)   # reproduced with the reference's defect (problem.rs:262)
            for k, cls in (("reltol", deq.Reltol), ("abstol", deq.Abstol), ("minstep", deq.Minstep), ("maxstep", deq.Maxstep), ("initstep", deq.Initstep)):
                if k in opt:
                    o.insert(cls(float.fromhex(opt[k])))
            if "max_steps" in opt:
                o.insert(deq.MaxSteps(opt["max_steps"]))
            if opt.get("libm_folded"):
                o.insert(deq.Libm.LlvmFolded)
            sol = prob.solve(getattr(deq.Ode, c["method"]), o, tableau_set=0 if c["tableau_set"] == "reference" else 1)
            tout, yout = sol.trajectory(0)
            assert (int(sol.accepted[0]), int(sol.rejected[0]), int(sol.status[0]), len(tout)) == (c["accepted"], c["rejected"], c["status"], c["nout"]), c["name"]
            assert sol.final[0].tolist() == fh(c["final"]), c["name"]
            assert hashlib.sha256(np.ascontiguousarray(tout, dtype="<f8").tobytes()).hexdigest() == c["sha256_tout"], c["name"]
            assert hashlib.sha256(np.ascontiguousarray(yout, dtype="<f8").tobytes()).hexdigest() == c["sha256_yout"], c["name"]
        ran += 1
    assert ran == len(cases) >= 34


def test_points_specified_reproduces_the_reference_defect(deq, oracle):
    """Points::Specified (problem.rs:284-300): y is re-read from the OUTPUT vector (:262), so the solver's notion of the state
    lags behind whenever the tspan grid is coarser than the steps.  The product reproduces that stream bit for bit."""
    from diffeq_b200 import synth
    n = 300
    y0, mu = synth.vdp_sweep(0, n, n)
    ts = deq.linspace(0.0, 10.0, 41)
    base = (deq.Reltol(1e-6), deq.Abstol(1e-8), deq.MaxSteps(20000), deq.Points.Specified)
    prob = _problem(deq, deq.System.VanDerPol, mu, y0, ts)
    sa = prob.solve(deq.Ode.Ode45fe, deq.OdeOptionMap().insert(*base, deq.Output.All))
    sf = prob.solve(deq.Ode.Ode45fe, deq.OdeOptionMap().insert(*base))                       # final state only: counting pass
    oo = oracle.opts(1e-6, 1e-8, points=oracle.POINTS_SPECIFIED, max_steps=20000)
    ens = oracle.ensemble_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu, y0, ts, oo)
    for s in (sa, sf):
        assert np.array_equal(s.final, ens["yfinal"]) and np.array_equal(s.accepted, ens["accepted"])
        assert np.array_equal(s.rejected, ens["rejected"]) and np.array_equal(s.status, ens["status"])
        assert np.array_equal(s.nfe, ens["nfe"]) and np.array_equal(s.t_reached, ens["t_reached"])
    for i in (0, 1, 150, n - 1):
        one = oracle.solve_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu[i], y0[i], ts, oo)
        tout, yout = sa.trajectory(i)
        assert np.array_equal(tout, one["tout"]) and np.array_equal(yout, one["yout"]), i
        assert np.array_equal(tout, ts[:len(tout)])                                           # rows are the tspan points
    # user systems (NVRTC) and a 3-dimensional FSAL method
    yl = synth.lorenz_ics(0, 64)
    tl = deq.linspace(0.0, 1.0, 201)
    ol = oracle.opts(1e-7, 1e-7, points=oracle.POINTS_SPECIFIED, max_steps=5000)
    el = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, yl, tl, ol)
    rhs = deq.Rhs.from_source(USER_LORENZ, 3, 3)
    su = deq.OdeProblem.builder().fun(rhs).params(LOR).init(yl).tspan(tl).build().solve(
        deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-7), deq.Abstol(1e-7), deq.MaxSteps(5000), deq.Points.Specified, deq.Output.All))
    assert np.array_equal(su.final, el["yfinal"]) and np.array_equal(su.accepted, el["accepted"])
    with pytest.raises(deq.OdeError) as e:
        prob.solve(deq.Ode.Ode45fe, deq.OdeOptionMap().insert(deq.Points.Specified, deq.Output.Tspan))
    assert e.value.kind == "Unsupported"
    with pytest.raises(deq.OdeError):
        prob.solve(deq.Ode.Ode45fe, deq.OdeOptionMap().insert(deq.Points.Specified, deq.Mode.Fast))


def test_full_size_config2_properties(deq, oracle):
    """BASELINE config 2 at its full size (2^20 trajectories, T = 10), checked through size-independent properties: the
    result of trajectory i is independent of the partition (one solve vs four shards) and of the scheduling (lane refill vs
    static warps), the device-side totals equal the sums of the per-trajectory counters, every trajectory finishes, and a
    random sample of 1500 trajectories equals the oracle bit for bit."""
    from diffeq_b200 import synth
    n = 1 << 20
    y0 = synth.lorenz_ics(0, n)
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8))
    full = _problem(deq, deq.System.Lorenz, LOR, y0, [0.0, 10.0]).solve(deq.Ode.Ode45, o)
    assert np.all(full.status == 0) and np.all(full.t_reached == 10.0)
    assert full.totals["accepted"] == int(full.accepted.sum(dtype=np.int64)) and full.totals["rejected"] == int(full.rejected.sum(dtype=np.int64))
    assert full.totals["nfe"] == int(full.nfe.sum(dtype=np.int64)) == 2 * n + 6 * (full.totals["accepted"] + full.totals["rejected"])
    for b, e in ((0, 300001), (300001, 524288), (524288, 1000000), (1000000, n)):
        part = _problem(deq, deq.System.Lorenz, LOR, y0[b:e], [0.0, 10.0]).solve(deq.Ode.Ode45, o)
        assert np.array_equal(part.final, full.final[b:e]) and np.array_equal(part.accepted, full.accepted[b:e])
    static = _problem(deq, deq.System.Lorenz, LOR, y0, [0.0, 10.0]).solve(deq.Ode.Ode45, deq.OdeOptionMap().insert(deq.Reltol(1e-8), deq.Abstol(1e-8), deq.Refill(0)))
    assert np.array_equal(static.final, full.final) and np.array_equal(static.rejected, full.rejected)
    idx = np.sort(np.random.default_rng(11).choice(n, 1500, replace=False))
    ref = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0[idx], [0.0, 10.0], oracle.opts(1e-8, 1e-8))
    assert np.array_equal(full.final[idx], ref["yfinal"]) and np.array_equal(full.accepted[idx], ref["accepted"])
    assert np.array_equal(full.rejected[idx], ref["rejected"])


def test_full_size_config3_dense_slices(deq, oracle):
    """BASELINE config 3 at 2^20 trajectories x 1000 rows (16.8 GB on the device, both layouts): three slices of the dense
    output are copied back and must equal the oracle's rows bit for bit; nvalid and the step counters likewise."""
    import ctypes as C
    from diffeq_b200 import synth
    n = 1 << 20
    y0, mu = synth.vdp_sweep(0, n, n)
    ts = deq.linspace(0.0, 20.0, 1000)
    L = deq.lib()
    prob = _problem(deq, deq.System.VanDerPol, mu, y0, ts)
    for layout in (deq.DenseLayout.SoA, deq.DenseLayout.TrajMajor):
        o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8), deq.Output.Tspan, layout).to_c()
        res = C.c_void_p()
        assert L.deq_solve(prob._h, int(deq.Ode.Ode45fe), 0, C.byref(o), C.byref(res)) == 0
        try:
            for b in (0, 524288 - 17, n - 64):
                e = b + 64
                ref = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu[b:e], y0[b:e], ts, oracle.opts(1e-6, 1e-8))
                shape = (64, 1000, 2) if layout == deq.DenseLayout.TrajMajor else (2, 1000, 64)
                got = np.empty(shape); nv = np.empty(64, np.uint32)
                assert L.deq_result_dense(res, C.c_size_t(b), C.c_size_t(e), got.ctypes.data_as(C.POINTER(C.c_double)),
                                          nv.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
                if layout == deq.DenseLayout.SoA:
                    got = got.transpose(2, 1, 0)
                want = ref["out"].transpose(2, 1, 0)
                m = ~np.isnan(want)
                assert np.array_equal(nv, ref["nvalid"]) and np.array_equal(got[m], want[m]), (layout, b)
        finally:
            L.deq_result_destroy(res)


def test_concurrent_handles_from_several_host_threads(deq, oracle):
    """Distinct handles may be used concurrently from different threads (include/diffeq_b200.h threading contract): four
    host threads solve four different ensembles at the same time; every result equals the oracle's."""
    import threading
    from diffeq_b200 import synth
    jobs = [("Ode45", "Lorenz", synth.lorenz_ics(0, 3000), np.array(LOR), [0.0, 3.0], 1e-8),
            ("Ode45fe", "VanDerPol", *synth.vdp_sweep(0, 2000, 2000), [0.0, 10.0], 1e-6),
            ("Ode45", "Lorenz", synth.lorenz_ics(5000, 7000), np.array(LOR), [0.0, 2.0], 1e-6),
            ("Ode78", "Pendulum", synth.pendulum_ics(0, 1500), synth.PENDULUM_PARAMS, [0.0, 5.0], 1e-10)]
    out, errs = [None] * len(jobs), []

    def run(i):
        try:
            m, s, y0, p, ts, tol = jobs[i]
            for _ in range(3):
                out[i] = _problem(deq, getattr(deq.System, s), p, y0, ts).solve(
                    getattr(deq.Ode, m), deq.OdeOptionMap().insert(deq.Reltol(tol), deq.Abstol(tol)), tableau_set=1)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i, (m, s, y0, p, ts, tol) in enumerate(jobs):
        ref = oracle.ensemble_adaptive(getattr(oracle, m.upper()), getattr(oracle, s.upper()), p, y0, ts, oracle.opts(tol, tol), 1)
        assert np.array_equal(out[i].final, ref["yfinal"]) and np.array_equal(out[i].accepted, ref["accepted"]), i


def test_device_resident_inputs_and_outputs(deq, oracle):
    """deq_ensemble_create_device / deq_result_device_ptrs: inputs and results stay on the GPU in the library's SoA layout
    (here owned by torch tensors); values equal the oracle's."""
    import ctypes as C
    import torch
    from diffeq_b200 import synth
    n = 3001
    y0v, mu = synth.vdp_sweep(0, n, n)
    d_y0 = torch.from_numpy(np.ascontiguousarray(y0v.T)).cuda()          # [2][n]
    d_mu = torch.from_numpy(np.ascontiguousarray(mu.T)).cuda()           # [1][n]
    torch.cuda.synchronize()
    rhs = deq.Rhs.builtin(deq.System.VanDerPol)
    prob = deq.OdeProblem.from_device(rhs, d_y0.data_ptr(), n, [0.0, 20.0], d_params_soa_ptr=d_mu.data_ptr())
    ref = oracle.ensemble_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu, y0v, [0.0, 20.0], oracle.opts(1e-6, 1e-8))
    sol = prob.solve(deq.Ode.Ode45fe, deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8)))
    assert np.array_equal(sol.final, ref["yfinal"]) and np.array_equal(sol.accepted, ref["accepted"])
    # raw device view of the result
    L = deq.lib()
    o = deq.OdeOptionMap().insert(deq.Reltol(1e-6), deq.Abstol(1e-8)).to_c()
    res = C.c_void_p(); yf = C.POINTER(C.c_double)(); dn = C.POINTER(C.c_double)()
    assert L.deq_solve(prob._h, int(deq.Ode.Ode45fe), 0, C.byref(o), C.byref(res)) == 0
    try:
        assert L.deq_result_device_ptrs(res, C.byref(yf), C.byref(dn)) == 0 and not dn
        addr = C.cast(yf, C.c_void_p).value
        rt = C.CDLL("libcudart.so.12")                      # any CUDA runtime can read the library's device buffer
        host = np.empty((2, n))
        assert rt.cudaMemcpy(host.ctypes.data_as(C.c_void_p), C.c_void_p(addr), C.c_size_t(2 * n * 8), 2) == 0
        assert np.array_equal(host.T, ref["yfinal"])
    finally:
        L.deq_result_destroy(res)
    # shared parameters variant
    y0l = synth.lorenz_ics(0, 500)
    d_l = torch.from_numpy(np.ascontiguousarray(y0l.T)).cuda()
    pl = deq.OdeProblem.from_device(deq.Rhs.builtin(deq.System.Lorenz), d_l.data_ptr(), 500, [0.0, 1.0], params=LOR)
    rl = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, LOR, y0l, [0.0, 1.0])
    assert np.array_equal(pl.solve(deq.Ode.Ode45).final, rl["yfinal"])


RING_SRC = r'''
// ring of DEQ_USER_N logistic cells with diffusive coupling (the oracle's RingF): same expression, same order
__device__ void rhs(double t, const double* v, double* d, const double* p) {
#pragma unroll
    for (int i = 0; i < DEQ_USER_N; ++i) {
        const int ip = (i + 1) % DEQ_USER_N, im = (i + DEQ_USER_N - 1) % DEQ_USER_N;
        d[i] = p[0] * (v[ip] - v[i]) + p[1] * (v[i] * (1.0 - v[im]));
    }
}
'''


@pytest.mark.parametrize("n", [8, 16, 32, 64])
def test_generic_dimension_user_systems_bit_exact(deq, oracle, n):
    """State dimension 8 and 16 (the `OdeType` impls of the reference reach 9-tuples and arbitrary Vecs, types.rs:138-278): the
    generic-dimension NVRTC path against the oracle's ring system — adaptive (FSAL and 13-stage non-FSAL), fixed step, final
    states, dense rows in both layouts and the ragged stream, all bit for bit."""
    osys = {8: oracle.RING8, 16: oracle.RING16, 32: oracle.RING32, 64: oracle.RING64}[n]   # beyond 16 the stage vectors live in local memory
    rng = np.random.default_rng(n)
    m = 300
    y0 = rng.uniform(0.1, 0.9, (m, n))
    prm = np.column_stack([rng.uniform(0.1, 0.5, m), rng.uniform(0.5, 2.5, m)])
    rhs = deq.Rhs.from_source(RING_SRC, n, 2)
    assert rhs.dim == n
    ts = deq.linspace(0.0, 4.0, 33)
    prob = deq.OdeProblem.builder().fun(rhs).params(prm).init(y0).tspan(ts).build()
    for method, tset, tol in ((deq.Ode.Ode45, 0, 1e-8), (deq.Ode.Ode78, 1, 1e-10), (deq.Ode.Ode45fe, 0, 1e-6)):
        om = {deq.Ode.Ode45: oracle.ODE45, deq.Ode.Ode78: oracle.ODE78, deq.Ode.Ode45fe: oracle.ODE45FE}[method]
        oo = oracle.opts(tol, tol)
        base = (deq.Reltol(tol), deq.Abstol(tol))
        ref = oracle.ensemble_adaptive(om, osys, prm, y0, ts, oo, tset)
        sol = prob.solve(method, deq.OdeOptionMap().insert(*base), tableau_set=tset)
        for k in ("accepted", "rejected", "nfe", "status", "t_reached"):
            assert np.array_equal(getattr(sol, k), ref[k]), (method, k)
        assert np.array_equal(sol.final, ref["yfinal"]), method
        rd = oracle.ensemble_dense(om, osys, prm, y0, ts, oo, tset)
        want = rd["out"].transpose(2, 1, 0)
        for lay in (deq.DenseLayout.SoA, deq.DenseLayout.TrajMajor):
            sd = prob.solve(method, deq.OdeOptionMap().insert(*base, deq.Output.Tspan, lay), tableau_set=tset)
            assert np.array_equal(sd.nvalid, rd["nvalid"]) and np.array_equal(sd.yout, want), (method, lay)
        sa = prob.solve(method, deq.OdeOptionMap().insert(*base, deq.Output.All), tableau_set=tset)
        for i in (0, m - 1):
            one = oracle.solve_adaptive(om, osys, prm[i], y0[i], ts, oo, tset=tset)
            tout, yout = sa.trajectory(i)
            assert np.array_equal(tout, one["tout"]) and np.array_equal(yout, one["yout"]), (method, i)
    fx = prob.solve(deq.Ode.Ode4, deq.OdeOptionMap().insert(deq.Output.Tspan))      # RK4 at dt = 1/8 blows up for the stiffer cells:
    rfx = oracle.ensemble_fixed(oracle.ODE4, osys, prm, y0, ts)                     # the overflow to NaN must match as well
    assert np.array_equal(fx.final, rfx, equal_nan=True) and np.isfinite(rfx).all(axis=1).mean() > 0.5
    for i in (3, 7):
        assert np.array_equal(fx.yout[i], oracle.solve_fixed(oracle.ODE4, osys, prm[i], y0[i], ts)["yout"], equal_nan=True)
    if n > 32:
        with pytest.raises(deq.OdeError) as e:       # the stiff path carries n <= 32
            prob.solve(deq.Ode.Ode23s)
        assert e.value.kind == "Unsupported"
