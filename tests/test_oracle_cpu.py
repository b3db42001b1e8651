"""CPU: the C++ oracle against the committed golden fixtures (tests/golden/, made by tools/gen_golden.py from the
reference's tableau source and the independent Python restatement tests/pyref.py), bit-for-bit."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
TABS = json.load(open(os.path.join(GOLD, "tableaus.json")))
CASES = json.load(open(os.path.join(GOLD, "solver_cases.json")))
NAME2TAB = {"Feuler": "feuler", "Heun": "heun", "Midpoint": "midpoint", "Ode4": "rk4", "Ode21": "rk21", "Ode23": "rk23",
            "Ode45": "dopri5", "Ode45fe": "rk45", "Ode78": "feh78"}


def fh(v):
    return [fh(x) for x in v] if isinstance(v, list) else float.fromhex(v)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<f8").tobytes()).hexdigest()


def tspan_of(desc):
    if desc[0] == "list":
        return [float(x) for x in desc[1:]]
    a, b, n = desc[1], desc[2], desc[3]
    step = (b - a) / (n - 1)
    return [a + step * i for i in range(n)]


@pytest.mark.parametrize("tset", ["reference", "textbook"])
@pytest.mark.parametrize("method", sorted(NAME2TAB))
def test_oracle_tableaus_match_reference_source(oracle, method, tset):
    """Coefficients parsed from runge_kutta.rs (row-major constructor semantics) == the oracle's transcription."""
    g = TABS[tset][NAME2TAB[method]]
    t = oracle.tableau(getattr(oracle, method.upper()), 0 if tset == "reference" else 1)
    assert t["S"] == g["S"] and t["adaptive"] == g["adaptive"] and t["order_min"] == g["order_min"] and t["fsal"] == g["fsal"]
    assert np.array_equal(t["a"], np.array(fh(g["a"])))
    assert np.array_equal(t["c"], np.array(fh(g["c"])))
    gb = np.array(fh(g["b"]))
    assert np.array_equal(t["b"] if g["adaptive"] else t["b"][:, 0], gb if g["adaptive"] else gb[:, 0])


def test_reference_defects_are_preserved():
    """the reference set keeps the mis-transcriptions, the textbook set fixes exactly those."""
    r, t = TABS["reference"], TABS["textbook"]
    assert fh(r["rk21"]["b"]) == [[0.5, 0.5], [1.0, 0.0]] and fh(t["rk21"]["b"]) == [[0.5, 1.0], [0.5, 0.0]]
    assert fh(r["rk23"]["b"])[0][1] == 1.0 / 9.0 and fh(t["rk23"]["b"])[0][1] == 2.0 / 9.0
    assert fh(r["dopri5"]["c"])[3] == 0.75 and fh(t["dopri5"]["c"])[3] == 0.8
    assert fh(r["feh78"]["a"])[5][:5] == [0.5, 0.0, 0.0, 0.75, 0.2]
    assert abs(sum(b[0] for b in fh(r["feh78"]["b"])) - 267.0 / 280.0) < 1e-15
    assert abs(sum(b[0] for b in fh(t["feh78"]["b"])) - 1.0) < 1e-15
    assert r["dopri5"]["fsal"] and not r["rk45"]["fsal"] and not r["feh78"]["fsal"]     # runge_kutta.rs:832-836


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "fixed"], ids=lambda c: c["name"])
def test_oracle_fixed_goldens(oracle, case):
    ts = tspan_of(case["tspan"])
    r = oracle.solve_fixed(getattr(oracle, case["method"].upper()), getattr(oracle, case["system"].upper()),
                           fh(case["params"]), fh(case["y0"]), ts)
    assert r["yfinal"].tolist() == fh(case["final"])
    for i, row in case["rows"].items():
        assert r["yout"][int(i)].tolist() == fh(row)
    assert digest(r["yout"]) == case["sha256_yout"]


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "adaptive"], ids=lambda c: c["name"])
def test_oracle_adaptive_goldens(oracle, case):
    ts = tspan_of(case["tspan"])
    kw = {}
    for k, v in case["options"].items():
        if k == "points_all":
            kw["points"] = 0 if v else 1
        elif k == "libm_folded":
            kw["libm_folded"] = int(v)
        else:
            kw[k] = float.fromhex(v) if isinstance(v, str) else v
    r = oracle.solve_adaptive(getattr(oracle, case["method"].upper()), getattr(oracle, case["system"].upper()),
                              fh(case["params"]), fh(case["y0"]), ts, oracle.opts(**kw),
                              tset=0 if case["tableau_set"] == "reference" else 1)
    assert r["rc"] == 0
    assert (r["accepted"], r["rejected"], r["status"], r["nout"]) == (case["accepted"], case["rejected"], case["status"], case["nout"])
    assert r["yfinal"].tolist() == fh(case["final"])
    for i, (t, y) in case["rows"].items():
        assert r["tout"][int(i)] == float.fromhex(t) and r["yout"][int(i)].tolist() == fh(y)
    assert digest(r["tout"]) == case["sha256_tout"] and digest(r["yout"]) == case["sha256_yout"]


def test_cfg1_known_answer(oracle):
    """BASELINE config 1 (the reference's own Lorenz fixture, problem.rs:978-1010) — survey-time restatement value."""
    c = [c for c in CASES if c["name"] == "cfg1_rk4_lorenz_t100"][0]
    assert c["final"] == ["0x1.12621c0124822p+1", "-0x1.2b36df19c01edp+2", "0x1.dd8c99d38f21ep+4"]


def test_oracle_error_paths(oracle):
    r = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, [10, 28, 8 / 3], [1, 2, 20], [0.0, 1.0], oracle.opts(initstep=-0.1))
    assert r["rc"] == 5   # InvalidInitstep (problem.rs:232-237)
    r = oracle.solve_adaptive(oracle.ODE4, oracle.LORENZ, [10, 28, 8 / 3], [1, 2, 20], [0.0, 1.0])
    assert r["rc"] == 2   # InvalidButcherTableauWeightType (problem.rs:204-209)
    r = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, [10, 28, 8 / 3], [1, 2, 20], [])
    assert r["rc"] == 0 and r["nout"] == 0   # empty tspan: nothing to solve (problem.rs:211-214)


def test_oracle_ensemble_matches_single(oracle):
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 64)
    e = oracle.ensemble_adaptive(oracle.ODE45, oracle.LORENZ, synth.LORENZ_PARAMS, y0, [0.0, 2.0], oracle.opts(1e-8, 1e-8), nthreads=2)
    for i in (0, 17, 63):
        s = oracle.solve_adaptive(oracle.ODE45, oracle.LORENZ, synth.LORENZ_PARAMS, y0[i], [0.0, 2.0], oracle.opts(1e-8, 1e-8))
        assert s["accepted"] == e["accepted"][i] and np.array_equal(s["yfinal"], e["yfinal"][i])
    # dense (tspan-filtered Points::All) == the interpolated rows of the ragged single-trajectory output
    y0v, mu = synth.vdp_sweep(0, 8, 8)
    ts = oracle.linspace(0.0, 20.0, 200)
    d = oracle.ensemble_dense(oracle.ODE45FE, oracle.VANDERPOL, mu, y0v, ts, oracle.opts(1e-6, 1e-8))
    s = oracle.solve_adaptive(oracle.ODE45FE, oracle.VANDERPOL, mu[3], y0v[3], ts, oracle.opts(1e-6, 1e-8))
    lut = {t: y for t, y in zip(s["tout"], s["yout"])}
    for k in range(1, int(d["nvalid"][3])):
        assert np.array_equal(d["out"][:, k, 3], lut[ts[k]])


# ---------------------------------------------------------------------------------------------------------------------
# Stiff path: ode23s / oderosenbrock.  Goldens come from tests/pyref.py (tools/gen_golden.py stiff), the
# Rosenbrock coefficient sets are parsed from the reference's rosenbrock.rs.
# ---------------------------------------------------------------------------------------------------------------------
STIFF = json.load(open(os.path.join(GOLD, "stiff_cases.json")))


def _stiff_opts(oracle, case):
    kw = {}
    for k, v in case["options"].items():
        if k == "points_all":
            kw["points"] = 0 if v else 1
        elif k == "libm_folded":
            kw["libm_folded"] = int(v)
        else:
            kw[k] = float.fromhex(v) if isinstance(v, str) else v
    return oracle.opts(**kw)


@pytest.mark.parametrize("which", ["kr4", "s4"])
def test_oracle_rosenbrock_coeffs_match_reference_source(oracle, which):
    g = STIFF["rosenbrock_coeffs"][which]
    co = oracle.rosenbrock_coeffs(oracle.KR4 if which == "kr4" else oracle.S4)
    assert co["gamma"] == float.fromhex(g["gamma"])
    assert co["a"].tolist() == fh(g["a"]) and co["b"].tolist() == fh(g["b"]) and co["c"].tolist() == fh(g["c"])
    # the transcription slips of the reference's kr4 versus Kaps-Rentrop (a31 = 4.5247..., c32 = 0.15975068...) are kept
    if which == "kr4":
        assert co["a"][2][0] == 4.452470820736 and co["c"][2][1] == 0.1597500684673


@pytest.mark.parametrize("case", [c for c in STIFF["cases"] if c["kind"] == "ode23s"], ids=lambda c: c["name"])
def test_oracle_ode23s_goldens(oracle, case):
    ts = tspan_of(case["tspan"])
    r = oracle.solve_ode23s(getattr(oracle, case["system"].upper()), fh(case["params"]), fh(case["y0"]), ts,
                            _stiff_opts(oracle, case), tset=0 if case["tableau_set"] == "reference" else 1)
    assert r["rc"] == 0
    assert (r["accepted"], r["rejected"], r["status"], r["nout"]) == (case["accepted"], case["rejected"], case["status"], case["nout"])
    assert r["yfinal"].tolist() == fh(case["final"]) and r["t_reached"] == float.fromhex(case["t_reached"])
    for i, (t, y) in case["rows"].items():
        assert r["tout"][int(i)] == float.fromhex(t) and r["yout"][int(i)].tolist() == fh(y)
    assert digest(r["tout"]) == case["sha256_tout"] and digest(r["yout"]) == case["sha256_yout"]


@pytest.mark.parametrize("case", [c for c in STIFF["cases"] if c["kind"] == "rosenbrock"], ids=lambda c: c["name"])
def test_oracle_rosenbrock_goldens(oracle, case):
    ts = tspan_of(case["tspan"])
    r = oracle.solve_rosenbrock(oracle.KR4 if case["which"] == "kr4" else oracle.S4, getattr(oracle, case["system"].upper()),
                                fh(case["params"]), fh(case["y0"]), ts)
    assert (r["status"], len(r["yout"])) == (case["status"], case["nout"])
    for i, y in case["rows"].items():
        assert r["yout"][int(i)].tolist() == fh(y)
    assert digest(r["tout"]) == case["sha256_tout"] and digest(r["yout"]) == case["sha256_yout"]
    assert len(r["tout"]) == max(len(ts) - 1, 0)          # `tout: diff(tspan)` (problem.rs:572,639), one short of yout


def test_oracle_stiff_ensemble_matches_single(oracle):
    """The OpenMP ensemble drivers run the same per-trajectory code as the single-trajectory entry points."""
    mus = np.array([[0.5], [5.0], [50.0], [500.0]])
    y0 = np.tile([2.0, 0.0], (4, 1))
    ts = np.linspace(0.0, 30.0, 16)
    o = oracle.opts(1e-5, 1e-8)
    ens = oracle.ensemble_ode23s(oracle.VANDERPOL, mus, y0, ts, o, dense=True)
    for i in range(4):
        one = oracle.solve_ode23s(oracle.VANDERPOL, mus[i], y0[i], ts, o)
        assert np.array_equal(ens["yfinal"][i], one["yfinal"]) and ens["accepted"][i] == one["accepted"]
        assert ens["nrows"][i] == one["nout"] and ens["nfe"][i] == one["nfe"]
        rows = {t: y for t, y in zip(one["tout"], one["yout"])}
        for q in range(int(ens["nvalid"][i])):
            assert np.array_equal(ens["out"][:, q, i], rows[ts[q]])
    er = oracle.ensemble_rosenbrock(oracle.S4, oracle.VANDERPOL, mus, y0, np.linspace(0, 0.2, 21), want_all=True)
    for i in range(4):
        one = oracle.solve_rosenbrock(oracle.S4, oracle.VANDERPOL, mus[i], y0[i], np.linspace(0, 0.2, 21))
        assert np.array_equal(er["yall"][i], one["yout"]) and np.array_equal(er["yfinal"][i], one["yfinal"])


def test_ode23s_textbook_variant_converges():
    """TEXTBOOK ode23s (true W^-1) is a working stiff solver: Van der Pol mu = 100 against scipy's Radau."""
    import sys
    sys.path.insert(0, HERE)
    import pyref
    from scipy.integrate import solve_ivp
    mu = 100.0
    r = pyref.ode23s(pyref.vanderpol([mu]), [2.0, 0.0], [0.0, 150.0], reltol=1e-6, abstol=1e-9, true_inverse=True)
    ref = solve_ivp(lambda t, v: [v[1], mu * (1 - v[0] ** 2) * v[1] - v[0]], (0.0, 150.0), [2.0, 0.0], method="Radau",
                    rtol=1e-10, atol=1e-12)
    assert r["status"] == 0 and abs(r["y"][0] - ref.y[0, -1]) < 2e-3 * abs(ref.y[0, -1])


def test_reference_unit_tests_mirrored(oracle):
    """The reference's own unit tests, restated against the oracle: runge_kutta.rs:823-836 (`is_consistent`, `is_fsal`),
    problem.rs:1012-1034 (`diff_test`, `fdjacobian_test` on the Lorenz fixture Y0 = [0.1, 0, 0])."""
    import sys
    sys.path.insert(0, HERE)
    import pyref
    for m in ("FEULER", "MIDPOINT", "HEUN", "ODE21", "ODE4"):          # is_consistent_rk: a[(i, i-1)] == c[i]
        t = oracle.tableau(getattr(oracle, m))
        assert all(t["a"][i][i - 1] == t["c"][i] for i in range(1, t["S"])), m
    assert not oracle.tableau(oracle.MIDPOINT)["fsal"] and oracle.tableau(oracle.ODE45)["fsal"]
    r = oracle.solve_rosenbrock(oracle.S4, oracle.VANDERPOL, [1.0], [2.0, 0.0], [2.0, 6.0, 4.0, 16.0])
    assert r["tout"].tolist() == [4.0, -2.0, 12.0]                      # diff(&[2, 6, 4, 16])
    assert oracle.solve_rosenbrock(oracle.S4, oracle.VANDERPOL, [1.0], [2.0, 0.0], [])["tout"].tolist() == []
    J = oracle.fdjacobian(oracle.LORENZ, [10.0, 28.0, 8.0 / 3.0], 0.0, [0.1, 0.0, 0.0])
    assert J.shape == (3, 3)
    assert J.tolist() == pyref.fdjacobian(pyref.lorenz([10.0, 28.0, 8.0 / 3.0]), 0.0, [0.1, 0.0, 0.0])
    exact = np.array([[-10.0, 10.0, 0.0], [28.0, -1.0, -0.1], [0.0, 0.1, -8.0 / 3.0]])   # analytic Jacobian at Y0
    assert np.allclose(J, exact, rtol=0, atol=2e-2)                      # "crude" indeed: dx = 1 % of x (or of 1 where x = 0)


def test_allocation_faithful_variant_is_bit_identical(oracle):
    """solve_adapt_vec (the reference's Vec<f64> / clone pattern) must not change a bit against solve_adapt."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from diffeq_b200 import synth
    y0 = synth.lorenz_ics(0, 300)
    for method, tset, kw in ((oracle.ODE45, 0, {}), (oracle.ODE45FE, 0, dict(maxstep=0.05)), (oracle.ODE78, 1, dict(max_steps=500)),
                             (oracle.ODE23, 0, dict(max_steps=300)), (oracle.ODE45, 0, dict(libm_folded=1, initstep=1e-3))):
        o = oracle.opts(1e-7, 1e-9, **kw)
        a = oracle.ensemble_adaptive(method, oracle.LORENZ, synth.LORENZ_PARAMS, y0, [0.0, 0.3, 1.0], o, tset)
        b = oracle.ensemble_adaptive_vec(method, oracle.LORENZ, synth.LORENZ_PARAMS, y0, [0.0, 0.3, 1.0], o, tset)
        for k in ("yfinal", "accepted", "rejected", "nfe", "status"):
            assert np.array_equal(a[k], b[k], equal_nan=(k == "yfinal")), (method, k)
