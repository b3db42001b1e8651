"""A bounded slice of the differential campaign (tests/gpu_fuzz.py; the long runs are logged in profiles/r1_fuzz_campaign.txt)
as part of the GPU suite: fresh seeds, every method and output kind, bit patterns compared against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("args", [["400", "101"], ["60", "102", "nvrtc"], ["300", "103", "fast"]], ids=["builtin", "nvrtc", "fast-sanity"])
def test_differential_fuzz_slice(deq, args):
    out = subprocess.run([sys.executable, os.path.join(HERE, "gpu_fuzz.py")] + args, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "mismatches 0" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _replay_options(D, O, c):
    kw = {str(k): float(v) for k, v in zip(c["opt_keys"], c["opt_vals"])}
    rtol, atol = float(c["rtol"]), float(c["atol"])
    o = D.OdeOptionMap().insert(D.Reltol(rtol), D.Abstol(atol), D.MaxSteps(int(kw["max_steps"])))
    if int(kw.get("libm_folded", 0)):
        o.insert(D.Libm.LlvmFolded)
    for key, opt in (("maxstep", D.Maxstep), ("minstep", D.Minstep), ("initstep", D.Initstep)):
        if key in kw:
            o.insert(opt(kw[key]))
    if int(kw.get("points", 0)):
        o.insert(D.Points.Specified)
    okw = dict(kw)
    for k in ("max_steps", "libm_folded", "points"):
        okw[k] = int(okw.get(k, 0))
    return o, O.opts(rtol, atol, **okw), int(okw["points"])


@pytest.mark.parametrize("case", sorted(f for f in os.listdir(os.path.join(HERE, "golden", "fuzz_regressions")) if f.endswith(".npz")))
def test_fuzz_regressions(deq, oracle, case):
    """Inputs of the cases a differential campaign once caught (tests/gpu_fuzz.py keeps them), replayed against the oracle
    through every scheduling flavour and both dense layouts.  Round 2: `nvalid` of the row-synchronous SoA kernel when the last
    step ends a rounding error PAST tspan[nt-1] (`t + (tend - t) > tend`: the reference's emission loop, problem.rs:303-319,
    then counts the end point as an interior row); sin of a huge argument (glibc's Payne–Hanek range reduction)."""
    import numpy as np
    D, O = deq, oracle
    c = np.load(os.path.join(HERE, "golden", "fuzz_regressions", case), allow_pickle=True)
    o, oo, pts = _replay_options(D, O, c)
    method, tset, name = str(c["method"]), int(c["tset"]), str(c["system"])
    y0, params, ts = c["y0"], c["params"], c["ts"]
    osys, om = getattr(O, name.upper()), getattr(O, method.upper())
    ref = O.ensemble_adaptive(om, osys, params, y0, ts, oo, tset)
    rd = None if pts else O.ensemble_dense(om, osys, params, y0, ts, oo, tset)
    for refill in (0, 1, 2):
        p = D.OdeProblem.builder().fun(D.Rhs.builtin(getattr(D.System, name))).params(params).init(y0).tspan(ts).build()
        sol = p.solve(getattr(D.Ode, method), D.OdeOptionMap(o).insert(D.Refill(refill)), tableau_set=tset)
        for k in ("accepted", "rejected", "nfe", "status", "t_reached"):
            assert np.array_equal(getattr(sol, k), ref[k], equal_nan=True), (case, refill, k)
        assert np.array_equal(sol.final, ref["yfinal"], equal_nan=True), (case, refill)
        if rd is None:
            continue
        want = rd["out"].transpose(2, 1, 0)
        for lay in (D.DenseLayout.SoA, D.DenseLayout.TrajMajor):
            sd = p.solve(getattr(D.Ode, method), D.OdeOptionMap(o).insert(D.Refill(refill), D.Output.Tspan, lay), tableau_set=tset)
            assert np.array_equal(sd.nvalid, rd["nvalid"]), (case, refill, int(lay))
            msk = ~np.isnan(want) & ~np.isnan(sd.yout)
            assert np.array_equal(sd.yout[msk], want[msk]), (case, refill, int(lay))
