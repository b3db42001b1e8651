"""Probe (not a test): a long seeded differential campaign, CUDA path vs CPU oracle, bit for bit.  Covers what the unit tests
cover, with many more random draws: explicit adaptive methods (all tableaus, both sets, Points::All / Specified, final / dense /
ragged output, both libm flavours), fixed-step methods, ode23s and oderosenbrock.
usage: python tests/gpu_fuzz.py [cases] [seed] [nvrtc] [fast]
   nvrtc: the systems come in as user CUDA source through NVRTC
   fast:  additionally run Mode.Fast on the adaptive cases and hold it to a sanity bar against the bit-exact result (it must
          finish, keep every trajectory's exit status, and take a total number of step attempts within 2 %)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import diffeq_b200 as D
from diffeq_b200 import synth
from oracle import oracle as O

LOR = np.array([10.0, 28.0, 8.0 / 3.0])
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
USER = {
    "Lorenz": (3, 3, "__device__ void rhs(double t, const double* v, double* d, const double* p) { const double x = v[0], y = v[1], z = v[2];"
                     " d[0] = p[0] * (y - x); d[1] = x * (p[1] - z) - y; d[2] = x * y - p[2] * z; }"),
    "VanDerPol": (2, 1, "__device__ void rhs(double t, const double* v, double* d, const double* p) { d[0] = v[1];"
                        " d[1] = p[0] * (1.0 - v[0] * v[0]) * v[1] - v[0]; }"),
    "Pendulum": (2, 3, "__device__ void rhs(double t, const double* v, double* d, const double* p) { d[0] = v[1];"
                       " d[1] = -p[0] * v[1] - deq_sin(v[0]) + p[1] * deq_cos(p[2] * t); }"),
}
RING = ("__device__ void rhs(double t, const double* v, double* d, const double* p) {\n#pragma unroll\n for (int i = 0; i < DEQ_USER_N; ++i) {"
        " const int ip = (i + 1) % DEQ_USER_N, im = (i + DEQ_USER_N - 1) % DEQ_USER_N; d[i] = p[0] * (v[ip] - v[i]) + p[1] * (v[i] * (1.0 - v[im])); } }")
NVRTC = "nvrtc" in sys.argv[3:]
FASTCHK = "fast" in sys.argv[3:]
fast_stats = []


def p_y0(p, i):
    return getattr(p, "_y0_dbg", [None])[i] if hasattr(p, "_y0_dbg") else None


def fast_check(tag, p, ode, o, tset, sol):
    """Mode.Fast against the PARITY solution `sol` of the same problem (sanity bar, see the module docstring)."""
    global bad
    f = p.solve(ode, D.OdeOptionMap(o).insert(D.Mode.Fast), tableau_set=tset)
    a0 = int(sol.accepted.sum(dtype=np.int64) + sol.rejected.sum(dtype=np.int64))
    a1 = int(f.accepted.sum(dtype=np.int64) + f.rejected.sum(dtype=np.int64))
    same_status = float((f.status == sol.status).mean())
    fin = np.isfinite(sol.final).all(axis=1) & (sol.status == 0) & (f.status == 0)
    dev = np.max(np.abs(f.final[fin] - sol.final[fin]) / (np.abs(sol.final[fin]) * float(o["Reltol"]) + float(o["Abstol"])), axis=1) if fin.any() else np.zeros(1)
    fast_stats.append((same_status, abs(a1 - a0) / max(a0, 1), float(np.median(dev)), float(dev.max())))
    # one attempt per trajectory of slack: with maxstep = span / integer the 1 % end-snap (problem.rs:337) sits on a knife edge and
    # FAST's dt * 0.01 vs PARITY's exact dt / 100 decides whether a last tiny step is taken; backward runs are excluded by the
    # caller (after the first rejection the reference runs away, chaotically)
    # (short runs — a handful of attempts per trajectory — are held to "no trajectory more than two attempts away" instead: with ~5
    # attempts each, one knife-edge attempt per trajectory plus a few second ones already exceeds any relative bar)
    per_traj = np.abs((f.accepted.astype(np.int64) + f.rejected) - (sol.accepted.astype(np.int64) + sol.rejected))
    if same_status < 0.9 or (abs(a1 - a0) > 0.02 * a0 + len(sol.accepted) and int(per_traj.max()) > 2):
        bad += 1
        print("FAST-MODE OUTLIER", tag, same_status, a0, a1, flush=True)
        if os.environ.get("DEQ_FUZZ_VERBOSE"):
            d = (f.accepted.astype(np.int64) + f.rejected) - (sol.accepted.astype(np.int64) + sol.rejected)
            i = int(np.argmax(np.abs(d)))
            print("   options", {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in o.items()}, "tspan", p.tspan[0], p.tspan[-1], len(p.tspan))
            print("   per-trajectory attempt differences: nonzero", int((d != 0).sum()), "of", len(d), "largest at", i, "parity acc/rej", int(sol.accepted[i]), int(sol.rejected[i]),
                  "fast acc/rej", int(f.accepted[i]), int(f.rejected[i]), "y0", p_y0(p, i), "final parity", sol.final[i], "fast", f.final[i], flush=True)
if NVRTC:   # plus the oracle's ring systems of dimension 8 and 16 (explicit methods only: the stiff path carries n <= 3)
    USER.update({"Ring8": (8, 2, RING), "Ring16": (16, 2, RING)})
RHS = {k: D.Rhs.from_source(src, n, np_) for k, (n, np_, src) in USER.items()} if NVRTC else {}
bad = 0
t_start = time.time()


def system(case, n, stiff=False):
    name = ["Lorenz", "VanDerPol", "Pendulum", "Ring8", "Ring16"][int(rng.integers(0, 5 if (NVRTC and not stiff) else 3))]
    if name.startswith("Ring"):
        d = int(name[4:])
        return name, rng.uniform(0.1, 0.9, (n, d)), np.column_stack([rng.uniform(0.1, 0.5, n), rng.uniform(0.5, 2.5, n)])
    if name == "Lorenz":
        return name, synth.lorenz_ics(case * 977, case * 977 + n), LOR
    if name == "VanDerPol":
        return name, np.tile([2.0, 0.0], (n, 1)) + rng.uniform(-1, 1, (n, 2)), 10.0 ** rng.uniform(-1, 1.5, (n, 1))
    return name, synth.pendulum_ics(case * 977, case * 977 + n), np.array(synth.PENDULUM_PARAMS)


def check(tag, pairs):
    global bad
    for name, a, b in pairs:
        a, b = np.asarray(a), np.asarray(b)
        if a.dtype == np.float64 and b.dtype == np.float64 and a.shape == b.shape:
            # bit patterns (so +0 / -0 are told apart); a NaN matches any NaN (payload and sign are platform details)
            same = (np.ascontiguousarray(a).view(np.uint64) == np.ascontiguousarray(b).view(np.uint64)) | (np.isnan(a) & np.isnan(b))
            ok_ = bool(same.all())
        else:
            ok_ = np.array_equal(a, b, equal_nan=True)
        if not ok_:
            bad += 1
            print("MISMATCH", name, tag, flush=True)
            try:   # keep the inputs of the failing case for analysis on the CPU
                os.makedirs("gpurun_out", exist_ok=True)
                np.savez(os.path.join("gpurun_out", "fuzz_fail_%d_%d_%s.npz" % (seed, tag[0], name)), got=a, want=b, tag=np.array(repr(tag)),
                         **{k: np.asarray(v) for k, v in CUR.items()})
            except Exception as e:   # noqa: BLE001
                print("could not save the failing case:", e, flush=True)
            return False
    return True


def grid(t0, span, npts, allow_dup=True):
    if npts < 2:
        return np.array([t0])
    g = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, max(npts - 2, 0))]))
    if allow_dup and npts > 4 and rng.random() < 0.2:
        g[2] = g[1]
    return t0 + span * g


CUR = {}
for case in range(cases):
    kind = rng.choice(["adapt", "adapt", "adapt", "fixed", "ode23s", "rosen"])
    n = int(rng.integers(1, 150)) if rng.random() < 0.97 else int(rng.integers(1000, 12000))   # a few multi-CTA ensembles (from 4096 on: the sorted dense dispatch)
    name, y0, params = system(case, n, stiff=kind in ("ode23s", "rosen"))
    if rng.random() < 0.04:   # special values in the initial state: NaN, infinities, signed zeros, huge and tiny magnitudes
        y0 = y0.copy()
        y0[int(rng.integers(0, n)), int(rng.integers(0, y0.shape[1]))] = float(rng.choice([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, 1e-300, 5e-324]))
    sysid, osys = getattr(D.System, name, None), getattr(O, name.upper())
    t0 = float(rng.uniform(-2, 2))
    span = float(10.0 ** rng.uniform(-2, 0.7))
    def prob(ts):
        pp = D.OdeProblem.builder().fun(RHS.get(name) or D.Rhs.builtin(sysid)).params(params).init(y0).tspan(ts).build()
        pp._y0_dbg = y0
        return pp
    if kind == "fixed":
        m = ["Feuler", "Heun", "Midpoint", "Ode4"][int(rng.integers(0, 4))]
        ts = grid(t0, span, int(rng.integers(2, 60)))
        ref = O.ensemble_fixed(getattr(O, m.upper()), osys, params, y0, ts)
        sol = prob(ts).solve(getattr(D.Ode, m), D.OdeOptionMap().insert(D.Output.Tspan))
        ok = check((case, kind, m, name, n), [("final", sol.final, ref)])
        if ok:
            one = O.solve_fixed(getattr(O, m.upper()), osys, params[0] if np.ndim(params) == 2 else params, y0[0], ts)
            check((case, kind, m, name, "rows"), [("rows", sol.yout[0], one["yout"])])
        continue
    if kind == "rosen":
        which = int(rng.integers(0, 2))
        ts = grid(t0, span * 0.05, int(rng.integers(1, 40)))
        ref = O.ensemble_rosenbrock(which, osys, params, y0, ts, want_all=True)
        sol = prob(ts).solve(D.Ode.Ode4skr if which == 0 else D.Ode.Ode4ss, D.OdeOptionMap().insert(D.Output.Tspan))
        m_ = np.arange(len(ts))[None, :] < sol.nvalid[:, None]
        check((case, kind, which, name, n), [("status", sol.status, ref["status"]), ("final", sol.final, ref["yfinal"]),
                                              ("rows", sol.yout[m_], ref["yall"][m_])])
        continue
    rtol, atol = 10.0 ** rng.uniform(-9, -3), 10.0 ** rng.uniform(-11, -4)
    folded = int(rng.random() < 0.4)
    kw = dict(max_steps=int(rng.choice([300, 1500, 4000])), libm_folded=folded)
    o = D.OdeOptionMap().insert(D.Reltol(rtol), D.Abstol(atol), D.MaxSteps(kw["max_steps"]), D.Refill(int(rng.integers(0, 3))))   # static / lane refill / + tail compaction and the dense-kernel choice
    if folded:
        o.insert(D.Libm.LlvmFolded)
    backward = rng.random() < 0.15
    if rng.random() < 0.3:
        kw["maxstep"] = span / float(rng.integers(3, 30)); o.insert(D.Maxstep(kw["maxstep"]))
    if rng.random() < 0.2:
        kw["minstep"] = span * 10.0 ** rng.uniform(-6, -3); o.insert(D.Minstep(kw["minstep"]))
    ts = grid(t0, span, int(rng.integers(2, 50)) if rng.random() < 0.95 else 1)
    if len(ts) < 2:
        ts = ts[:1]
    if backward:
        ts = ts[::-1].copy()
    tdir = 1.0 if ts[-1] >= ts[0] else -1.0
    if kind == "ode23s":
        tset = int(rng.integers(0, 2))
        pts = int(rng.random() < 0.4)
        if rng.random() < 0.3:
            kw["initstep"] = float(rng.choice([-1.0, 1.0]) * span * 10.0 ** rng.uniform(-5, -1)); o.insert(D.Initstep(kw["initstep"]))
        kw["points"] = pts
        o.insert(D.Points.Specified if pts else D.Points.All)
        oo = O.opts(rtol, atol, **kw)
        ref = O.ensemble_ode23s(osys, params, y0, ts, oo, tset=tset, dense=True)
        p = prob(ts)
        sol = p.solve(D.Ode.Ode23s, o, tableau_set=tset)
        tag = (case, kind, name, n, tset, folded, pts, backward)
        ok = check(tag, [(k, getattr(sol, k), ref[k]) for k in ("accepted", "rejected", "nfe", "status", "t_reached")] + [("final", sol.final, ref["yfinal"])])
        if FASTCHK and not folded and not backward:
            fast_check(tag, p, D.Ode.Ode23s, o, tset, sol)
        if ok:
            sa = p.solve(D.Ode.Ode23s, D.OdeOptionMap(o).insert(D.Output.All), tableau_set=tset)
            i = int(rng.integers(0, n))
            one = O.solve_ode23s(osys, params[i] if np.ndim(params) == 2 else params, y0[i], ts, oo, tset=tset)
            tout, yout = sa.trajectory(i)
            check(tag + ("stream", i), [("tout", tout, one["tout"]), ("yout", yout, one["yout"]), ("nrows", np.diff(sa.offsets).astype(np.uint32), ref["nrows"])])
            if len(ts) >= 2:
                sd = p.solve(D.Ode.Ode23s, D.OdeOptionMap(o).insert(D.Output.Tspan), tableau_set=tset)
                want = ref["out"].transpose(2, 1, 0)
                msk = (np.arange(len(ts))[None, :] < ref["nvalid"][:, None]) & ~np.isnan(want[:, :, 0])
                check(tag + ("dense",), [("nvalid", sd.nvalid, ref["nvalid"]), ("rows", sd.yout[msk], want[msk])])
            if rng.random() < 0.2:
                oh = D.OdeOptionMap(o).insert(D.ShardChunk(int(rng.choice([0, 1, 7, 64]))))
                hs = D.solve_host_multi(RHS.get(name) or D.Rhs.builtin(sysid), y0, params, ts, D.Ode.Ode23s, oh, devices=[0] * int(rng.integers(1, 5)), tableau_set=tset)
                check(tag + ("host",), [("final", hs["final"], ref["yfinal"]), ("accepted", hs["accepted"], ref["accepted"]), ("status", hs["status"], ref["status"])])
        continue
    # explicit adaptive
    m, tset = [("Ode45", 0), ("Ode45", 1), ("Ode45fe", 0), ("Ode23", 1), ("Ode23", 0), ("Ode21", 1), ("Ode21", 0), ("Ode78", 1), ("Ode78", 0)][int(rng.integers(0, 9))]
    if rng.random() < 0.3:
        kw["initstep"] = float(tdir * span * 10.0 ** rng.uniform(-5, -1)); o.insert(D.Initstep(kw["initstep"]))
    pts = int(rng.random() < 0.25)
    kw["points"] = pts
    if pts:
        o.insert(D.Points.Specified)
    oo = O.opts(rtol, atol, **kw)
    CUR = dict(y0=y0, params=params, ts=ts, rtol=rtol, atol=atol, method=m, tset=tset, system=name, opt_keys=list(kw.keys()),
               opt_vals=[float(v) for v in kw.values()], refill=int(o["Refill"]) if "Refill" in o else -1)
    ref = O.ensemble_adaptive(getattr(O, m.upper()), osys, params, y0, ts, oo, tset)
    p = prob(ts)
    sol = p.solve(getattr(D.Ode, m), o, tableau_set=tset)
    tag = (case, kind, m, tset, name, n, folded, pts, backward)
    ok = check(tag, [(k, getattr(sol, k), ref[k]) for k in ("accepted", "rejected", "nfe", "status", "t_reached")] + [("final", sol.final, ref["yfinal"])])
    if FASTCHK and not pts and not folded and not backward and (m, tset) not in (("Ode23", 0), ("Ode21", 0), ("Ode78", 0)):   # not the collapsing reference tableaus
        fast_check(tag, p, getattr(D.Ode, m), o, tset, sol)
    if ok:
        sa = p.solve(getattr(D.Ode, m), D.OdeOptionMap(o).insert(D.Output.All), tableau_set=tset)
        i = int(rng.integers(0, n))
        one = O.solve_adaptive(getattr(O, m.upper()), osys, params[i] if np.ndim(params) == 2 else params, y0[i], ts, oo, tset=tset)
        tout, yout = sa.trajectory(i)
        check(tag + ("stream", i), [("tout", tout, one["tout"]), ("yout", yout, one["yout"])])
        if rng.random() < 0.15:
            oh = D.OdeOptionMap(o).insert(D.ShardChunk(int(rng.choice([0, 1, 7, 64]))))
            hs = D.solve_host_multi(RHS.get(name) or D.Rhs.builtin(sysid), y0, params, ts, getattr(D.Ode, m), oh, devices=[0] * int(rng.integers(1, 5)), tableau_set=tset)
            check(tag + ("host",), [("final", hs["final"], ref["yfinal"]), ("accepted", hs["accepted"], ref["accepted"]), ("t_reached", hs["t_reached"], ref["t_reached"])])
        if not pts and len(ts) >= 2:
            for lay in (D.DenseLayout.SoA, D.DenseLayout.TrajMajor):
                sd = p.solve(getattr(D.Ode, m), D.OdeOptionMap(o).insert(D.Output.Tspan, lay), tableau_set=tset)
                rd = O.ensemble_dense(getattr(O, m.upper()), osys, params, y0, ts, oo, tset)
                want = rd["out"].transpose(2, 1, 0)
                msk = ~np.isnan(want) & ~np.isnan(sd.yout)   # rows the reference never emitted are unset on either side
                check(tag + ("dense", int(lay)), [("nvalid", sd.nvalid, rd["nvalid"]), ("rows", sd.yout[msk], want[msk])])
if fast_stats:
    fs = np.array(fast_stats)
    print(f"fast-mode sanity: {len(fs)} adaptive cases; same exit status mean {fs[:, 0].mean():.4f} min {fs[:, 0].min():.3f}; step-attempt total differs by "
          f"mean {fs[:, 1].mean():.2e} max {fs[:, 1].max():.2e}; final-state deviation in units of (rtol |y| + atol): median of medians {np.median(fs[:, 2]):.2e}, "
          f"median of maxima {np.median(fs[:, 3]):.2e}")
print(f"fuzz done: {cases} cases, seed {seed}, mismatches {bad}, {time.time() - t_start:.1f} s")
sys.exit(1 if bad else 0)
