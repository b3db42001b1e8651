"""Probe (not a test, CPU only): seeded differential campaign between the two independent restatements of the reference —
tests/pyref.py (pure Python, tableaus parsed from the reference source) and oracle/diffeq_oracle.cpp — on single trajectories:
explicit adaptive methods (both tableau sets, Points::All / Specified, both libm flavours), fixed-step methods, ode23s (both
W-inverse variants) and oderosenbrock.  Every (tout, yout) row must agree bit for bit.
usage: python tests/cpu_fuzz_pyref.py [cases] [seed]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

import pyref
from oracle import oracle as O

fh = lambda v: [fh(x) for x in v] if isinstance(v, list) else (float.fromhex(v) if isinstance(v, str) else v)  # noqa: E731
TABS = json.load(open(os.path.join(HERE, "golden", "tableaus.json")))
TABS = {k: {m: dict(t, a=fh(t["a"]), b=fh(t["b"]), c=fh(t["c"])) for m, t in v.items()} for k, v in TABS.items()}
NAME2TAB = {"Feuler": "feuler", "Heun": "heun", "Midpoint": "midpoint", "Ode4": "rk4", "Ode21": "rk21", "Ode23": "rk23",
            "Ode45": "dopri5", "Ode45fe": "rk45", "Ode78": "feh78"}
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 500
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = 0
t_start = time.time()


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    return bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())


for case in range(cases):
    kind = rng.choice(["adapt", "adapt", "fixed", "ode23s", "rosen"])
    sysname = ["Lorenz", "VanDerPol", "Pendulum", "Ring8", "Ring16"][int(rng.integers(0, 3 if kind in ("ode23s", "rosen") else 5))]
    if sysname == "Lorenz":
        prm, y0, f = [10.0, 28.0, 8.0 / 3.0], rng.uniform([-20, -20, 0], [20, 20, 50]).tolist(), None
        f = pyref.lorenz(prm)
    elif sysname == "VanDerPol":
        prm, y0 = [float(10.0 ** rng.uniform(-1, 2.5))], (np.array([2.0, 0.0]) + rng.uniform(-1, 1, 2)).tolist()
        f = pyref.vanderpol(prm)
    elif sysname.startswith("Ring"):
        d = int(sysname[4:])
        prm, y0 = [float(rng.uniform(0.1, 0.5)), float(rng.uniform(0.5, 2.5))], rng.uniform(0.1, 0.9, d).tolist()
        f = pyref.ring(prm, d)
    else:
        prm, y0 = [0.5, 1.2, 2.0 / 3.0], rng.uniform([-3.1, -2], [3.1, 2]).tolist()
        f = pyref.pendulum(prm)
    osys = getattr(O, sysname.upper())
    t0, span = float(rng.uniform(-2, 2)), float(10.0 ** rng.uniform(-2, 0.5))
    npts = int(rng.integers(2, 30)) if rng.random() < 0.9 else 1
    g = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, max(npts - 2, 0))])) if npts >= 2 else np.array([0.0])
    if npts > 4 and rng.random() < 0.2:
        g[2] = g[1]
    ts = (t0 + span * g)
    if rng.random() < 0.15:
        ts = ts[::-1].copy()
    ts = ts.tolist()
    tag = (case, kind, sysname, len(ts))
    if kind == "fixed":
        m = ["Feuler", "Heun", "Midpoint", "Ode4"][int(rng.integers(0, 4))]
        tout, ys = pyref.oderk_fixed(f, TABS["reference"][NAME2TAB[m]], y0, ts)
        r = O.solve_fixed(getattr(O, m.upper()), osys, prm, y0, ts)
        if not same(ys, r["yout"]):
            bad += 1; print("MISMATCH", tag, m)
        continue
    if kind == "rosen":
        which = int(rng.integers(0, 2))
        tsr = (t0 + 0.05 * span * g).tolist()
        pr = pyref.oderosenbrock(f, pyref.ROSENBROCK["kr4" if which == 0 else "s4"], y0, tsr)
        r = O.solve_rosenbrock(which, osys, prm, y0, tsr)
        if not (same(pr["yout"], r["yout"]) and pr["status"] == r["status"] and same(pr["tout"], r["tout"])):
            bad += 1; print("MISMATCH", tag, which)
        continue
    rtol, atol = float(10.0 ** rng.uniform(-8, -3)), float(10.0 ** rng.uniform(-10, -4))
    folded = bool(rng.random() < 0.4)
    kw = dict(reltol=rtol, abstol=atol, max_steps=int(rng.choice([100, 300, 600])), libm_folded=folded)
    tdir = 1.0 if ts[-1] >= ts[0] else -1.0
    if rng.random() < 0.3:
        kw["maxstep"] = span / float(rng.integers(3, 30))
    if rng.random() < 0.2:
        kw["minstep"] = span * float(10.0 ** rng.uniform(-6, -3))
    pts_all = bool(rng.random() < 0.7)
    if kind == "ode23s":
        ti = bool(rng.integers(0, 2))
        if rng.random() < 0.3:
            kw["initstep"] = float(rng.choice([-1.0, 1.0]) * span * 10.0 ** rng.uniform(-5, -1))
        pr = pyref.ode23s(f, y0, ts, points_all=pts_all, true_inverse=ti, **kw)
        oo = O.opts(rtol, atol, minstep=kw.get("minstep"), maxstep=kw.get("maxstep"), initstep=kw.get("initstep", 0.0),
                    points=0 if pts_all else 1, max_steps=kw["max_steps"], libm_folded=int(folded))
        r = O.solve_ode23s(osys, prm, y0, ts, oo, tset=int(ti))
        ok = same(pr["tout"], r["tout"]) and same(pr["yout"], r["yout"]) and (pr["accepted"], pr["rejected"], pr["status"]) == (r["accepted"], r["rejected"], r["status"])
        ok = ok and same(pr["y"], r["yfinal"]) and same([pr["t"]], [r["t_reached"]])
        if not ok:
            bad += 1; print("MISMATCH", tag, ti, folded, pts_all, kw)
        continue
    m, tset = [("Ode45", "reference"), ("Ode45", "textbook"), ("Ode45fe", "reference"), ("Ode23", "textbook"), ("Ode23", "reference"),
               ("Ode21", "textbook"), ("Ode21", "reference"), ("Ode78", "textbook"), ("Ode78", "reference")][int(rng.integers(0, 9))]
    if rng.random() < 0.3:
        kw["initstep"] = float(tdir * span * 10.0 ** rng.uniform(-5, -1))
    pr = pyref.oderk_adapt(f, TABS[tset][NAME2TAB[m]], y0, ts, points_all=pts_all, **kw)
    oo = O.opts(rtol, atol, minstep=kw.get("minstep"), maxstep=kw.get("maxstep"), initstep=kw.get("initstep", 0.0),
                points=0 if pts_all else 1, max_steps=kw["max_steps"], libm_folded=int(folded))
    r = O.solve_adaptive(getattr(O, m.upper()), osys, prm, y0, ts, oo, tset=0 if tset == "reference" else 1)
    ok = same(pr["tout"], r["tout"]) and same(pr["yout"], r["yout"]) and (pr["accepted"], pr["rejected"], pr["status"]) == (r["accepted"], r["rejected"], r["status"])
    if not ok:
        bad += 1; print("MISMATCH", tag, m, tset, folded, pts_all, kw)
print(f"pyref-vs-oracle fuzz done: {cases} cases, seed {seed}, mismatches {bad}, {time.time() - t_start:.1f} s")
sys.exit(1 if bad else 0)
