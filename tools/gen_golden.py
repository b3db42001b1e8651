#!/usr/bin/env python
"""gen_golden.py — generates tests/golden/*.json (run in the build container, where /root/reference exists).

1. tableaus.json: every Butcher tableau PARSED FROM THE REFERENCE SOURCE (src/ode/runge_kutta.rs): the f64 literal
   expressions are evaluated in IEEE double (Python floats) and laid out with nalgebra's row-major constructor
   semantics.  This pins both the oracle's and the product's transcription of the coefficients to the reference.
2. solver_cases.json: known-answer vectors for the RK hot path computed by tests/pyref.py (pure-Python restatement,
   using the parsed tableaus) — bit patterns as hex floats.  The C++ oracle must reproduce them exactly.
The reference itself cannot be executed here (no Rust toolchain), so these are restatement-derived goldens; the
oracle header and DESIGN.md say "parity unpinned by the reference".
"""
import hashlib
import json
import os
import re
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref  # noqa: E402

REF = "/root/reference/src/ode/runge_kutta.rs"
OUT = os.path.join(ROOT, "tests", "golden")


def _args(body, start_token):
    """Returns the comma-separated numeric expressions inside the first (...) or &[...] after start_token."""
    i = body.index(start_token)
    j = body.index("(", i)
    depth, k = 0, j
    while True:
        ch = body[k]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                break
        k += 1
    inner = body[j + 1:k]
    while "(" in inner and "&[" not in inner:  # Weights::Explicit(Vector2::new(..)) -> innermost argument list
        inner = inner[inner.index("(") + 1:inner.rindex(")")]
    if "&[" in inner:
        inner = inner[inner.index("&[") + 2:inner.rindex("]")]
    inner = re.sub(r"U\d+\s*,", "", inner)  # generic-dimension arguments of from_row_slice_generic
    vals = []
    for tok in inner.split(","):
        tok = tok.strip().replace("_", "")
        if not tok:
            continue
        assert re.fullmatch(r"-?[0-9.]+(\s*/\s*[0-9.]+)?", tok), tok
        vals.append(float(eval(tok)))  # noqa: S307 - numeric literal expression checked by the regex above
    return vals


def parse_tableaus():
    src = open(REF).read()
    fns = {}
    for m in re.finditer(r"pub fn (\w+)\(\) -> Self \{", src):
        name = m.group(1)
        end = src.index("Self {", m.end())
        fns[name] = src[m.end():end]
    meta = {"feuler": (1, False, 1), "midpoint": (2, False, 2), "heun": (2, False, 2), "rk21": (2, True, 1),
            "rk23": (4, True, 2), "rk4": (4, False, 4), "rk45": (6, True, 4), "dopri5": (7, True, 4), "feh78": (13, True, 7)}
    out = {}
    for name, (S, adaptive, order_min) in meta.items():
        body = fns[name]
        if name == "feuler":
            a, b, c = [0.0], [1.0], [0.0]  # Matrix1::zero(), Vector1::one(), Vector1::zero()
        else:
            a = _args(body, "let a =")
            b = _args(body, "let b =")
            c = _args(body, "let c =")
        assert len(a) == S * S and len(c) == S and len(b) == (2 * S if adaptive else S), (name, len(a), len(b), len(c))
        A = [[a[r * S + cc] for cc in range(S)] for r in range(S)]            # row-major constructors
        B = [[b[2 * s], b[2 * s + 1]] for s in range(S)] if adaptive else [[b[s], 0.0] for s in range(S)]
        out[name] = dict(S=S, adaptive=adaptive, order_min=order_min, a=A, b=B, c=c)
        out[name]["fsal"] = pyref.is_fsal(out[name])
    return out


def textbook(tabs):
    """the corrected coefficients."""
    import copy
    t = copy.deepcopy(tabs)
    t["rk21"]["b"] = [[0.5, 1.0], [0.5, 0.0]]
    t["rk23"]["b"][0][1] = 2.0 / 9.0
    t["dopri5"]["c"][3] = 4.0 / 5.0
    f = t["feh78"]
    f["a"][5][0], f["a"][5][3], f["a"][5][4] = 1.0 / 20.0, 1.0 / 4.0, 1.0 / 5.0
    b7 = [41. / 840., 0., 0., 0., 0., 34. / 105., 9. / 35., 9. / 35., 9. / 280., 9. / 280., 41. / 840., 0., 0.]
    b8 = [0., 0., 0., 0., 0., 34. / 105., 9. / 35., 9. / 35., 9. / 280., 9. / 280., 0., 41. / 840., 41. / 840.]
    f["b"] = [[b7[s], b8[s]] for s in range(13)]
    for v in t.values():
        v["fsal"] = pyref.is_fsal(v)
    return t


def hx(v):
    return [hx(x) for x in v] if isinstance(v, (list, tuple)) else float(v).hex()


def digest(rows):
    h = hashlib.sha256()
    for r in rows:
        for x in (r if isinstance(r, (list, tuple)) else [r]):
            h.update(struct.pack("<d", x))
    return h.hexdigest()


def linspace(a, b, n):
    step = (b - a) / (n - 1)
    return [a + step * i for i in range(n)]


METHOD_TAB = {"Feuler": "feuler", "Heun": "heun", "Midpoint": "midpoint", "Ode4": "rk4", "Ode21": "rk21", "Ode23": "rk23",
              "Ode45": "dopri5", "Ode45fe": "rk45", "Ode78": "feh78"}
SYS = {"Lorenz": pyref.lorenz, "VanDerPol": pyref.vanderpol, "Pendulum": pyref.pendulum,
       "Cvdp2": lambda p: pyref.coupled_vdp(p, 2), "Cvdp3": lambda p: pyref.coupled_vdp(p, 3),
       "Cvdp6": lambda p: pyref.coupled_vdp(p, 6), "Cvdp12": lambda p: pyref.coupled_vdp(p, 12)}


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = parse_tableaus()
    txt = textbook(ref)
    json.dump({"reference": {k: dict(v, a=hx(v["a"]), b=hx(v["b"]), c=hx(v["c"])) for k, v in ref.items()},
               "textbook": {k: dict(v, a=hx(v["a"]), b=hx(v["b"]), c=hx(v["c"])) for k, v in txt.items()}},
              open(os.path.join(OUT, "tableaus.json"), "w"), indent=0)
    LOR = [10.0, 28.0, 8.0 / 3.0]
    cases = []

    def fixed(name, method, system, params, y0, tspan_desc, tspan, keep_every):
        tout, ys = pyref.oderk_fixed(SYS[system](params), ref[METHOD_TAB[method]], y0, tspan)
        cases.append(dict(name=name, kind="fixed", method=method, system=system, params=hx(params), y0=hx(y0),
                          tspan=tspan_desc, final=hx(ys[-1]), rows={str(i): hx(ys[i]) for i in range(0, len(ys), keep_every)},
                          sha256_yout=digest(ys)))

    def adapt(name, method, system, params, y0, tspan_desc, tspan, tset="reference", **kw):
        tb = (ref if tset == "reference" else txt)[METHOD_TAB[method]]
        r = pyref.oderk_adapt(SYS[system](params), tb, y0, tspan, **kw)
        n = len(r["tout"])
        keep = sorted(set([0, 1, 2, n // 2, n - 2, n - 1]) & set(range(n)))
        cases.append(dict(name=name, kind="adaptive", method=method, system=system, tableau_set=tset, params=hx(params),
                          y0=hx(y0), tspan=tspan_desc, options={k: (hx(v) if isinstance(v, float) else v) for k, v in kw.items()},
                          accepted=r["accepted"], rejected=r["rejected"], status=r["status"], nout=n,
                          final=hx(r["yout"][-1]) if n else [], t_last=hx(r["tout"][-1]) if n else None,
                          rows={str(i): [hx(r["tout"][i]), hx(r["yout"][i])] for i in keep},
                          sha256_tout=digest(r["tout"]), sha256_yout=digest(r["yout"])))

    # BASELINE config 1 — the reference's own fixture (problem.rs:978-1010)
    fixed("cfg1_rk4_lorenz_t100", "Ode4", "Lorenz", LOR, [0.1, 0.0, 0.0], ["linspace", 0.0, 100.0, 100001],
          linspace(0.0, 100.0, 100001), 10000)
    for m in ("Feuler", "Heun", "Midpoint", "Ode4"):
        fixed(f"fixed_{m}_lorenz", m, "Lorenz", LOR, [1.0, -2.0, 20.0], ["linspace", 0.0, 2.0, 401], linspace(0.0, 2.0, 401), 100)
        fixed(f"fixed_{m}_vdp", m, "VanDerPol", [1.5], [2.0, 0.0], ["linspace", 0.0, 5.0, 501], linspace(0.0, 5.0, 501), 100)
    fixed("fixed_Ode4_pendulum", "Ode4", "Pendulum", [0.5, 1.2, 2.0 / 3.0], [0.3, 0.1], ["linspace", 0.0, 5.0, 501],
          linspace(0.0, 5.0, 501), 100)
    # adaptive
    adapt("dopri5_lorenz_t10_1e-8", "Ode45", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 10.0], [0.0, 10.0], reltol=1e-8, abstol=1e-8)
    adapt("rk45_lorenz_t10_1e-8", "Ode45fe", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 10.0], [0.0, 10.0], reltol=1e-8, abstol=1e-8)
    adapt("dopri5_lorenz_t100_default", "Ode45", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 100.0], [0.0, 100.0])
    # backward integration: after the first rejection the timeout clamp `new_dt.min(dt)` (problem.rs:835-838) flips the sign
    # of dt and the reference runs away forward forever — pinned bug-for-bug under the guard
    adapt("dopri5_lorenz_backward_runaway", "Ode45", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 1.0, 0.0], [1.0, 0.0], reltol=1e-6,
          max_steps=2000)
    adapt("dopri5_lorenz_backward_short", "Ode45", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 0.05, 0.0], [0.05, 0.0], reltol=1e-4)
    adapt("dopri5_lorenz_initstep", "Ode45", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 0.0, 1.0], [0.0, 1.0], initstep=1e-3)
    adapt("dopri5_lorenz_maxstep", "Ode45", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 0.0, 1.0], [0.0, 1.0], maxstep=0.01)
    adapt("dopri5_lorenz_minstep_break", "Ode45", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 0.0, 5.0], [0.0, 5.0], initstep=0.5,
          minstep=5e-2)
    adapt("dopri5_zero_span", "Ode45", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 1.0, 1.0], [1.0, 1.0])
    adapt("dopri5_textbook_pendulum", "Ode45", "Pendulum", [0.5, 1.2, 2.0 / 3.0], [0.3, 0.1], ["list", 0.0, 10.0], [0.0, 10.0],
          tset="textbook", reltol=1e-8, abstol=1e-8)
    adapt("dopri5_reference_pendulum", "Ode45", "Pendulum", [0.5, 1.2, 2.0 / 3.0], [0.3, 0.1], ["list", 0.0, 10.0], [0.0, 10.0],
          reltol=1e-8, abstol=1e-8)
    for mu in (0.1, 1.0, 10.0):
        adapt(f"rk45_vdp_mu{mu}_dense1000", "Ode45fe", "VanDerPol", [mu], [2.0, 0.0], ["linspace", 0.0, 20.0, 1000],
              linspace(0.0, 20.0, 1000), reltol=1e-6, abstol=1e-8)
    adapt("rk45_vdp_specified_points", "Ode45fe", "VanDerPol", [1.0], [2.0, 0.0], ["linspace", 0.0, 20.0, 50],
          linspace(0.0, 20.0, 50), points_all=False)
    adapt("feh78_textbook_pendulum_1e-12", "Ode78", "Pendulum", [0.5, 1.2, 2.0 / 3.0], [0.3, 0.1], ["list", 0.0, 10.0], [0.0, 10.0],
          tset="textbook", reltol=1e-12, abstol=1e-12)
    adapt("rk23_textbook_lorenz", "Ode23", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 2.0], [0.0, 2.0], tset="textbook")
    adapt("rk21_textbook_lorenz", "Ode21", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 1.0], [0.0, 1.0], tset="textbook", reltol=1e-3,
          abstol=1e-3)
    # the libm calls as an LLVM-optimised build folds them (pow(x,0.5) -> sqrt, pow(10,x) -> exp10)
    adapt("dopri5_lorenz_t10_1e-8_libm_folded", "Ode45", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 10.0], [0.0, 10.0], reltol=1e-8,
          abstol=1e-8, libm_folded=True)
    adapt("dopri5_lorenz_t100_default_libm_folded", "Ode45", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 100.0], [0.0, 100.0],
          libm_folded=True)
    adapt("rk45_vdp_mu1.0_dense1000_libm_folded", "Ode45fe", "VanDerPol", [1.0], [2.0, 0.0], ["linspace", 0.0, 20.0, 1000],
          linspace(0.0, 20.0, 1000), reltol=1e-6, abstol=1e-8, libm_folded=True)
    for m in ("Ode21", "Ode23", "Ode78"):  # the reference's broken tableaus: step-size collapse, guarded
        adapt(f"{m}_reference_collapse", m, "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 0.0, 0.01], [0.0, 0.01], reltol=1e-3, abstol=1e-3,
              max_steps=3000)
    json.dump(cases, open(os.path.join(OUT, "solver_cases.json"), "w"), indent=0)
    print(f"wrote {len(cases)} solver cases and {len(ref)} x 2 tableaus to {OUT}")


def parse_rosenbrock():
    """RosenbrockCoeffs::kr4 / s4 parsed from src/ode/rosenbrock.rs (Matrix4::new / Vector4::new are row-major)."""
    src = open(os.path.join(os.path.dirname(REF), "rosenbrock.rs")).read()
    out = {}
    for m in re.finditer(r"pub fn (\w+)\(\) -> Self \{", src):
        name = m.group(1)
        end = src.index("Self {", m.end())
        body = src[m.end():end]
        tail = src[end:src.index("}", end)]
        gamma = float(re.search(r"gamma:\s*([0-9.]+)", tail).group(1))
        a, b, c = _args(body, "let a ="), _args(body, "let b ="), _args(body, "let c =")
        assert len(a) == 16 and len(b) == 4 and len(c) == 16
        out[name] = dict(gamma=gamma, a=[a[4 * i:4 * i + 4] for i in range(4)], b=b, c=[c[4 * i:4 * i + 4] for i in range(4)])
    return out


def main_stiff():
    """stiff_cases.json: ode23s / oderosenbrock known-answer vectors from tests/pyref.py."""
    ros = parse_rosenbrock()
    assert ros["kr4"] == pyref.ROSENBROCK["kr4"] and ros["s4"] == pyref.ROSENBROCK["s4"], "pyref coefficients differ from rosenbrock.rs"
    LOR = [10.0, 28.0, 8.0 / 3.0]
    PEN = [0.5, 1.2, 2.0 / 3.0]
    cases = []

    def ode23s(name, system, params, y0, tspan_desc, tspan, tset="reference", **kw):
        r = pyref.ode23s(SYS[system](params), y0, tspan, true_inverse=(tset == "textbook"), **kw)
        n = len(r["tout"])
        keep = sorted(set([0, 1, 2, n // 2, n - 2, n - 1]) & set(range(n)))
        cases.append(dict(name=name, kind="ode23s", system=system, tableau_set=tset, params=hx(params), y0=hx(y0), tspan=tspan_desc,
                          options={k: (hx(v) if isinstance(v, float) else v) for k, v in kw.items()},
                          accepted=r["accepted"], rejected=r["rejected"], status=r["status"], nout=n, final=hx(r["y"]),
                          t_reached=hx(r["t"]), rows={str(i): [hx(r["tout"][i]), hx(r["yout"][i])] for i in keep},
                          sha256_tout=digest(r["tout"]), sha256_yout=digest(r["yout"])))

    def rosen(name, which, system, params, y0, tspan_desc, tspan):
        r = pyref.oderosenbrock(SYS[system](params), ros[which], y0, tspan)
        n = len(r["yout"])
        keep = sorted(set([0, 1, 2, n // 2, n - 1]) & set(range(n)))
        cases.append(dict(name=name, kind="rosenbrock", which=which, system=system, params=hx(params), y0=hx(y0), tspan=tspan_desc,
                          status=r["status"], nout=n, final=hx(r["yout"][-1]), rows={str(i): hx(r["yout"][i]) for i in keep},
                          sha256_tout=digest(r["tout"]), sha256_yout=digest(r["yout"])))

    for tset in ("reference", "textbook"):
        for mu in (1.0, 10.0, 1000.0):
            ode23s(f"ode23s_{tset}_vdp_mu{mu}", "VanDerPol", [mu], [2.0, 0.0], ["list", 0.0, 20.0], [0.0, 20.0], tset=tset)
        ode23s(f"ode23s_{tset}_vdp_mu100_dense", "VanDerPol", [100.0], [2.0, 0.0], ["linspace", 0.0, 200.0, 400],
               linspace(0.0, 200.0, 400), tset=tset, reltol=1e-6, abstol=1e-8)
        ode23s(f"ode23s_{tset}_lorenz", "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0, 1.0, 2.5, 5.0], [0.0, 1.0, 2.5, 5.0], tset=tset)
        ode23s(f"ode23s_{tset}_pendulum", "Pendulum", PEN, [1.0, 0.5], ["list", 0.0, 3.0, 10.0], [0.0, 3.0, 10.0], tset=tset,
               reltol=1e-7, abstol=1e-9)
    ode23s("ode23s_specified_points", "VanDerPol", [5.0], [2.0, 0.0], ["linspace", 0.0, 10.0, 101], linspace(0.0, 10.0, 101),
           points_all=False)
    ode23s("ode23s_backward", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 0.0, -0.5, -1.0], [0.0, -0.5, -1.0], max_steps=3000)
    ode23s("ode23s_initstep_maxstep", "VanDerPol", [3.0], [2.0, 0.0], ["list", 0.0, 5.0], [0.0, 5.0], initstep=0.01, maxstep=0.05)
    ode23s("ode23s_negative_initstep_is_abs", "VanDerPol", [3.0], [2.0, 0.0], ["list", 0.0, 1.0], [0.0, 1.0], initstep=-0.01)
    ode23s("ode23s_minstep_exit", "VanDerPol", [50.0], [2.0, 0.0], ["list", 0.0, 100.0], [0.0, 100.0], minstep=0.05)
    ode23s("ode23s_zero_state_jacobian", "Lorenz", LOR, [0.0, 0.0, 0.0], ["list", 0.0, 1.0], [0.0, 1.0], max_steps=500)
    ode23s("ode23s_libm_folded", "VanDerPol", [10.0], [2.0, 0.0], ["list", 0.0, 20.0], [0.0, 20.0], libm_folded=True)
    # OdeError::InvalidMatrix (problem.rs:481,588): Lorenz at the origin with rho = 0, beta = -4 has d f_z / d z = 4 exactly, so
    # W_zz = 1 - 4 (h d) vanishes for h d = 0.25 (h = 0x1.b504f333f9de6p-1) and (1/(gamma h) - 4) for gamma h = 0.25
    ode23s("ode23s_invalid_matrix", "Lorenz", [10.0, 0.0, -4.0], [0.0, 0.0, 0.0], ["list", 0.0, 10.0], [0.0, 10.0],
           initstep=float.fromhex("0x1.b504f333f9de6p-1"), maxstep=5.0)
    ode23s("ode23s_minstep_midrun", "VanDerPol", [1000.0], [2.0, 0.0], ["list", 0.0, 2000.0], [0.0, 2000.0], minstep=3e-4)
    # backward run that overshoots tfinal by 8 ulp and returns with a tiny step h > 0: the full tspan rescan (problem.rs:519-521)
    # then pushes the row for tfinal although the run is backward (found by tests/gpu_fuzz.py, seed 8 case 1987)
    fh_ = float.fromhex
    ode23s("ode23s_backward_overshoot_correction", "VanDerPol", [fh_("0x1.a989391c904bbp-1")],
           [fh_("0x1.f9432062ca40fp+0"), fh_("-0x1.4b5804af8ef58p-2")], ["list", fh_("0x1.0a4941f00af94p+1"), fh_("-0x1.6f0d005933b00p-7")],
           [fh_("0x1.0a4941f00af94p+1"), fh_("-0x1.6f0d005933b00p-7")], tset="textbook", reltol=fh_("0x1.3f0fbaf4e45fep-12"),
           abstol=fh_("0x1.4adc292f2891ap-30"), maxstep=fh_("0x1.0bb84ef0642cfp-1"), initstep=fh_("-0x1.58b45d6392d05p-4"))
    ode23s("ode23s_zero_span", "Lorenz", LOR, [1.0, 2.0, 20.0], ["list", 1.0, 1.0], [1.0, 1.0])
    # n = 4 (nalgebra's 4x4 cofactor inverse) and n = 6 (LU-based try_inverse): coupled Van der Pol oscillators
    Y4, Y6 = [2.0, 0.0, -1.5, 0.5], [2.0, 0.0, -1.5, 0.5, 0.3, -1.0]
    for tset in ("reference", "textbook"):
        ode23s(f"ode23s_{tset}_cvdp2_mu5", "Cvdp2", [5.0, 0.7], Y4, ["list", 0.0, 10.0], [0.0, 10.0], tset=tset, max_steps=20000)
        ode23s(f"ode23s_{tset}_cvdp3_mu5", "Cvdp3", [5.0, 0.7], Y6, ["list", 0.0, 10.0], [0.0, 10.0], tset=tset, max_steps=20000)
    ode23s("ode23s_textbook_cvdp2_stiff_dense", "Cvdp2", [200.0, 1.5], Y4, ["linspace", 0.0, 100.0, 51], linspace(0.0, 100.0, 51), tset="textbook",
           reltol=1e-6, abstol=1e-8)
    ode23s("ode23s_textbook_cvdp3_stiff_dense", "Cvdp3", [200.0, 1.5], Y6, ["linspace", 0.0, 100.0, 51], linspace(0.0, 100.0, 51), tset="textbook",
           reltol=1e-6, abstol=1e-8)
    ode23s("ode23s_cvdp2_singular_w", "Cvdp2", [0.0, 0.0], [0.0, 0.0, 0.0, 0.0], ["list", 0.0, 1.0], [0.0, 1.0], tset="textbook", max_steps=200)
    # n = 12 and n = 24 (matrices in thread-local memory on the device): the same LU-based try_inverse on larger systems
    Y12 = [2.0, 0.0, -1.5, 0.5, 0.3, -1.0, 1.1, 0.2, -0.4, 0.9, 1.7, -0.6]
    Y24 = Y12 + [0.25 * ((7 * i) % 11) - 1.25 for i in range(12)]
    for tset in ("reference", "textbook"):
        ode23s(f"ode23s_{tset}_cvdp6_mu5", "Cvdp6", [5.0, 0.7], Y12, ["list", 0.0, 4.0], [0.0, 4.0], tset=tset, max_steps=3000)
        ode23s(f"ode23s_{tset}_cvdp12_mu3", "Cvdp12", [3.0, 0.4], Y24, ["list", 0.0, 2.0], [0.0, 2.0], tset=tset, max_steps=2000)
    for which in ("kr4", "s4"):
        rosen(f"{which}_vdp", which, "VanDerPol", [5.0], [2.0, 0.0], ["linspace", 0.0, 0.39, 40], linspace(0.0, 0.39, 40))
        rosen(f"{which}_lorenz", which, "Lorenz", LOR, [0.1, 0.0, 0.0], ["linspace", 0.0, 0.05, 51], linspace(0.0, 0.05, 51))
        rosen(f"{which}_pendulum", which, "Pendulum", PEN, [1.0, 0.5], ["list", 0.0, 0.1, 0.15, 0.4], [0.0, 0.1, 0.15, 0.4])
        rosen(f"{which}_invalid_matrix_step2", which, "Lorenz", [10.0, 0.0, -4.0], [0.0, 0.0, 0.0],
              ["list", 0.0, 0.25, 0.25 + 0.25 / ros[which]["gamma"], 9.0], [0.0, 0.25, 0.25 + 0.25 / ros[which]["gamma"], 9.0])
        rosen(f"{which}_single_point", which, "Lorenz", LOR, [0.1, 0.0, 0.0], ["list", 0.0], [0.0])
        rosen(f"{which}_cvdp2", which, "Cvdp2", [5.0, 0.7], Y4, ["linspace", 0.0, 0.2, 21], linspace(0.0, 0.2, 21))
        rosen(f"{which}_cvdp3", which, "Cvdp3", [5.0, 0.7], Y6, ["linspace", 0.0, 0.2, 21], linspace(0.0, 0.2, 21))
        rosen(f"{which}_cvdp6", which, "Cvdp6", [5.0, 0.7], Y12, ["linspace", 0.0, 0.1, 11], linspace(0.0, 0.1, 11))
    json.dump({"rosenbrock_coeffs": {k: dict(gamma=hx(v["gamma"]), a=hx(v["a"]), b=hx(v["b"]), c=hx(v["c"])) for k, v in ros.items()},
               "cases": cases}, open(os.path.join(OUT, "stiff_cases.json"), "w"), indent=0)
    print(f"wrote {len(cases)} stiff cases to {OUT}")


if __name__ == "__main__":
    if "stiff" in sys.argv[1:]:
        main_stiff()       # leaves tableaus.json / solver_cases.json untouched
    else:
        main()
        main_stiff()
