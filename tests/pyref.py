"""pyref.py — a second, independent restatement of the reference's RK hot path in pure Python floats.

Purpose: pin the C++ oracle (oracle/diffeq_oracle.cpp).  The reference's own tests assert no numeric solver result
 and Rust cannot be built here, so "parity unpinned by the reference" is mitigated by requiring two
restatements, written separately from the Rust text (one in C++ with templates and fixed arrays, this one with Python
lists, its tableaus parsed straight out of runge_kutta.rs by tools/gen_golden.py), to agree BIT-FOR-BIT on every
golden case (tests/test_oracle_cpu.py).  Python floats are IEEE f64, CPython never fuses a*b+c, and math.pow /
math.log10 call the same glibc the Rust crate would.

Follows /root/reference/src/ode/problem.rs:193-406 (loops), :702-899 (helpers), types.rs:91-116 (pnorm).
"""
import ctypes
import math

_libm = ctypes.CDLL("libm.so.6")
_libm.exp10.restype = ctypes.c_double
_libm.exp10.argtypes = [ctypes.c_double]

EPS = 2.220446049250313e-16


def signum(x):
    return x if x != x else math.copysign(1.0, x)


def fmax(a, b):  # f64::max (NaN-ignoring)
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def fmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def lorenz(p):
    s, r, b = p
    return lambda t, v: [s * (v[1] - v[0]), v[0] * (r - v[2]) - v[1], v[0] * v[1] - b * v[2]]


def vanderpol(p):
    mu = p[0]
    return lambda t, v: [v[1], mu * (1.0 - v[0] * v[0]) * v[1] - v[0]]


def pendulum(p):
    g, a, w = p
    return lambda t, v: [v[1], -g * v[1] - math.sin(v[0]) + a * math.cos(w * t)]


def is_fsal(tb):
    """runge_kutta.rs:185-193 (b.as_slice() is column-major: the first S entries are column 0)."""
    S = len(tb["c"])
    for col in range(S):
        if tb["a"][S - 1][col] != tb["b"][col][0]:
            return False
    return tb["c"][S - 1] == 1.0


def calc_coefficients(f, tb, t, k0, y, dt):
    """problem.rs:702-737"""
    S = len(tb["c"])
    ks = [k0]
    for row in range(1, S):
        yi = list(y)
        for col, k in enumerate(ks):
            for d in range(len(yi)):
                yi[d] += k[d] * (tb["a"][row][col] * dt)
        tn = t + tb["c"][row] * dt
        ks.append(f(tn, yi))
    return ks


def oderk_fixed(f, tb, y0, tspan):
    """problem.rs:361-406"""
    ys = [list(y0)]
    b = [row[0] for row in tb["b"]]
    for i in range(len(tspan) - 1):
        dt = tspan[i + 1] - tspan[i]
        yi = list(ys[i])
        ks = calc_coefficients(f, tb, tspan[i], f(tspan[i], yi), list(yi), dt)
        for s, k in enumerate(ks):
            for d in range(len(yi)):
                yi[d] += k[d] * b[s] * dt
        ys.append(yi)
    return list(tspan), ys


def hinit(f, x0, t0, tend, order, reltol, abstol, libm_folded=False):
    """problem.rs:850-899"""
    tdir = signum(tend - t0)
    norm = 0.0
    for v in x0:
        if abs(v) > norm:
            norm = abs(v)
    tau = fmax(norm * reltol, 1.0 * abstol)
    d0 = norm / tau
    f0 = f(t0, x0)
    n1 = 0.0
    for v in f0:
        if abs(v) > n1:
            n1 = abs(v)
    d1 = n1 / tau
    h0 = 1.0e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
    x1 = [x0[d] + f0[d] * h0 * tdir for d in range(len(x0))]
    f1 = f(t0 + tdir * h0, x1)
    n2 = 0.0
    for d in range(len(f1)):
        v = abs(f1[d] - f0[d])
        if v > n2:
            n2 = v
    d2 = n2 / (tau * h0)
    m = fmax(d1, d2)
    if m < 1e-15:
        h1 = fmax(1.0e-6, 1.0e-3 * h0)
    else:
        pw = -(2.0 + math.log10(m)) / float(order + 1)
        h1 = _libm.exp10(pw) if libm_folded else math.pow(10.0, pw)
    h = tdir * fmin(fmin(h1, 100.0 * h0), tdir * (tend - t0))
    return h, tdir, f0


def stepsize_hw92(dt, tdir, x0, xtrial, xerr, order, timeout, abstol, reltol, maxstep, libm_folded=False):
    """problem.rs:801-845 with pnorm P(2) (types.rs:109-115)"""
    xerr = list(xerr)
    for d in range(len(x0)):
        if xtrial[d] != xtrial[d]:
            return 10.0, 0.2 * dt, 5
        xerr[d] /= fmax(abs(x0[d]), abs(xtrial[d])) * reltol + abstol
    s = 0.0
    for v in xerr:
        s = s + abs(v) * abs(v)
    try:
        err = math.sqrt(s) if libm_folded else math.pow(s, 1.0 * (1.0 / 2.0))   # LLVM: pow(x, 0.5) -> sqrt(x)
    except OverflowError:
        err = math.inf
    pw = 1.0 / float(order + 1)
    try:
        g = math.pow(1.0 / err, pw) if err != 0.0 else math.inf
    except (OverflowError, ZeroDivisionError):
        g = math.inf
    new_dt = fmin(maxstep, fmax(0.2, g * 0.8) * tdir * dt)
    if timeout > 0:
        new_dt = fmin(new_dt, dt)
        timeout -= 1
    return err, tdir * new_dt, timeout


def hermite_interp(tq, t, dt, y0, y1, f0, f1):
    """problem.rs:783-798"""
    theta = (tq - t) / dt
    out = []
    for i in range(len(y0)):
        out.append((y0[i] * (1.0 - theta) + y1[i] * theta)
                   + ((y1[i] - y0[i]) * (1.0 - 2.0 * theta) + f0[i] * (theta - 1.0) * dt + f1[i] * theta * dt)
                   * theta * (theta - 1.0))
    return out


def oderk_adapt(f, tb, y0, tspan, reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0, points_all=True,
                max_steps=0, libm_folded=False):
    """problem.rs:193-358.  Returns dict(tout, yout, accepted, rejected, status) — status 0 done, 1 minstep break,
    2 max_steps guard."""
    if not tspan:
        return dict(tout=[], yout=[], accepted=0, rejected=0, status=0)
    order = tb["order_min"]
    fsal = is_fsal(tb)
    S = len(tb["c"])
    t = tspan[0]
    tend = tspan[-1]
    minstep = abs(tend - t) / 1e18 if minstep is None else minstep
    maxstep = abs(tend - t) / 2.5 if maxstep is None else maxstep
    h, tdir, f0 = hinit(f, y0, t, tend, order, reltol, abstol, libm_folded)
    if initstep != 0.0:
        if abs(signum(initstep) - tdir) < EPS:
            dt = initstep
        else:
            raise ValueError("InvalidInitstep")
    else:
        dt = h
    timeout = 0
    last_step = abs(t + dt - tend) <= EPS
    tout = [t]
    ys = [list(y0)]
    ck, cy = list(f0), list(y0)
    it = 1
    acc = rej = 0
    status = 0
    attempts = 0
    while True:
        if max_steps and attempts >= max_steps:
            status = 2
            break
        attempts += 1
        ks = calc_coefficients(f, tb, t, ck, cy, dt)
        y = list(ys[-1])
        ytrial = [0.0] * len(y)
        yerr = [0.0] * len(y)
        for s, k in enumerate(ks):
            for d in range(len(y)):
                ytrial[d] += k[d] * tb["b"][s][0]
                yerr[d] += k[d] * tb["b"][s][1]
        for d in range(len(y)):
            yerr[d] = (ytrial[d] - yerr[d]) * dt
            ytrial[d] = y[d] + (ytrial[d] * dt)
        err, ndt, timeout = stepsize_hw92(dt, tdir, y, ytrial, yerr, order, timeout, abstol, reltol, maxstep, libm_folded)
        if err < 1.0:
            acc += 1
            f1 = list(ks[S - 1]) if fsal else f(t + dt, ytrial)
            if not points_all:
                while it < len(tspan) and (tdir * tspan[it] < tdir * (t + dt) or last_step):
                    ys.append(hermite_interp(tspan[it], t, dt, y, ytrial, ks[0], f1))
                    tout.append(tspan[it])
                    it += 1
            else:
                while it < len(tspan) and tdir * t < tdir * tspan[it] and tdir * tspan[it] < tdir * (t + dt):
                    ys.append(hermite_interp(tspan[it], t, dt, y, ytrial, ks[0], f1))
                    tout.append(tspan[it])
                    it += 1
                ys.append(list(ytrial))
                tout.append(t + dt)
            ck, cy = f1, list(ytrial)
            if last_step:
                break
            t += dt
            dt = ndt
            if tdir * (t + dt + dt / 100.0) >= tdir * tend:
                dt = tend - t
                last_step = True
        elif abs(ndt) < minstep:
            status = 1
            break
        else:
            rej += 1
            last_step = False
            dt = ndt
            timeout = 5
    return dict(tout=tout, yout=ys, accepted=acc, rejected=rej, status=status)
