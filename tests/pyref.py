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


def fdiv(a, b):
    """IEEE-754 division (Python raises on a zero divisor)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


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


def ring(p, n):
    """The oracle's RingF test system (dimension n): d_i = kappa (y_{i+1} - y_i) + r (y_i (1 - y_{i-1})), indices mod n."""
    kappa, r = p
    return lambda t, v: [kappa * (v[(i + 1) % n] - v[i]) + r * (v[i] * (1.0 - v[(i + n - 1) % n])) for i in range(n)]


def coupled_vdp(p, m):
    """The oracle's CoupledVdpF test system (dimension 2 m): m Van der Pol oscillators in a ring, state (x_0, y_0, x_1, ...),
    x_i' = y_i, y_i' = mu (1 - x_i^2) y_i - x_i + kappa (x_{i+1} - x_i)."""
    mu, kappa = p

    def f(t, v):
        d = []
        for i in range(m):
            x, y, xn = v[2 * i], v[2 * i + 1], v[2 * ((i + 1) % m)]
            d.append(y)
            d.append(mu * (1.0 - x * x) * y - x + kappa * (xn - x))
        return d
    return f


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


# ---------------------------------------------------------------------------------------------------------------------
# Stiff path: ode23s (problem.rs:409-557), oderosenbrock (problem.rs:560-650), fdjacobian (:903-926),
# coefficient sets rosenbrock.rs:14-118.  The dense linear algebra is nalgebra 0.20 (Cargo.toml: nalgebra = "0.20"; not
# vendored, no Cargo.lock), restated here from its published algorithms:
#   * `&A * s`, `&A - &B`, `&a + &b`: elementwise;
#   * `&M * &v` (gemm -> gemv -> axcpy): y = M[:,0]*v0, then y = M[:,j]*v_j + y for j = 1.. (column order);
#   * `M.lu()` (linalg/lu.rs): partial pivoting on the first largest |.| (icamax), multipliers = a_ji * (1/pivot),
#     trailing update a_jk = (-p_k)*l_j + a_jk, columns with a zero pivot skipped; `lu_internal()` is the packed L\U
#     matrix WITHOUT the permutation;
#   * `M.try_inverse()` (linalg/inverse.rs): closed-form adjugate formulas for n = 1, 2, 3; the 4x4 cofactor routine
#     (`do_inverse4`) for n = 4; `lu::try_invert_to` for n >= 5.
# ---------------------------------------------------------------------------------------------------------------------
SQRT2 = math.sqrt(2.0)


def na_lu_packed(m):
    n = len(m)
    m = [list(r) for r in m]
    for i in range(n):
        piv, best = i, abs(m[i][i])
        for r in range(i + 1, n):
            if abs(m[r][i]) > best:
                best, piv = abs(m[r][i]), r
        diag = m[piv][i]
        if diag == 0.0:
            continue
        if piv != i:
            m[i], m[piv] = m[piv], m[i]          # columns < i, column i and columns > i are all swapped
        inv = 1.0 / diag
        for r in range(i + 1, n):
            m[r][i] = m[r][i] * inv
        for k in range(i + 1, n):
            a = -m[i][k]
            for r in range(i + 1, n):
                m[r][k] = a * m[r][i] + m[r][k]
    return m


def na_try_inverse(m):
    n = len(m)
    if n == 1:
        d = m[0][0]
        return None if d == 0.0 else [[1.0 / d]]
    if n == 2:
        m11, m12, m21, m22 = m[0][0], m[0][1], m[1][0], m[1][1]
        det = m11 * m22 - m21 * m12
        if det == 0.0:
            return None
        return [[m22 / det, -m12 / det], [-m21 / det, m11 / det]]
    if n == 3:
        (m11, m12, m13), (m21, m22, m23), (m31, m32, m33) = m
        minor_m12_m23 = m22 * m33 - m32 * m23
        minor_m11_m23 = m21 * m33 - m31 * m23
        minor_m11_m22 = m21 * m32 - m31 * m22
        det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
        if det == 0.0:
            return None
        return [[minor_m12_m23 / det, (m13 * m32 - m33 * m12) / det, (m12 * m23 - m22 * m13) / det],
                [-minor_m11_m23 / det, (m11 * m33 - m31 * m13) / det, (m13 * m21 - m23 * m11) / det],
                [minor_m11_m22 / det, (m12 * m31 - m32 * m11) / det, (m11 * m22 - m21 * m12) / det]]
    if n == 4:
        return _na_inverse4(m)
    return _na_lu_invert(m)


def _na_inverse4(mat):
    """`do_inverse4` (nalgebra 0.20 linalg/inverse.rs): the cofactor expansion of MESA's gluInvertMatrix on the column-major
    slice m[k] = M[k % 4][k / 4]; products left to right, sums left to right; det from the first row of cofactors; the
    cofactors are then MULTIPLIED by 1/det (not divided, unlike the n <= 3 formulas)."""
    m = [mat[k % 4][k // 4] for k in range(16)]
    o = [0.0] * 16
    o[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10]
    o[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10]
    o[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6]
    o[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6]
    o[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10]
    o[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10]
    o[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6]
    o[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6]
    o[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9]
    o[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9]
    o[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5]
    o[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5]
    o[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9]
    o[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9]
    o[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5]
    o[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5]
    det = m[0] * o[0] + m[1] * o[4] + m[2] * o[8] + m[3] * o[12]
    if det == 0.0:
        return None
    inv_det = 1.0 / det
    return [[o[c * 4 + r] * inv_det for c in range(4)] for r in range(4)]


def _na_lu_invert(mat):
    """`lu::try_invert_to` (nalgebra 0.20 linalg/lu.rs), what try_inverse does for n >= 5: Gaussian elimination with partial
    pivoting on a copy (first largest |.|, multipliers a_ji * (1/pivot), update a_jk = (-p_k) l_j + a_jk; a zero pivot fails)
    with the row swaps applied to an identity `out`; then out <- L^-1 out (unit diagonal, forward, column by column) and
    out <- U^-1 out (backward: coeff = b_i / u_ii, b_j = (-coeff) u_ji + b_j for j < i)."""
    n = len(mat)
    m = [list(r) for r in mat]
    out = [[1.0 if r == c else 0.0 for c in range(n)] for r in range(n)]
    for i in range(n):
        piv, best = i, abs(m[i][i])
        for r in range(i + 1, n):
            if abs(m[r][i]) > best:
                best, piv = abs(m[r][i]), r
        diag = m[piv][i]
        if diag == 0.0:
            return None
        if piv != i:
            out[i], out[piv] = out[piv], out[i]
            m[i], m[piv] = m[piv], m[i]
        inv = 1.0 / diag
        for r in range(i + 1, n):
            m[r][i] = m[r][i] * inv
        for k in range(i + 1, n):
            a = -m[i][k]
            for r in range(i + 1, n):
                m[r][k] = a * m[r][i] + m[r][k]
    for k in range(n):          # solve_lower_triangular_with_diag_mut(out, 1.0)
        for i in range(n - 1):
            coeff = out[i][k] / 1.0
            for r in range(i + 1, n):
                out[r][k] = (-coeff) * m[r][i] + out[r][k]
    for k in range(n):          # solve_upper_triangular_mut(out)
        for i in range(n - 1, -1, -1):
            diag = m[i][i]
            if diag == 0.0:
                return None
            coeff = out[i][k] / diag
            out[i][k] = coeff
            for r in range(i):
                out[r][k] = (-coeff) * m[r][i] + out[r][k]
    return out


def na_gemv(m, v):
    n = len(v)
    y = [m[r][0] * v[0] for r in range(n)]
    for j in range(1, n):
        for r in range(n):
            y[r] = m[r][j] * v[j] + y[r]
    return y


def fdjacobian(f, t, x):
    """problem.rs:903-926"""
    ftx = f(t, x)
    n = len(ftx)
    J = [[0.0] * n for _ in range(n)]
    for c in range(n):
        xj = x[c]
        if xj == 0.0:
            xj += 1.0
        dxj = xj * 0.01
        tmp = list(x)
        tmp[c] += dxj
        yj = f(t, tmp)
        for r in range(n):
            J[r][c] = fdiv(yj[r] - ftx[r], dxj)
    return J


def pnorm2(v, libm_folded=False):
    s = 0.0
    for x in v:
        s = s + abs(x) * abs(x)
    try:
        return math.sqrt(s) if libm_folded else math.pow(s, 1.0 * (1.0 / 2.0))
    except OverflowError:
        return math.inf


def ode23s(f, y0, tspan, reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0, points_all=True,
           max_steps=0, libm_folded=False, true_inverse=False):
    """problem.rs:409-557.  status 0 done, 1 loop left because |h| <= minstep before tfinal, 2 max_steps guard,
    3 InvalidMatrix.  true_inverse=True is the TEXTBOOK variant (W^-1 instead of the inverse of the packed LU)."""
    if not tspan:
        return dict(tout=[], yout=[], accepted=0, rejected=0, status=0, t=0.0, y=[])
    t = tspan[0]
    tfinal = tspan[-1]
    minstep = abs(tfinal - t) / 1e18 if minstep is None else minstep
    maxstep = abs(tfinal - t) / 2.5 if maxstep is None else maxstep
    d = 1.0 / (2.0 + SQRT2)
    e32 = 6.0 + SQRT2
    if initstep == 0.0:
        h0, tdir, f0 = hinit(f, y0, t, tfinal, 3, reltol, abstol, libm_folded)
    else:
        h0, tdir, f0 = initstep, signum(tfinal - t), f(t, y0)
    h = tdir * fmin(abs(h0), maxstep)
    tout = [t]
    yout = [list(y0)]
    jac = fdjacobian(f, t, y0)
    n = len(y0)
    y = list(y0)
    f0 = list(f0)
    acc = rej = attempts = 0
    status = 0
    while abs(t - tfinal) > 0.0 and minstep < abs(h):
        if max_steps and attempts >= max_steps:
            status = 2
            break
        attempts += 1
        if abs(t - tfinal) < abs(h):
            h = tfinal - t
        hd = h * d
        w = [[(1.0 if r == c else 0.0) - jac[r][c] * (1.0 * hd) for c in range(n)] for r in range(n)]
        if n * n != 1 and not true_inverse:
            w = na_lu_packed(w)
        fdt = f(t + h / 100.0, y)
        for i in range(n):
            fdt[i] = (fdt[i] - f0[i]) * (hd / (h / 100.0))
        winv = na_try_inverse(w)
        if winv is None:
            status = 3
            break
        k1 = na_gemv(winv, [f0[i] + fdt[i] for i in range(n)])
        f1y = [y[i] + k1[i] * 0.5 * h for i in range(n)]
        f1 = f(t + 0.5 * h, f1y)
        k2 = na_gemv(winv, [f1[i] - k1[i] for i in range(n)])
        k2 = [k2[i] + k1[i] for i in range(n)]
        ynew = [y[i] + k2[i] * h for i in range(n)]
        f2 = f(t + h, ynew)
        k3 = na_gemv(winv, [((f2[i] - (k2[i] - f1[i]) * (1.0 * e32)) - (k1[i] - f0[i]) * (1.0 * 2.0)) + fdt[i] for i in range(n)])
        kerr = [(k1[i] - k2[i] * (1.0 * 2.0)) + k3[i] for i in range(n)]
        err = pnorm2(kerr, libm_folded) * (abs(h) / 6.0)
        delta = fmax(fmax(pnorm2(y, libm_folded), pnorm2(ynew, libm_folded)) * reltol, 1.0 * abstol)
        if err <= delta:
            acc += 1
            for toi in tspan:
                if toi > t and toi <= t + h:
                    s = fdiv(toi - t, h)
                    tout.append(toi)
                    c1 = 1.0 * (s * (1.0 - s) / (1.0 - 2.0 * d))
                    c2 = 1.0 * (s * (s - 2.0 * d) / (1.0 - 2.0 * d))
                    yout.append([y[i] + (k1[i] * c1 + k2[i] * c2) * h for i in range(n)])
            if points_all and abs(tout[-1] - (t + h)) > EPS:
                tout.append(t + h)
                yout.append(list(ynew))
            t += h
            y = ynew
            f0 = f2
            jac = fdjacobian(f, t, y)
        else:
            rej += 1
        r = delta / err if err != 0.0 else (math.inf if delta > 0.0 else math.nan)
        try:
            g = math.pow(r, 1.0 / 3.0)
        except (OverflowError, ValueError):
            g = math.nan if r != r else math.inf
        h = fmin(maxstep, g * abs(h) * 0.8) * tdir
    else:
        status = 0 if not (abs(t - tfinal) > 0.0) else 1
    return dict(tout=tout, yout=yout, accepted=acc, rejected=rej, status=status, t=t, y=y)


ROSENBROCK = {
    # rosenbrock.rs:16-68 (Matrix4::new is row-major: a[i][j] below is a[(i, j)])
    "kr4": dict(gamma=0.231,
                a=[[0., 0., 0., 0.], [2., 0., 0., 0.], [4.452_470_820_736, 4.163_528_788_60, 0., 0.],
                   [4.452_470_820_736, 4.163_528_788_60, 0., 0.]],
                b=[3.957_503_746_63, 4.624_892_388_36, 0.617_477_263_873, 1.282_612_945_6],
                c=[[0., 0., 0., 0.], [-5.071_675_338_77, 0., 0., 0.], [6.020_152_728_65, 0.159_750_068_467_3, 0., 0.],
                   [-1.856_343_618_677, -8.505_380_858_19, -2.084_075_136_02, 0.0]]),
    # rosenbrock.rs:71-117
    "s4": dict(gamma=0.5,
               a=[[0., 0., 0., 0.], [2., 0., 0., 0.], [48. / 25., 6. / 250., 0., 0.], [48. / 25., 6. / 250., 0., 0.]],
               b=[19. / 9., 1. / 2., 25. / 108., 125. / 108.],
               c=[[0., 0., 0., 0.], [-8., 0., 0., 0.], [372. / 25., 12. / 5., 0., 0.],
                  [-112. / 125., -54. / 125., -2. / 5., 0.]]),
}


def oderosenbrock(f, co, y0, tspan):
    """problem.rs:560-640.  Returns dict(tout = diff(tspan) (sic), yout, status 0 | 3 InvalidMatrix)."""
    if not tspan:
        return dict(tout=[], yout=[], status=0)
    hvec = [tspan[i + 1] - tspan[i] for i in range(len(tspan) - 1)]
    n = len(y0)
    x = [list(y0)]
    for solstep, hs in enumerate(hvec):
        ts = tspan[solstep]
        xs = list(x[solstep])
        dfdx = fdjacobian(f, ts, xs)
        dg = 1.0 * fdiv(1.0, co["gamma"] * hs)
        jac = [[(dg if r == c else 0.0) - dfdx[r][c] for c in range(n)] for r in range(n)]
        jinv = na_try_inverse(jac)
        if jinv is None:
            return dict(tout=hvec, yout=x, status=3)
        g = [na_gemv(jinv, f(ts + co["b"][0] * hs, xs))]
        nx = list(x[-1])
        for i in range(n):
            nx[i] += g[0][i] * co["b"][0]
        for i in range(1, 4):
            dx = [0.0] * n
            df = [0.0] * n
            for j in range(min(len(g), i - 1)):
                for dd in range(n):
                    dx[dd] += g[j][dd] * co["a"][i][j]
                    df[dd] += g[j][dd] * co["c"][i][j]
            fx = f(ts + co["b"][i] * hs, [xs[dd] + dx[dd] for dd in range(n)])
            gv = na_gemv(jinv, fx)
            gv = [gv[dd] + df[dd] * fdiv(1.0, hs) for dd in range(n)]
            for dd in range(n):
                nx[dd] += gv[dd] * co["b"][i]
            g.append(gv)
        x.append(nx)
    return dict(tout=hvec, yout=x, status=0)
