"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (diffeq_b200) never imports this module.

The oracle restates /root/reference/src/ode/problem.rs:193-406,702-899 and runge_kutta.rs:231-817
(see the header of diffeq_oracle.cpp).  Parity is UNPINNED by the reference (it has no numeric tests and
no Rust toolchain exists here); the oracle is pinned by tests/pyref.py + tests/golden/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# `Ode` enum order (mod.rs:14-26) + ODE21 extension
FEULER, HEUN, MIDPOINT, ODE23, ODE23S, ODE4, ODE45, ODE45FE, ODE4SKR, ODE4SS, ODE78, ODE21 = range(12)
LORENZ, VANDERPOL, PENDULUM = 0, 1, 2
CVDP2, CVDP3 = 7, 8   # 2 / 3 Van der Pol oscillators in a ring (dimension 4 / 6): the stiff path beyond the closed-form inverses
CVDP6, CVDP12 = 9, 10  # 6 / 12 oscillators (dimension 12 / 24): the stiff path with its matrices in thread-local memory
RING8, RING16, RING32, RING64 = 3, 4, 5, 6   # test systems of dimension 8 .. 64 (ring of coupled logistic cells) for the generic-dimension path
REFERENCE, TEXTBOOK = 0, 1
POINTS_ALL, POINTS_SPECIFIED = 0, 1
ST_DONE, ST_MINSTEP_BREAK, ST_MAX_STEPS, ST_INVALID_MATRIX = 0, 1, 2, 3
KR4, S4 = 0, 1   # RosenbrockCoeffs::kr4 / s4 (rosenbrock.rs:16,71) = Ode::Ode4skr / Ode::Ode4ss


class Opts(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("has_minstep", C.c_int), ("minstep", C.c_double),
                ("has_maxstep", C.c_int), ("maxstep", C.c_double), ("initstep", C.c_double), ("points", C.c_int),
                ("max_steps", C.c_uint64), ("libm_folded", C.c_int)]


def opts(reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0, points=POINTS_ALL, max_steps=0, libm_folded=0):
    """Defaults follow options.rs:283-299 (Reltol 1e-5, Abstol 1e-8, Initstep 0, Points::All).  libm_folded=1 selects what
    an LLVM-optimised build evaluates (sqrt / exp10 instead of pow(.,0.5) / pow(10,.)); see diffeq_oracle.cpp."""
    return Opts(reltol, abstol, int(minstep is not None), minstep or 0.0, int(maxstep is not None), maxstep or 0.0,
                initstep, points, max_steps, libm_folded)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "diffeq_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.oracle_pow.restype = C.c_double
        _LIB.oracle_pow.argtypes = [C.c_double, C.c_double]
        _LIB.oracle_log10.restype = C.c_double
        _LIB.oracle_log10.argtypes = [C.c_double]
        for fn in ("oracle_exp10", "oracle_sin", "oracle_cos"):
            getattr(_LIB, fn).restype = C.c_double
            getattr(_LIB, fn).argtypes = [C.c_double]
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def rhs_dim(rhs):
    return lib().oracle_rhs_dim(rhs)


def rhs_nparams(rhs):
    return lib().oracle_rhs_nparams(rhs)


def linspace(a, b, n):
    out = np.empty(n, dtype=np.float64)
    lib().oracle_linspace(C.c_double(a), C.c_double(b), C.c_size_t(n), _p(out))
    return out


def tableau(method, tset=REFERENCE):
    S = C.c_int(); ad = C.c_int(); om = C.c_int(); fs = C.c_int()
    a = np.zeros(169); b = np.zeros(26); c = np.zeros(13)
    rc = lib().oracle_tableau(method, tset, C.byref(S), _p(a), _p(b), _p(c), C.byref(ad), C.byref(om), C.byref(fs))
    if rc:
        raise ValueError(f"oracle_tableau rc={rc}")
    s = S.value
    return dict(S=s, a=a[:s * s].reshape(s, s).copy(), b=b[:2 * s].reshape(s, 2).copy(), c=c[:s].copy(),
                adaptive=bool(ad.value), order_min=om.value, fsal=bool(fs.value))


def solve_fixed(method, rhs, params, y0, tspan, want_all=True):
    """Single trajectory, oderk_fixed.  Returns (tout, yout[nt][n]) like OdeSolution."""
    params, y0, tspan = _f64(params), _f64(y0), _f64(tspan)
    n = rhs_dim(rhs)
    yout = np.empty((len(tspan), n)) if want_all else None
    yfin = np.empty(n)
    nfe = C.c_uint32()
    rc = lib().oracle_solve_fixed(method, rhs, _p(params), _p(y0), _p(tspan), C.c_size_t(len(tspan)), _p(yout), _p(yfin),
                                  C.byref(nfe))
    if rc:
        raise RuntimeError(f"oracle_solve_fixed rc={rc}")
    return dict(tout=tspan, yout=yout, yfinal=yfin, nfe=nfe.value)


def solve_adaptive(method, rhs, params, y0, tspan, o=None, tset=REFERENCE, want_all=True):
    """Single trajectory, oderk_adapt.  Returns the ragged (tout, yout) of Points::All/Specified + counters."""
    o = o or opts()
    params, y0, tspan = _f64(params), _f64(y0), _f64(tspan)
    n = rhs_dim(rhs)
    L = lib()
    nout = C.c_size_t(); cnt = (C.c_uint32 * 3)(); st = C.c_uint8(); tr = C.c_double()
    yfin = np.empty(n)
    args = [method, tset, rhs, _p(params), _p(y0), _p(tspan), C.c_size_t(len(tspan)), C.byref(o)]
    rc = L.oracle_solve_adaptive(*args, None, None, C.c_size_t(0), C.byref(nout), _p(yfin), cnt, C.byref(st), C.byref(tr))
    if rc:
        return dict(rc=rc)
    res = dict(rc=0, nout=nout.value, yfinal=yfin, accepted=cnt[0], rejected=cnt[1], nfe=cnt[2], status=st.value,
               t_reached=tr.value)
    if want_all:
        tout = np.empty(nout.value); yout = np.empty((nout.value, n))
        L.oracle_solve_adaptive(*args, _p(tout), _p(yout), C.c_size_t(nout.value), C.byref(nout), _p(yfin), cnt,
                                C.byref(st), C.byref(tr))
        res.update(tout=tout, yout=yout)
    return res


def _ens_params(params, ntraj, np_):
    params = _f64(params)
    per = int(params.ndim == 2)
    if per:
        assert params.shape == (ntraj, np_)
    else:
        assert params.shape == (np_,)
    return params, per


def ensemble_adaptive(method, rhs, params, y0, tspan, o=None, tset=REFERENCE, nthreads=0):
    o = o or opts()
    y0, tspan = _f64(y0), _f64(tspan)
    ntraj, n = y0.shape
    params, per = _ens_params(params, ntraj, rhs_nparams(rhs))
    yf = np.empty((ntraj, n)); acc = np.empty(ntraj, np.uint32); rej = np.empty(ntraj, np.uint32)
    nfe = np.empty(ntraj, np.uint32); st = np.empty(ntraj, np.uint8); tr = np.empty(ntraj)
    rc = lib().oracle_ensemble_adaptive(method, tset, rhs, C.c_size_t(ntraj), _p(y0), _p(params), per, _p(tspan),
                                        C.c_size_t(len(tspan)), C.byref(o), nthreads, _p(yf), _p(acc, C.c_uint32),
                                        _p(rej, C.c_uint32), _p(nfe, C.c_uint32), _p(st, C.c_uint8), _p(tr))
    if rc:
        raise RuntimeError(f"oracle_ensemble_adaptive rc={rc}")
    return dict(yfinal=yf, accepted=acc, rejected=rej, nfe=nfe, status=st, t_reached=tr)


def ensemble_adaptive_vec(method, rhs, params, y0, tspan, o=None, tset=REFERENCE, nthreads=0):
    """ensemble_adaptive with the reference's allocation pattern (heap vector per state, ~20 allocations per step attempt);
    same bits, used to report how the CPU baseline depends on it."""
    o = o or opts()
    y0, tspan = _f64(y0), _f64(tspan)
    ntraj, n = y0.shape
    params, per = _ens_params(params, ntraj, rhs_nparams(rhs))
    yf = np.empty((ntraj, n)); acc = np.empty(ntraj, np.uint32); rej = np.empty(ntraj, np.uint32)
    nfe = np.empty(ntraj, np.uint32); st = np.empty(ntraj, np.uint8)
    rc = lib().oracle_ensemble_adaptive_vec(method, tset, rhs, C.c_size_t(ntraj), _p(y0), _p(params), per, _p(tspan),
                                            C.c_size_t(len(tspan)), C.byref(o), nthreads, _p(yf), _p(acc, C.c_uint32),
                                            _p(rej, C.c_uint32), _p(nfe, C.c_uint32), _p(st, C.c_uint8))
    if rc:
        raise RuntimeError(f"oracle_ensemble_adaptive_vec rc={rc}")
    return dict(yfinal=yf, accepted=acc, rejected=rej, nfe=nfe, status=st)


def ensemble_dense(method, rhs, params, y0, tspan, o=None, tset=REFERENCE, nthreads=0):
    """Points::All filtered to the tspan times; out is SoA [n][nt][ntraj] (rows >= nvalid, except the last, are NaN)."""
    o = o or opts()
    y0, tspan = _f64(y0), _f64(tspan)
    ntraj, n = y0.shape
    params, per = _ens_params(params, ntraj, rhs_nparams(rhs))
    out = np.full((n, len(tspan), ntraj), np.nan)
    nv = np.empty(ntraj, np.uint32); acc = np.empty(ntraj, np.uint32); rej = np.empty(ntraj, np.uint32)
    st = np.empty(ntraj, np.uint8)
    rc = lib().oracle_ensemble_dense(method, tset, rhs, C.c_size_t(ntraj), _p(y0), _p(params), per, _p(tspan),
                                     C.c_size_t(len(tspan)), C.byref(o), nthreads, _p(out), _p(nv, C.c_uint32),
                                     _p(acc, C.c_uint32), _p(rej, C.c_uint32), _p(st, C.c_uint8))
    if rc:
        raise RuntimeError(f"oracle_ensemble_dense rc={rc}")
    return dict(out=out, nvalid=nv, accepted=acc, rejected=rej, status=st)


def ensemble_fixed(method, rhs, params, y0, tspan, nthreads=0):
    y0, tspan = _f64(y0), _f64(tspan)
    ntraj, n = y0.shape
    params, per = _ens_params(params, ntraj, rhs_nparams(rhs))
    yf = np.empty((ntraj, n))
    rc = lib().oracle_ensemble_fixed(method, rhs, C.c_size_t(ntraj), _p(y0), _p(params), per, _p(tspan),
                                     C.c_size_t(len(tspan)), nthreads, _p(yf))
    if rc:
        raise RuntimeError(f"oracle_ensemble_fixed rc={rc}")
    return yf


# ---- stiff path: ode23s problem.rs:409-557, oderosenbrock problem.rs:560-640 ----
def solve_ode23s(rhs, params, y0, tspan, o=None, tset=REFERENCE, want_all=True):
    """Single trajectory ode23s.  tset=TEXTBOOK inverts W itself instead of its packed LU factors."""
    o = o or opts()
    params, y0, tspan = _f64(params), _f64(y0), _f64(tspan)
    n = rhs_dim(rhs)
    L = lib()
    nout = C.c_size_t(); cnt = (C.c_uint32 * 3)(); st = C.c_uint8(); tr = C.c_double()
    yfin = np.empty(n)
    args = [tset, rhs, _p(params), _p(y0), _p(tspan), C.c_size_t(len(tspan)), C.byref(o)]
    rc = L.oracle_solve_ode23s(*args, None, None, C.c_size_t(0), C.byref(nout), _p(yfin), cnt, C.byref(st), C.byref(tr))
    if rc:
        return dict(rc=rc)
    res = dict(rc=0, nout=nout.value, yfinal=yfin, accepted=cnt[0], rejected=cnt[1], nfe=cnt[2], status=st.value,
               t_reached=tr.value)
    if want_all:
        tout = np.empty(nout.value); yout = np.empty((nout.value, n))
        L.oracle_solve_ode23s(*args, _p(tout), _p(yout), C.c_size_t(nout.value), C.byref(nout), _p(yfin), cnt,
                              C.byref(st), C.byref(tr))
        res.update(tout=tout, yout=yout)
    return res


def ensemble_ode23s(rhs, params, y0, tspan, o=None, tset=REFERENCE, nthreads=0, dense=False):
    o = o or opts()
    y0, tspan = _f64(y0), _f64(tspan)
    ntraj, n = y0.shape
    params, per = _ens_params(params, ntraj, rhs_nparams(rhs))
    yf = np.empty((ntraj, n)); acc = np.empty(ntraj, np.uint32); rej = np.empty(ntraj, np.uint32)
    nfe = np.empty(ntraj, np.uint32); st = np.empty(ntraj, np.uint8); tr = np.empty(ntraj)
    nv = np.empty(ntraj, np.uint32); nrows = np.empty(ntraj, np.uint32)
    out = np.full((n, len(tspan), ntraj), np.nan) if dense else None
    rc = lib().oracle_ensemble_ode23s(tset, rhs, C.c_size_t(ntraj), _p(y0), _p(params), per, _p(tspan),
                                      C.c_size_t(len(tspan)), C.byref(o), nthreads, _p(yf), _p(acc, C.c_uint32),
                                      _p(rej, C.c_uint32), _p(nfe, C.c_uint32), _p(st, C.c_uint8), _p(tr), _p(out),
                                      _p(nv, C.c_uint32), _p(nrows, C.c_uint32))
    if rc:
        raise RuntimeError(f"oracle_ensemble_ode23s rc={rc}")
    return dict(yfinal=yf, accepted=acc, rejected=rej, nfe=nfe, status=st, t_reached=tr, out=out, nvalid=nv, nrows=nrows)


def solve_rosenbrock(which, rhs, params, y0, tspan):
    """Single trajectory oderosenbrock.  Returns yout[rows][n] and tout = diff(tspan) as the reference does (sic)."""
    params, y0, tspan = _f64(params), _f64(y0), _f64(tspan)
    n = rhs_dim(rhs)
    yout = np.full((len(tspan), n), np.nan); yfin = np.empty(n)
    st = C.c_uint8(); nfe = C.c_uint32(); rows = C.c_size_t()
    rc = lib().oracle_solve_rosenbrock(which, rhs, _p(params), _p(y0), _p(tspan), C.c_size_t(len(tspan)), _p(yout),
                                       _p(yfin), C.byref(st), C.byref(nfe), C.byref(rows))
    if rc:
        raise RuntimeError(f"oracle_solve_rosenbrock rc={rc}")
    return dict(tout=np.diff(tspan), yout=yout[:rows.value], yfinal=yfin, status=st.value, nfe=nfe.value)


def ensemble_rosenbrock(which, rhs, params, y0, tspan, nthreads=0, want_all=False):
    y0, tspan = _f64(y0), _f64(tspan)
    ntraj, n = y0.shape
    params, per = _ens_params(params, ntraj, rhs_nparams(rhs))
    yf = np.empty((ntraj, n)); st = np.empty(ntraj, np.uint8)
    yall = np.full((ntraj, len(tspan), n), np.nan) if want_all else None
    rc = lib().oracle_ensemble_rosenbrock(which, rhs, C.c_size_t(ntraj), _p(y0), _p(params), per, _p(tspan),
                                          C.c_size_t(len(tspan)), nthreads, _p(yf), _p(st, C.c_uint8), _p(yall))
    if rc:
        raise RuntimeError(f"oracle_ensemble_rosenbrock rc={rc}")
    return dict(yfinal=yf, status=st, yall=yall)


def fdjacobian(rhs, params, t, x):
    """`fdjacobian` (problem.rs:903-926): forward-difference Jacobian, [n][n]."""
    params, x = _f64(params), _f64(x)
    n = rhs_dim(rhs)
    J = np.empty((n, n))
    rc = lib().oracle_fdjacobian(rhs, _p(params), C.c_double(t), _p(x), _p(J))
    if rc:
        raise RuntimeError(f"oracle_fdjacobian rc={rc}")
    return J


def rosenbrock_coeffs(which):
    g = C.c_double(); a = np.zeros(16); b = np.zeros(4); c = np.zeros(16)
    rc = lib().oracle_rosenbrock_coeffs(which, C.byref(g), _p(a), _p(b), _p(c))
    if rc:
        raise ValueError(f"oracle_rosenbrock_coeffs rc={rc}")
    return dict(gamma=g.value, a=a.reshape(4, 4), b=b, c=c.reshape(4, 4))


def num_threads():
    return lib().oracle_num_threads()
