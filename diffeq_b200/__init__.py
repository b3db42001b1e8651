"""diffeq_b200 — host-side mirror of mattsse/diffeq's problem/solver surface over libdiffeq_b200.so.

The reference API for this path (all /root/reference/src/ode/):
    OdeProblem::builder().fun(f).init(y0).tspan(ts) / .tspan_linspace(a, b, n).build()?   problem.rs:60-131
    problem.solve(Ode::X, OdeOptionMap) / .ode45(opts) / .ode4() ...                      problem.rs:133-190
    OdeSolution { tout, yout }                                                            solution.rs:22-29
    Ode enum + FromStr                                                                    mod.rs:14-46
    option newtypes Reltol/Abstol/Minstep/Maxstep/Initstep/Points, OdeOptionMap           options.rs
    OdeError                                                                              error.rs:4-23
Differences forced by the device: `fun` takes a device functor (`Rhs.builtin(...)` or `Rhs.from_source(...)`),
its captured constants are passed with `.params(...)`, and `init` may be one state or an [ntraj, n] ensemble.

Everything numeric happens in the CUDA library behind the C ABI (include/diffeq_b200.h).  There is no CPU
fallback: importing works without a GPU (so the ABI can be inspected), but any solve without the extension or
without a CUDA device raises.
"""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("DEQ_LIB", os.path.join(_HERE, "libdiffeq_b200.so"))  # DEQ_LIB: experiment variants
_lib = None


class OdeError(Exception):
    """Mirrors `OdeError` (error.rs:4-23); `.status` is the deq_status code, `.kind` the variant name."""
    KINDS = {1: "Uninitialized", 2: "InvalidButcherTableauWeightType", 3: "NAN", 4: "ZeroTimeSpan",
             5: "InvalidInitstep", 6: "InvalidMatrix", 7: "Cuda", 8: "Nvrtc", 9: "Unsupported", 10: "InvalidArg"}

    def __init__(self, status, msg):
        self.status = status
        self.kind = self.KINDS.get(status, "Unknown")
        super().__init__(f"{self.kind}: {msg}")


class Ode(enum.IntEnum):
    """`enum Ode` (mod.rs:14-26), same order; Ode21 is the `.ode21()` method (problem.rs:164-166)."""
    Feuler = 0
    Heun = 1
    Midpoint = 2
    Ode23 = 3
    Ode23s = 4
    Ode4 = 5
    Ode45 = 6
    Ode45fe = 7
    Ode4skr = 8
    Ode4ss = 9
    Ode78 = 10
    Ode21 = 11

    @staticmethod
    def from_str(s):
        """`impl FromStr for Ode` (mod.rs:28-46) — including its quirks: no name for Ode45fe, "ode4s" -> Ode4ss."""
        table = {"feuler": Ode.Feuler, "heun": Ode.Heun, "midpoint": Ode.Midpoint, "ode23": Ode.Ode23,
                 "ode23s": Ode.Ode23s, "ode4": Ode.Ode4, "ode45": Ode.Ode45, "ode4skr": Ode.Ode4skr,
                 "ode4s": Ode.Ode4ss, "ode78": Ode.Ode78}
        if s not in table:
            raise ValueError(f"{s} is not a valid Ode identifier")
        return table[s]


class TableauSet(enum.IntEnum):
    Reference = 0
    Textbook = 1


class Points(enum.IntEnum):
    """`Points` (options.rs:120-137)."""
    All = 0
    Specified = 1


class Output(enum.IntEnum):
    Final = 0
    Tspan = 1
    All = 2


class Mode(enum.IntEnum):
    """Arithmetic mode (deq_mode): Parity = bit-identical to the reference restatement, Fast = FMA / zero-skipping."""
    Parity = 0
    Fast = 1


class Libm(enum.IntEnum):
    """deq_libm: the libm calls as written in the source, or as LLVM folds them in an optimised build (sqrt / exp10)."""
    AsWritten = 0
    LlvmFolded = 1


class DenseLayout(enum.IntEnum):
    """deq_dense_layout: SoA [n][ntspan][ntraj] (sorted sweeps) or TrajMajor [ntraj][ntspan][n] (smem-staged, any order)."""
    SoA = 0
    TrajMajor = 1


class Precision(enum.IntEnum):
    F64 = 0
    F32 = 1


class System(enum.IntEnum):
    Lorenz = 0
    VanDerPol = 1
    Pendulum = 2


class _Options(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("has_minstep", C.c_int), ("minstep", C.c_double),
                ("has_maxstep", C.c_int), ("maxstep", C.c_double), ("initstep", C.c_double), ("points", C.c_int),
                ("output", C.c_int), ("max_steps", C.c_uint64), ("refill", C.c_int), ("async_", C.c_int),
                ("mode", C.c_int), ("precision", C.c_int), ("libm", C.c_int), ("dense_layout", C.c_int),
                ("shard_chunk", C.c_uint64)]


# The symbols include/diffeq_b200.h declares (checked by tests/test_abi.py).
ABI_SYMBOLS = [
    "deq_last_error_message", "deq_version", "deq_abi_options_size", "deq_rhs_builtin", "deq_rhs_from_source", "deq_rhs_dim",
    "deq_rhs_nparams", "deq_rhs_destroy", "deq_tspan_linspace", "deq_ensemble_create", "deq_ensemble_create_device",
    "deq_ensemble_destroy", "deq_ensemble_ntraj", "deq_set_stream", "deq_set_device", "deq_get_device", "deq_solve_host", "deq_solve_host_dense", "deq_options_default", "deq_solve",
    "deq_result_final", "deq_result_counts", "deq_result_totals", "deq_result_dense", "deq_result_all_offsets",
    "deq_result_all_copy", "deq_result_device_ptrs",
    "deq_result_trajectory", "deq_result_kernel_ms", "deq_result_destroy", "deq_tableau_get", "deq_rosenbrock_get", "deq_host_alloc",
    "deq_host_free", "deq_test_pow", "deq_test_log10", "deq_test_div", "deq_measure_fp64_peak",
]


def lib():
    """Loads the CUDA extension; fails loudly if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m diffeq_b200.build` "
                              "(diffeq_b200 has no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        L.deq_last_error_message.restype = C.c_char_p
        L.deq_version.restype = C.c_char_p
        L.deq_options_default.restype = _Options
        L.deq_ensemble_ntraj.restype = C.c_size_t
        L.deq_abi_options_size.restype = C.c_size_t
        for name in ABI_SYMBOLS:
            getattr(L, name)
        if L.deq_abi_options_size() != C.sizeof(_Options):
            raise ImportError("diffeq_b200: deq_options layout of the Python binding does not match the library")
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise OdeError(rc, lib().deq_last_error_message().decode())


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _check_shapes(rhs, y0, params):
    """The C ABI reads ntraj * n state and nparams (or ntraj * nparams) parameter doubles through raw pointers: refuse
    anything else here.  Returns (y0 [ntraj, n], params or None, params_per_traj)."""
    y0 = _f64(y0)
    if y0.ndim == 1:
        y0 = y0[None, :]
    if y0.ndim != 2 or y0.shape[1] != rhs.dim:
        raise OdeError(10, f"state array has shape {y0.shape}, the right-hand side expects [ntraj, {rhs.dim}]")
    npar = rhs.nparams
    if npar == 0:
        return np.ascontiguousarray(y0), None, 0
    if params is None:
        raise OdeError(1, "Required problem parameters must be initialized")
    params = _f64(params)
    if params.ndim == 1 and params.shape[0] == npar:
        return np.ascontiguousarray(y0), params, 0
    if params.ndim == 2 and params.shape == (y0.shape[0], npar):
        return np.ascontiguousarray(y0), params, 1
    raise OdeError(10, f"parameter array has shape {params.shape}, expected [{npar}] or [{y0.shape[0]}, {npar}]")


def _out_array(out, key, shape, dtype):
    """A caller-provided result array must be C-contiguous with the exact shape and an item size matching the C type."""
    a = out.get(key) if out is not None else None
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"] or tuple(a.shape) != tuple(shape) or a.dtype.itemsize != np.dtype(dtype).itemsize:
        raise OdeError(10, f"out[{key!r}] must be a C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
    return a


def test_div(seed, count):
    """Device self-test of the controller's division / reciprocal sequences: (bit mismatches, results checked)."""
    m, c = C.c_uint64(), C.c_uint64()
    _check(lib().deq_test_div(C.c_uint64(seed), C.c_uint64(count), C.byref(m), C.byref(c)))
    return m.value, c.value


def set_device(device):
    """Select the GPU for this thread's subsequent library calls (mirror torch.cuda.set_device here)."""
    _check(lib().deq_set_device(int(device)))


def set_stream(cuda_stream_ptr):
    """Enqueue this thread's subsequent library work on the given cudaStream_t (int), e.g. torch's current stream."""
    _check(lib().deq_set_stream(C.c_void_p(cuda_stream_ptr or 0)))


def linspace(a, b, n):
    """`tspan_linspace` (problem.rs:84-87)."""
    out = np.empty(n, dtype=np.float64)
    _check(lib().deq_tspan_linspace(C.c_double(a), C.c_double(b), C.c_size_t(n), _ptr(out)))
    return out


def tableau(method, tableau_set=TableauSet.Reference):
    S = C.c_int(); ad = C.c_int(); om = C.c_int(); fs = C.c_int()
    a = np.zeros(169); b = np.zeros(26); c = np.zeros(13)
    _check(lib().deq_tableau_get(int(method), int(tableau_set), C.byref(S), _ptr(a), _ptr(b), _ptr(c), C.byref(ad),
                                 C.byref(om), C.byref(fs)))
    s = S.value
    return dict(S=s, a=a[:s * s].reshape(s, s).copy(), b=b[:2 * s].reshape(s, 2).copy(), c=c[:s].copy(),
                adaptive=bool(ad.value), order_min=om.value, fsal=bool(fs.value))


STIFF = (Ode.Ode23s, Ode.Ode4skr, Ode.Ode4ss)


def _traits(method, tableau_set=TableauSet.Reference):
    """(adaptive, has_counters, stiff) of a method.  The stiff members of `Ode` (problem.rs:139,143-144) have no Butcher
    tableau; ode23s is adaptive, oderosenbrock walks the tspan grid but can stop a trajectory (InvalidMatrix)."""
    method = Ode(int(method))
    if method in STIFF:
        return method == Ode.Ode23s, True, True
    ad = tableau(method, tableau_set)["adaptive"]
    return ad, ad, False


def rosenbrock_coeffs(method):
    """`RosenbrockCoeffs::kr4()` / `s4()` (rosenbrock.rs:14-118) as the kernels hold them."""
    g = C.c_double(); a = np.zeros(16); b = np.zeros(4); c = np.zeros(16)
    _check(lib().deq_rosenbrock_get(int(method), C.byref(g), _ptr(a), _ptr(b), _ptr(c)))
    return dict(gamma=g.value, a=a.reshape(4, 4), b=b, c=c.reshape(4, 4))


def measure_fp64_peak():
    tf = C.c_double(); mhz = C.c_double()
    _check(lib().deq_measure_fp64_peak(C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def device_pow(x, y):
    x, y = _f64(x), _f64(y)
    out = np.empty_like(x)
    _check(lib().deq_test_pow(_ptr(x), _ptr(y), _ptr(out), C.c_size_t(x.size)))
    return out


def device_log10(x):
    x = _f64(x)
    out = np.empty_like(x)
    _check(lib().deq_test_log10(_ptr(x), _ptr(out), C.c_size_t(x.size)))
    return out


class Rhs:
    """Device right-hand side: replaces the closure `F: Fn(f64,&Y)->Y` (problem.rs:32-36)."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def builtin(system):
        h = C.c_void_p()
        _check(lib().deq_rhs_builtin(int(system), C.byref(h)))
        return Rhs(h)

    @staticmethod
    def from_source(cuda_src, n, nparams):
        h = C.c_void_p()
        _check(lib().deq_rhs_from_source(cuda_src.encode(), int(n), int(nparams), C.byref(h)))
        return Rhs(h)

    @property
    def dim(self):
        return lib().deq_rhs_dim(self._h)

    @property
    def nparams(self):
        return lib().deq_rhs_nparams(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.deq_rhs_destroy(self._h)
            self._h = None


# ---- option newtypes (options.rs:258-281) ----
class _Opt:
    name = None

    def __init__(self, v):
        self.v = v


class Reltol(_Opt): name = "Reltol"
class Abstol(_Opt): name = "Abstol"
class Minstep(_Opt): name = "Minstep"
class Maxstep(_Opt): name = "Maxstep"
class Initstep(_Opt): name = "Initstep"
class MaxSteps(_Opt): name = "MaxSteps"      # extension (guard; the reference has none)
class Refill(_Opt): name = "Refill"          # extension (scheduling only; results do not depend on it)
class OutputOpt(_Opt): name = "Output"       # extension (what to materialise)
class ShardChunk(_Opt): name = "ShardChunk"  # extension (multi-device deq_solve_host: block-cyclic chunk size, 0 = contiguous)


class OdeOptionMap(dict):
    """`OdeOptionMap` (options.rs:8-25): a map keyed by option name.  Unknown/unused reference options
    (Norm, StepTimeout, Retries) are accepted and ignored exactly like the reference ignores them."""

    def insert(self, *opts):
        for o in opts:
            if isinstance(o, Points):
                self["Points"] = o
            elif isinstance(o, Output):
                self["Output"] = o
            elif isinstance(o, Mode):
                self["Mode"] = o
            elif isinstance(o, Precision):
                self["Precision"] = o
            elif isinstance(o, Libm):
                self["Libm"] = o
            elif isinstance(o, DenseLayout):
                self["DenseLayout"] = o
            else:
                self[o.name] = o.v
        return self

    def to_c(self):
        o = lib().deq_options_default()
        if "Reltol" in self: o.reltol = float(self["Reltol"])
        if "Abstol" in self: o.abstol = float(self["Abstol"])
        if "Minstep" in self: o.has_minstep, o.minstep = 1, float(self["Minstep"])
        if "Maxstep" in self: o.has_maxstep, o.maxstep = 1, float(self["Maxstep"])
        if "Initstep" in self: o.initstep = float(self["Initstep"])
        if "Points" in self: o.points = int(self["Points"])
        if "Output" in self: o.output = int(self["Output"])
        if "MaxSteps" in self: o.max_steps = int(self["MaxSteps"])
        if "Refill" in self: o.refill = int(self["Refill"])
        if "Mode" in self: o.mode = int(self["Mode"])
        if "Precision" in self: o.precision = int(self["Precision"])
        if "Libm" in self: o.libm = int(self["Libm"])
        if "DenseLayout" in self: o.dense_layout = int(self["DenseLayout"])
        if "ShardChunk" in self: o.shard_chunk = int(self["ShardChunk"])
        return o


class OdeSolution:
    """`OdeSolution{tout, yout}` (solution.rs:22-29) for an ensemble, plus the diagnostics the reference drops.

    final      [ntraj, n]  last element of each trajectory's yout
    accepted / rejected / nfe / status / t_reached   per trajectory (adaptive methods)
    tout, yout  tspan and [ntraj, ntspan, n] values when solved with Output.Tspan (rows >= nvalid[i] except the
                last are NaN: the reference never emitted them)
    """

    def __init__(self):
        self.final = None
        self.accepted = self.rejected = self.nfe = self.status = self.t_reached = None
        self.tout = self.yout = self.nvalid = None
        self.kernel_ms = None
        self.totals = None
        self.offsets = self.tout_flat = self.yout_flat = None   # Output.All: ragged Points::All streams

    def trajectory(self, i):
        """Output.All: the reference-shaped (tout, yout) of trajectory i (problem.rs:354-357)."""
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.tout_flat[a:b], self.yout_flat[a:b]

    def zipped(self, traj=0):
        """`OdeSolution::zipped` (solution.rs:33-36) for one trajectory."""
        return list(zip(self.tout, self.yout[traj]))


class OdeProblem:
    """`OdeProblem` (problem.rs:32-47) over an ensemble of initial states / parameter sets."""

    def __init__(self, rhs, y0, tspan, params, single):
        self._rhs = rhs
        self._single = single
        y0, params, per = _check_shapes(rhs, y0, params)
        self.ntraj, self.n = y0.shape
        self.tspan = tspan
        h = C.c_void_p()
        _check(lib().deq_ensemble_create(rhs._h, C.c_size_t(self.ntraj), _ptr(y0), _ptr(params), per, _ptr(tspan),
                                         C.c_size_t(len(tspan)), C.byref(h)))
        self._h = h

    @staticmethod
    def builder():
        return OdeBuilder()

    @classmethod
    def from_device(cls, rhs, d_y0_soa_ptr, ntraj, tspan, params=None, d_params_soa_ptr=None):
        """Ensemble whose inputs are already resident on the current GPU in the library layout (`deq_ensemble_create_device`):
        d_y0_soa_ptr -> f64 [n][ntraj]; per-trajectory parameters as a device pointer to [nparams][ntraj], or shared host
        values in `params`.  The caller keeps the device buffers alive for the lifetime of the problem."""
        self = cls.__new__(cls)
        self._rhs, self._single = rhs, False
        self.ntraj, self.n = int(ntraj), rhs.dim
        self.tspan = _f64(tspan)
        h = C.c_void_p()
        if d_params_soa_ptr is not None:
            pp, per = C.cast(C.c_void_p(int(d_params_soa_ptr)), C.POINTER(C.c_double)), 1
        else:
            self._params_host = _f64(params) if params is not None else None
            pp, per = _ptr(self._params_host), 0
        _check(lib().deq_ensemble_create_device(rhs._h, C.c_size_t(self.ntraj), C.cast(C.c_void_p(int(d_y0_soa_ptr)), C.POINTER(C.c_double)),
                                                pp, per, _ptr(self.tspan), C.c_size_t(len(self.tspan)), C.byref(h)))
        self._h = h
        return self

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.deq_ensemble_destroy(self._h)
            self._h = None

    # -- solve + per-method wrappers (problem.rs:133-190) --
    def solve(self, ode, opts=None, tableau_set=TableauSet.Reference, fetch=True):
        opts = opts if opts is not None else OdeOptionMap()
        o = opts.to_c()
        res = C.c_void_p()
        L = lib()
        _check(L.deq_solve(self._h, int(ode), int(tableau_set), C.byref(o), C.byref(res)))
        try:
            return self._collect(res, o, fetch, _traits(ode, tableau_set))
        finally:
            L.deq_result_destroy(res)

    def _collect(self, res, o, fetch, traits):
        adaptive, has_counts, stiff = traits
        L = lib()
        sol = OdeSolution()
        ms = C.c_float()
        _check(L.deq_result_kernel_ms(res, C.byref(ms)))
        sol.kernel_ms = ms.value
        a = C.c_uint64(); r = C.c_uint64(); f = C.c_uint64()
        _check(L.deq_result_totals(res, C.byref(a), C.byref(r), C.byref(f)))
        sol.totals = dict(accepted=a.value, rejected=r.value, nfe=f.value)
        if not fetch:
            return sol
        nt = len(self.tspan)
        if nt == 0:   # problem.rs:211-214: an empty tspan yields an empty OdeSolution
            sol.final = np.empty((0, self.n))
            sol.tout = np.empty(0)
            sol.yout = np.empty((self.ntraj, 0, self.n))
            return sol
        sol.final = np.empty((self.ntraj, self.n))
        if self.ntraj:
            _check(L.deq_result_final(res, _ptr(sol.final)))
        if has_counts and nt > 0:
            sol.accepted = np.empty(self.ntraj, np.uint32); sol.rejected = np.empty(self.ntraj, np.uint32)
            sol.nfe = np.empty(self.ntraj, np.uint32); sol.status = np.empty(self.ntraj, np.uint8)
            sol.t_reached = np.empty(self.ntraj)
            _check(L.deq_result_counts(res, _ptr(sol.accepted, C.c_uint32), _ptr(sol.rejected, C.c_uint32),
                                       _ptr(sol.nfe, C.c_uint32), _ptr(sol.status, C.c_uint8), _ptr(sol.t_reached)))
        if o.output == Output.All and nt > 0 and adaptive:
            sol.offsets = np.zeros(self.ntraj + 1, np.uint64)
            _check(L.deq_result_all_offsets(res, _ptr(sol.offsets, C.c_uint64)))
            rows = int(sol.offsets[-1])
            sol.tout_flat = np.empty(rows)
            sol.yout_flat = np.empty((rows, self.n))
            _check(L.deq_result_all_copy(res, C.c_size_t(0), C.c_size_t(self.ntraj), _ptr(sol.tout_flat), _ptr(sol.yout_flat)))
        if o.output == Output.Tspan and nt > 0:
            tm = o.dense_layout == DenseLayout.TrajMajor and adaptive and not stiff
            dense = np.full((self.ntraj, nt, self.n) if tm else (self.n, nt, self.ntraj), np.nan)
            nv = np.empty(self.ntraj, np.uint32)
            _check(L.deq_result_dense(res, C.c_size_t(0), C.c_size_t(self.ntraj), _ptr(dense), _ptr(nv, C.c_uint32)))
            idx = np.arange(nt)[:, None]
            invalid = idx >= nv[None, :]
            if not stiff:   # oderk_adapt always pushes the end point as the last row (problem.rs:320-322)
                invalid &= idx != nt - 1
            if tm:
                dense[invalid.T] = np.nan
                sol.yout = dense
            else:
                dense[:, invalid] = np.nan
                sol.yout = np.ascontiguousarray(dense.transpose(2, 1, 0))
            # oderosenbrock returns `tout: diff(tspan)` (problem.rs:572,639 — sic, one element short of yout)
            sol.tout = np.diff(self.tspan) if (stiff and not adaptive) else self.tspan.copy()
            sol.nvalid = nv
        return sol

    def feuler(self): return self.solve(Ode.Feuler)
    def heun(self): return self.solve(Ode.Heun)
    def midpoint(self): return self.solve(Ode.Midpoint)
    def ode4(self): return self.solve(Ode.Ode4)
    def ode21(self, opts=None, **kw): return self.solve(Ode.Ode21, opts, **kw)
    def ode23(self, opts=None, **kw): return self.solve(Ode.Ode23, opts, **kw)
    def ode45(self, opts=None, **kw): return self.solve(Ode.Ode45, opts, **kw)
    def ode45_dp(self, opts=None, **kw): return self.solve(Ode.Ode45, opts, **kw)
    def ode45_fe(self, opts=None, **kw): return self.solve(Ode.Ode45fe, opts, **kw)
    def ode78(self, opts=None, **kw): return self.solve(Ode.Ode78, opts, **kw)
    def ode23s(self, opts=None, **kw): return self.solve(Ode.Ode23s, opts, **kw)     # problem.rs:409
    def ode4s_kr(self, opts=None, **kw): return self.solve(Ode.Ode4skr, opts, **kw)  # problem.rs:643
    def ode4s_s(self, opts=None, **kw): return self.solve(Ode.Ode4ss, opts, **kw)    # problem.rs:648


class OdeBuilder:
    """`OdeBuilder` (problem.rs:49-119): fun / init / tspan / tspan_linspace / build."""

    def __init__(self):
        self._f = None
        self._y0 = None
        self._tspan = None
        self._params = None

    def fun(self, rhs):
        self._f = rhs
        return self

    def params(self, p):
        self._params = _f64(p)
        return self

    def init(self, y0):
        self._y0 = _f64(y0)
        return self

    def tspan(self, ts):
        self._tspan = _f64(ts)
        return self

    def tspan_linspace(self, a, b, n):
        self._tspan = linspace(a, b, n)
        return self

    def build(self):
        # problem.rs:92-104
        if self._f is None:
            raise OdeError(1, "Required problem must be initialized")
        if self._y0 is None:
            raise OdeError(1, "Initial starting point must be initialized")
        if self._tspan is None:
            raise OdeError(1, "Time span must be initialized")
        y0 = self._y0
        single = y0.ndim == 1
        if single:
            y0 = y0[None, :]
        if y0.shape[1] != self._f.dim:
            raise OdeError(10, f"state has {y0.shape[1]} components, the right-hand side expects {self._f.dim}")
        if self._f.nparams and self._params is None:
            raise OdeError(1, "Required problem parameters must be initialized")
        return OdeProblem(self._f, np.ascontiguousarray(y0), self._tspan, self._params, single)


def solve_host(rhs, y0, params, tspan, ode, opts=None, out_final=None, out_accepted=None, out_rejected=None,
               tableau_set=TableauSet.Reference, return_totals=False):
    """End-to-end call on HOST buffers (the drop-in shape of `OdeProblem::solve` for an ensemble): uploads y0 / params /
    tspan, integrates on the GPU, and copies the final states and step counters back into the given (ideally pinned)
    buffers.  Returns (final, accepted, rejected) — plus the device-reduced totals dict when return_totals is set."""
    L = lib()
    y0, params, per = _check_shapes(rhs, y0, params)
    ntraj, n = y0.shape
    tspan = _f64(tspan)
    o = (opts if opts is not None else OdeOptionMap()).to_c()
    adaptive = _traits(ode, tableau_set)[1]
    ens = C.c_void_p(); res = C.c_void_p()
    _check(L.deq_ensemble_create(rhs._h, C.c_size_t(ntraj), _ptr(y0), _ptr(params), per, _ptr(tspan),
                                 C.c_size_t(len(tspan)), C.byref(ens)))
    totals = None
    try:
        o.async_ = 1
        _check(L.deq_solve(ens, int(ode), int(tableau_set), C.byref(o), C.byref(res)))
        try:
            if out_final is None:
                out_final = np.empty((ntraj, n))
            _check(L.deq_result_final(res, _ptr(out_final)))
            if adaptive:
                if out_accepted is None:
                    out_accepted = np.empty(ntraj, np.uint32)
                if out_rejected is None:
                    out_rejected = np.empty(ntraj, np.uint32)
                _check(L.deq_result_counts(res, _ptr(out_accepted, C.c_uint32), _ptr(out_rejected, C.c_uint32), None,
                                           None, None))
            if return_totals:
                a = C.c_uint64(); r = C.c_uint64(); f = C.c_uint64()
                _check(L.deq_result_totals(res, C.byref(a), C.byref(r), C.byref(f)))
                totals = dict(accepted=a.value, rejected=r.value, nfe=f.value)
        finally:
            L.deq_result_destroy(res)
    finally:
        L.deq_ensemble_destroy(ens)
    if return_totals:
        return out_final, out_accepted, out_rejected, totals
    return out_final, out_accepted, out_rejected


def solve_host_multi(rhs, y0, params, tspan, ode, opts=None, devices=None, tableau_set=TableauSet.Reference, out=None):
    """`deq_solve_host`: one C-ABI call on host buffers, sharded over `devices` (list of CUDA ordinals; None = current device)
    with one host thread per GPU inside the library.  Returns dict(final, accepted, rejected, nfe, status, t_reached); pass
    `out` (a dict of preallocated, ideally pinned, arrays with those keys) to avoid allocating the results per call.
    With `Output.Tspan` in `opts` the call is `deq_solve_host_dense`: the result also carries `dense` ([ntraj, ntspan, n] for
    DenseLayout.TrajMajor on the explicit adaptive methods, else [n, ntspan, ntraj]) and `nvalid`; chunks of the ensemble are
    integrated while earlier chunks travel to the host."""
    L = lib()
    y0, params, per = _check_shapes(rhs, y0, params)
    ntraj, n = y0.shape
    tspan = _f64(tspan)
    nt = len(tspan)
    o = (opts if opts is not None else OdeOptionMap()).to_c()
    dev = (C.c_int * len(devices))(*devices) if devices else None
    dense = o.output == Output.Tspan
    adaptive, _, stiff = _traits(ode, tableau_set)
    tm = dense and o.dense_layout == DenseLayout.TrajMajor and adaptive and not stiff
    dshape = (ntraj, nt, n) if tm else (n, nt, ntraj)
    if out is None:
        out = dict(final=np.empty((ntraj, n)), accepted=np.zeros(ntraj, np.uint32), rejected=np.zeros(ntraj, np.uint32),
                   nfe=np.zeros(ntraj, np.uint32), status=np.zeros(ntraj, np.uint8), t_reached=np.zeros(ntraj))
        if dense:
            out["dense"] = np.empty(dshape)
            out["nvalid"] = np.zeros(ntraj, np.uint32)
    final = _out_array(out, "final", (ntraj, n), np.float64)
    if final is None:
        raise OdeError(10, "out['final'] is required")
    args = [rhs._h, C.c_size_t(ntraj), _ptr(y0), _ptr(params), per, _ptr(tspan), C.c_size_t(nt), int(ode), int(tableau_set), C.byref(o),
            dev, len(devices) if devices else 0, _ptr(final),
            _ptr(_out_array(out, "accepted", (ntraj,), np.uint32), C.c_uint32), _ptr(_out_array(out, "rejected", (ntraj,), np.uint32), C.c_uint32),
            _ptr(_out_array(out, "nfe", (ntraj,), np.uint32), C.c_uint32), _ptr(_out_array(out, "status", (ntraj,), np.uint8), C.c_uint8),
            _ptr(_out_array(out, "t_reached", (ntraj,), np.float64))]
    if dense:
        d = _out_array(out, "dense", dshape, np.float64)
        if d is None:
            raise OdeError(10, "out['dense'] is required with Output.Tspan")
        _check(L.deq_solve_host_dense(*args, _ptr(d), _ptr(_out_array(out, "nvalid", (ntraj,), np.uint32), C.c_uint32)))
    else:
        _check(L.deq_solve_host(*args))
    return out
