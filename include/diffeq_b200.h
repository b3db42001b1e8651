/* diffeq_b200.h — C ABI of libdiffeq_b200.so: B200-native ensemble Runge–Kutta integration behind the
 * problem/solver surface of the Rust crate mattsse/diffeq.
 *
 * The reference has no FFI of its own; its boundary for this path is the Rust API
 *     OdeProblem::builder().fun(f).init(y0).tspan(ts).build()?      src/ode/problem.rs:60-131
 *     problem.solve(Ode::X, OdeOptionMap) -> Result<OdeSolution, OdeError>   src/ode/problem.rs:133-147
 * Every entry point below names the reference item it replaces.  A Rust `-sys` crate binds these symbols
 * one-to-one (see INTEGRATION.md and rust/); everything is `extern "C"`, plain pointers and sizes.
 *
 * Conventions: all calls return deq_status (DEQ_OK == 0); deq_last_error_message() gives a thread-local detail
 * string.  The caller owns every buffer it passes; inputs are copied during *_create, outputs are copied into
 * caller-provided buffers.  Handles are not internally synchronised (one thread per handle at a time).  One
 * handle lives on one GPU (the current CUDA device at creation); an ensemble is sharded over GPUs by creating one
 * handle per device/process over a contiguous slice of the trajectories (no inter-GPU traffic is needed: every
 * trajectory is an independent OdeProblem, problem.rs:32-47).
 */
#ifndef DIFFEQ_B200_H
#define DIFFEQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors OdeError (src/error.rs:4-23) + backend codes. */
typedef enum deq_status {
    DEQ_OK = 0,
    DEQ_ERR_UNINITIALIZED = 1,     /* OdeError::Uninitialized (problem.rs:92-104) */
    DEQ_ERR_WEIGHT_TYPE = 2,       /* OdeError::InvalidButcherTableauWeightType (problem.rs:204-209) */
    DEQ_ERR_NAN = 3,               /* OdeError::NAN (never produced by the RK path) */
    DEQ_ERR_ZERO_TIMESPAN = 4,     /* OdeError::ZeroTimeSpan (unreachable for f64: signum(0) == 1) */
    DEQ_ERR_INVALID_INITSTEP = 5,  /* OdeError::InvalidInitstep (problem.rs:232-237) */
    DEQ_ERR_INVALID_MATRIX = 6,    /* OdeError::InvalidMatrix (problem.rs:481,588).  In an ensemble the failing trajectory
                                      stops with DEQ_TRAJ_INVALID_MATRIX and deq_solve itself succeeds */
    DEQ_ERR_CUDA = 7,
    DEQ_ERR_NVRTC = 8,
    DEQ_ERR_UNSUPPORTED = 9,       /* a combination this build does not carry (e.g. stiff methods with n > 32) */
    DEQ_ERR_INVALID_ARG = 10
} deq_status;

/* Same variants and order as `enum Ode` (src/ode/mod.rs:14-26); DEQ_ODE21 exposes `.ode21()` (problem.rs:164). */
typedef enum deq_method {
    DEQ_FEULER = 0, DEQ_HEUN = 1, DEQ_MIDPOINT = 2, DEQ_ODE23 = 3, DEQ_ODE23S = 4, DEQ_ODE4 = 5, DEQ_ODE45 = 6,
    DEQ_ODE45FE = 7, DEQ_ODE4SKR = 8, DEQ_ODE4SS = 9, DEQ_ODE78 = 10, DEQ_ODE21 = 11
} deq_method;

/* Which coefficients the adaptive tableaus use: REFERENCE = bug-for-bug runge_kutta.rs:231-817,
 * TEXTBOOK = corrected rk21 / rk23 / dopri5 c[3] / feh78.  For DEQ_ODE23S: REFERENCE inverts the packed LU factors of
 * W = I - h d J like problem.rs:469,481 does; TEXTBOOK inverts W itself (Shampine & Reichelt's ode23s). */
typedef enum deq_tableau_set { DEQ_TABLEAU_REFERENCE = 0, DEQ_TABLEAU_TEXTBOOK = 1 } deq_tableau_set;

typedef enum deq_system { DEQ_SYS_LORENZ = 0, DEQ_SYS_VANDERPOL = 1, DEQ_SYS_PENDULUM = 2 } deq_system;

/* `Points` (src/ode/options.rs:120-137).  Points::All is the reference-valid choice.  Points::Specified on the explicit
 * adaptive methods re-reads y from the output vector (problem.rs:262) and yields wrong trajectories whenever the tspan grid
 * is coarser than the steps; it is reproduced bug-for-bug (f64 parity kernels, output = DEQ_OUT_FINAL or DEQ_OUT_ALL — the
 * rows of that stream are exactly the tspan points; DEQ_OUT_TSPAN answers DEQ_ERR_UNSUPPORTED).  For correct tspan-only rows
 * use DEQ_POINTS_ALL with output = DEQ_OUT_TSPAN.  DEQ_ODE23S honours both kinds (problem.rs:536): with DEQ_OUT_ALL,
 * Specified yields the rows at the tspan times only, All adds every step end point. */
typedef enum deq_points { DEQ_POINTS_ALL = 0, DEQ_POINTS_SPECIFIED = 1 } deq_points;

/* What deq_solve materialises. */
typedef enum deq_output {
    DEQ_OUT_FINAL = 0,  /* last element of yout + per-trajectory counters */
    DEQ_OUT_TSPAN = 1,  /* additionally the rows of (tout,yout) that fall on tspan times, SoA [n][ntspan][ntraj] */
    DEQ_OUT_ALL = 2     /* the reference's complete Points::All output per trajectory (problem.rs:301-323): y0, every
                           Hermite value at a tspan point inside a step, every step end point — ragged, two passes */
} deq_output;

/* Arithmetic of the adaptive kernels.  PARITY (default): every operation rounded and ordered as in the reference,
 * glibc-exact pow — results are bit-identical to the CPU restatement of the crate.  FAST: fused multiply-adds,
 * structurally-zero tableau entries skipped, one pow per step, ~2x the throughput.  FAST is NOT bit-exact.  What is
 * measured (bench.py `mode_check`, 2^20 Lorenz trajectories, dopri5, rtol = atol = 1e-8, T = 10): identical accepted and
 * rejected step counts on > 99.9 % of the trajectories; final states within 10 x (rtol |y| + atol) of PARITY on 99.99 %
 * of them, median deviation ~1e-2 of that unit, largest several hundred units — a chaotic trajectory amplifies any
 * rounding difference by up to ~1e10 over that horizon, so no arithmetic other than the bit-identical one can stay
 * within 10 x rtol of PARITY on all of them.  Against a high-precision solution FAST is as accurate as PARITY
 * (tests/test_gpu_parity.py::test_fast_mode_accuracy_against_truth).  DEQ_ODE23S has a FAST flavour too (FMA, one
 * reciprocal per inverse / Jacobian column, sqrt norms); the other methods always run PARITY. */
typedef enum deq_mode { DEQ_MODE_PARITY = 0, DEQ_MODE_FAST = 1 } deq_mode;

/* Optional single-precision state (north_star "optional f32 mode").  The reference cannot instantiate f32 (types.rs:262-278),
 * so F32 has no reference result to match; it runs the FAST formulas in float (time stays f64) and is validated against
 * the f64 path at f32 tolerances.  Adaptive methods, built-in systems, FINAL / TSPAN output. */
typedef enum deq_precision { DEQ_PRECISION_F64 = 0, DEQ_PRECISION_F32 = 1 } deq_precision;

/* Which libm calls the PARITY arithmetic reproduces.  The source calls `powf(1/p)` with p = 2 for the error norm
 * (types.rs:115) and `10f64.powf(x)` in hinit (problem.rs:893).  AS_WRITTEN evaluates exactly those (glibc pow).
 * LLVM_FOLDED evaluates what an optimised rustc build runs after LLVM's libcall simplification: pow(x, 0.5) -> sqrt(x)
 * (the exponent folds to the constant 0.5 once pnorm(PNorm::default()) is inlined) and pow(10.0, x) -> exp10(x).  The two
 * differ by one ulp in roughly 1 of 2000 norm evaluations; which one a given build of the crate contains cannot be
 * determined without a Rust toolchain, so both are provided, each bit-exact against the oracle in the same flavour. */
typedef enum deq_libm { DEQ_LIBM_AS_WRITTEN = 0, DEQ_LIBM_LLVM_FOLDED = 1 } deq_libm;

/* Memory layout of DEQ_OUT_TSPAN (adaptive methods).  SOA [n][ntspan][ntraj] is written by one of two kernels: in the
 * lane-refill kernel (refill = 1) every lane emits inside its own steps, which coalesces only while neighbouring
 * trajectories advance together (parameter-sorted sweeps: the fastest choice there); in the row-synchronous kernel
 * (refill = 0) a warp owns 32 consecutive trajectories and emits each row together — one 256-byte store per component and
 * row whatever the order of the ensemble, at the price of lanes waiting for each other.  refill = 2 (default) decides after
 * the initial-step pass: if neighbouring trajectories start with (almost) equal steps the lane-refill kernel runs; if they do
 * once the ensemble is SORTED by first step size (a sweep handed over in arbitrary order), the lane-refill kernel runs the
 * trajectories in that order and the columns are permuted back in place afterwards; otherwise the row-synchronous kernel.
 * The output is bit-identical in every case.  TRAJ_MAJOR [ntraj][ntspan][n] is one contiguous stream per trajectory —
 * the reference's Vec<Y> shape — written through per-lane shared-memory staging with cooperative 128-byte stores; it is
 * the layout for ensembles in arbitrary order (a sweep in disguise is dispatched in first-step order here too, and needs no
 * permutation afterwards: every stream is written where it belongs). */
typedef enum deq_dense_layout { DEQ_LAYOUT_SOA = 0, DEQ_LAYOUT_TRAJ_MAJOR = 1 } deq_dense_layout;

/* Per-trajectory completion code (the reference returns Ok(partial) silently in the first two cases). */
typedef enum deq_traj_status {
    DEQ_TRAJ_DONE = 0,
    DEQ_TRAJ_MINSTEP_BREAK = 1,  /* |dt| < minstep: problem.rs:342-344 */
    DEQ_TRAJ_MAX_STEPS = 2,      /* extension: options.max_steps reached */
    DEQ_TRAJ_INVALID_MATRIX = 3  /* stiff methods: try_inverse() hit a zero determinant (OdeError::InvalidMatrix) */
} deq_traj_status;

/* AdaptiveOptions (src/ode/options.rs:49-70); defaults from options.rs:283-299 and problem.rs:219-225.
 * `norm`, `step_timeout` and `retries` are accepted by the reference but never used and are
 * therefore absent. */
typedef struct deq_options {
    double reltol;       /* Reltol, default 1e-5 */
    double abstol;       /* Abstol, default 1e-8 */
    int has_minstep;     /* Minstep: None -> |tend - t0| / 1e18 */
    double minstep;
    int has_maxstep;     /* Maxstep: None -> |tend - t0| / 2.5 */
    double maxstep;
    double initstep;     /* Initstep, 0 = estimate with hinit (problem.rs:232-240) */
    int points;          /* deq_points, default DEQ_POINTS_ALL */
    int output;          /* deq_output, default DEQ_OUT_FINAL */
    uint64_t max_steps;  /* extension: stop a trajectory after this many step attempts; default 10^7, 0 = unlimited
                          (the reference has no guard and can loop forever: broken tableaus, backward runs) */
    int refill;          /* scheduling of the adaptive kernels (never changes a result): 2 (default) persistent lanes refilled
                            individually + tail compaction once the pool is empty; 1: lane refill only; 0: static warp assignment */
    int async;           /* 0 (default): deq_solve returns after the GPU finished; 1: returns after enqueueing */
    int mode;            /* deq_mode, default DEQ_MODE_PARITY (adaptive methods only; fixed-step is always parity) */
    int precision;       /* deq_precision, default DEQ_PRECISION_F64 */
    int libm;            /* deq_libm, default DEQ_LIBM_AS_WRITTEN (PARITY mode only) */
    int dense_layout;    /* deq_dense_layout, default DEQ_LAYOUT_SOA */
    uint64_t shard_chunk; /* deq_solve_host over several devices: 0 (default) = contiguous equal blocks of the trajectory range;
                            c > 0 = block-cyclic, chunk j of c trajectories goes to device j % ndevices — balances a
                            parameter-sorted sweep whose step counts grow along the index range.  Results
                            do not depend on it */
} deq_options;

typedef struct deq_rhs deq_rhs;
typedef struct deq_ensemble deq_ensemble;
typedef struct deq_result deq_result;

const char* deq_last_error_message(void);
const char* deq_version(void);
/* sizeof(deq_options) as compiled into the library: lets a foreign binding verify its struct layout. */
size_t deq_abi_options_size(void);

/* ---- right-hand side: replaces the closure `F: Fn(f64,&Y)->Y` (problem.rs:32-36) ---- */
deq_status deq_rhs_builtin(int system /* deq_system */, deq_rhs** out);
/* User CUDA source defining
 *     __device__ void rhs(double t, const double* y, double* dy, const double* p);
 * for an n-dimensional state (n <= 64) and nparams parameters (<= 8), compiled with NVRTC for sm_100a.  The source may call
 * deq_sin(x) / deq_cos(x): sin and cos with glibc's bits for every double — what f64::sin / f64::cos of a Rust closure
 * evaluate to — so that a right-hand side with trigonometric terms stays bit-identical to the reference.  Optional: a source
 * that also defines
 *     #define DEQ_USER_NG <count>
 *     __device__ void forcing(double t, double* g, const double* p);                                 // the time-only terms
 *     __device__ void rhs_g(double t, const double* y, double* dy, const double* p, const double* g); // rhs, given those terms
 * with rhs(t, y, dy, p) == rhs_g(t, y, dy, p, g) for g = forcing(t) has its forcing evaluated once per DISTINCT stage time
 * (rows of a tableau with equal c, the end of an accepted step, the start of the next one): same bits, fewer evaluations. */
deq_status deq_rhs_from_source(const char* cuda_src, int n, int nparams, deq_rhs** out);
int deq_rhs_dim(const deq_rhs* rhs);
int deq_rhs_nparams(const deq_rhs* rhs);
void deq_rhs_destroy(deq_rhs* rhs);

/* OdeBuilder::tspan_linspace (problem.rs:84-87; itertools_num::linspace): out[i] = from + ((to-from)/(n-1))*i */
deq_status deq_tspan_linspace(double from, double to, size_t n, double* out);

/* ---- problem: replaces OdeProblem / OdeBuilder (problem.rs:32-119), one problem per trajectory ----
 * y0     host, [ntraj][n] row-major (one `Y` per trajectory, like the reference's Vec<f64> state)
 * params host, [ntraj][nparams] if params_per_traj else [nparams] (may be NULL if nparams == 0)
 * tspan  host, [ntspan], shared by all trajectories (may be NULL -> DEQ_ERR_UNINITIALIZED like build()) */
deq_status deq_ensemble_create(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params,
                               int params_per_traj, const double* tspan, size_t ntspan, deq_ensemble** out);
/* Same, with inputs already resident on the current device in the library's layout:
 * y0_soa [n][ntraj], params_soa [nparams][ntraj] (or host [nparams] shared values if !params_per_traj). */
deq_status deq_ensemble_create_device(const deq_rhs* rhs, size_t ntraj, const double* d_y0_soa,
                                      const double* params, int params_per_traj, const double* tspan,
                                      size_t ntspan, deq_ensemble** out);
void deq_ensemble_destroy(deq_ensemble* ens);
size_t deq_ensemble_ntraj(const deq_ensemble* ens);

/* Select the GPU this thread's subsequent calls use (the library links its own CUDA runtime, so a caller that picked a
 * device with another runtime instance — e.g. torch.cuda.set_device — must tell the library too). */
deq_status deq_set_device(int device);
deq_status deq_get_device(int* device);
/* CUDA stream (cudaStream_t as void*) all work of this thread's subsequent calls is enqueued on; NULL = default. */
deq_status deq_set_stream(void* cuda_stream);

/* ---- solve: replaces OdeProblem::solve(Ode, OdeOptionMap) (problem.rs:133-147) ---- */
deq_options deq_options_default(void);
deq_status deq_solve(deq_ensemble* ens, int method /* deq_method */, int tableau_set /* deq_tableau_set */,
                     const deq_options* opts, deq_result** out);

/* One call for an ensemble on HOST buffers, optionally sharded over several GPUs of the box: trajectories are
 * split into contiguous blocks, one host thread + stream per device, no inter-GPU traffic; results land in the caller's
 * arrays (yfinal [ntraj][n]; the counter arrays [ntraj] may each be NULL; ignored for fixed-step methods).  devices == NULL
 * or ndevices <= 0 uses the current device.  The result of trajectory i does not depend on the device count. */
deq_status deq_solve_host(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                          const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options* opts,
                          const int* devices, int ndevices, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                          uint32_t* nfe, uint8_t* status, double* t_reached);

/* deq_solve_host with the tspan rows (opts->output must be DEQ_OUT_TSPAN): `dense` is a host buffer in the layout
 * opts->dense_layout names — TRAJ_MAJOR [ntraj][ntspan][n] (each trajectory's rows contiguous, the reference's Vec<Y> per
 * problem; explicit adaptive methods) or SOA [n][ntspan][ntraj] — and nvalid[ntraj] (may be NULL) receives the number of
 * leading rows the reference would have emitted.  The trajectory range is cut into chunks of ~128 MiB of rows that are dealt
 * round-robin to the devices; each device integrates chunk i + 1 while chunk i travels to the host on a second stream, so the
 * call runs at the speed of the slower of the two (for dense rows that is the host link: use pinned buffers, deq_host_alloc).
 * DEQ_OUT_ALL is not offered on host buffers: its size is data dependent (use deq_solve + deq_result_all_*). */
deq_status deq_solve_host_dense(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                                const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options* opts,
                                const int* devices, int ndevices, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                                uint32_t* nfe, uint8_t* status, double* t_reached, double* dense, uint32_t* nvalid);

/* ---- result: replaces OdeSolution{tout,yout} (src/ode/solution.rs:22-29) + the Diagnostics the reference
 * counts but drops (problem.rs:957-970).  All getters synchronise with the solve first. ---- */
deq_status deq_result_final(deq_result* res, double* y /* host [ntraj][n] */);
deq_status deq_result_counts(deq_result* res, uint32_t* accepted, uint32_t* rejected, uint32_t* nfe,
                             uint8_t* status, double* t_reached /* each host [ntraj], any may be NULL */);
deq_status deq_result_totals(deq_result* res, uint64_t* accepted, uint64_t* rejected, uint64_t* nfe);
/* DEQ_OUT_TSPAN: rows of trajectories [traj_begin, traj_end) into host out[n][ntspan][traj_end-traj_begin] (SOA) or
 * out[traj_end-traj_begin][ntspan][n] (TRAJ_MAJOR);
 * nvalid (may be NULL) receives, per trajectory, how many leading tspan rows the reference would have emitted. */
deq_status deq_result_dense(deq_result* res, size_t traj_begin, size_t traj_end, double* out, uint32_t* nvalid);
/* DEQ_OUT_ALL: offsets[ntraj+1] — trajectory i owns rows [offsets[i], offsets[i+1]) of the ragged (tout, yout). */
deq_status deq_result_all_offsets(deq_result* res, uint64_t* offsets);
/* DEQ_OUT_ALL: rows of trajectories [traj_begin, traj_end): tout[rows], yout[rows][n] (one `Y` per row like Vec<Y>). */
deq_status deq_result_all_copy(deq_result* res, size_t traj_begin, size_t traj_end, double* tout, double* yout);
/* Device views (library layout, valid until deq_result_destroy): yfinal [n][ntraj], dense [n][ntspan][ntraj]. */
deq_status deq_result_device_ptrs(deq_result* res, const double** d_yfinal_soa, const double** d_dense_soa);
/* Fixed-step methods with DEQ_OUT_TSPAN: the full OdeSolution of one trajectory, yout host [ntspan][n]. */
deq_status deq_result_trajectory(deq_result* res, size_t traj, double* yout);
/* Milliseconds the main stepping kernel took on the device (CUDA events around that launch). */
deq_status deq_result_kernel_ms(deq_result* res, float* ms);
void deq_result_destroy(deq_result* res);

/* ---- utilities ---- */
/* Tableau introspection (runge_kutta.rs:87-100): a[S*S] row-major, b[S*2] (stepping, error), c[S]. */
deq_status deq_tableau_get(int method, int tableau_set, int* stages, double* a, double* b, double* c,
                           int* adaptive, int* order_min, int* fsal);
/* Rosenbrock coefficient sets (rosenbrock.rs:14-118) of DEQ_ODE4SKR (kr4) / DEQ_ODE4SS (s4): gamma, a[16], b[4], c[16],
 * matrices row-major. */
deq_status deq_rosenbrock_get(int method, double* gamma, double* a, double* b, double* c);
/* Pinned host memory for fast H2D / D2H of caller buffers. */
deq_status deq_host_alloc(size_t bytes, void** out);
void deq_host_free(void* p);
/* Device self-tests used by the parity suite: evaluate the library's pow / log10 on the GPU. */
deq_status deq_test_pow(const double* x, const double* y, double* out, size_t n);
deq_status deq_test_log10(const double* x, double* out, size_t n);
/* The controller's straight-line division / reciprocal sequences against the IEEE operators on `count` hashed operand pairs:
 * number of bit mismatches among the `checked` results whose main-path flag stayed clear (must be 0). */
deq_status deq_test_div(uint64_t seed, uint64_t count, uint64_t* mismatches, uint64_t* checked);
/* FP64 roofline denominator: sustained DFMA rate of this GPU in TFLOP/s (FMA = 2 flop). */
deq_status deq_measure_fp64_peak(double* tflops, double* sm_clock_mhz_hint);

#ifdef __cplusplus
}
#endif
#endif /* DIFFEQ_B200_H */
