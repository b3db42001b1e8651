// stepper.cuh — sm_100a kernels for the Runge–Kutta hot path: one trajectory per thread, stage vectors and
// controller state in registers, tableau coefficients and libm constants in the constant bank.
//
// Replaces (reference file:line, /root/reference/src/ode/problem.rs):
//   oderk_fixed :361-406            -> fixed_kernel
//   hinit :850-899                  -> hinit_kernel          (pre-pass, once per trajectory)
//   oderk_adapt :193-358            -> adapt_kernel          (calc_coefficients :702-737, embedded_step :740-780,
//                                                             stepsize_hw92 :801-845, hermite_interp :783-798 inlined)
// PARITY CONTRACT: every + - * / below is one IEEE f64 operation in the reference's evaluation order (SURVEY.md
// App. A); this translation unit must be compiled with -fmad=false so nvcc never contracts, and pow/log10 come
// from deq_math.cuh (glibc's algorithm).  Zero tableau entries are multiplied through like the reference does
// (they matter for NaN/Inf propagation and the sign of zero).
//
// Execution model (B200): 148 SMs x 4 resident CTAs x 128 threads.  Trajectories of an adaptive ensemble need
// different numbers of steps, so the adaptive kernel is PERSISTENT: lanes whose trajectory finished are re-filled
// from a global work counter (one warp-aggregated atomicAdd per refill, lanes ranked with a ballot), which keeps
// the FP64 pipe busy instead of idling until the slowest lane of the warp finishes.  HBM traffic is
// (2N + NP + 1) doubles in and N doubles + 21 bytes of counters out per trajectory for ~10^3 steps: the kernel
// is bound by the FP64 pipe, not by memory.  Measured on B200 (tools/micro/): an FP64 instruction holds the SMSP issue
// port for two cycles and nothing co-issues, so kernel time ~ 2 * N_fp64 + N_other per warp-step — every integer, move
// and branch instruction counts, which is why the loop below checks for refills once per burst of attempts, keeps
// constants in the constant bank and avoids fmax/fmin's NaN plumbing where the operands cannot be NaN.
// Dense output (MODE_DENSE) writes either SoA [N][ntspan][ntraj] (lanes at the same output row store to consecutive
// addresses: parameter-sorted sweeps) or trajectory-major [ntraj][ntspan][N] staged through shared memory with
// cooperative 128-byte drains (ensembles in arbitrary order); DEQ_OUT_ALL reuses the same kernel in two passes.
#pragma once
#include "deq_rtc_compat.h"
#include "deq_math.cuh"
#include "tableaus.cuh"
#include "rhs.cuh"
#include "params.h"

namespace deqk {

__device__ const uint64_t g_exp_tab[256] = {DEQ_EXP_TAB_INIT};
__device__ const double g_powlog_tab[384] = {DEQ_POWLOG_TAB_INIT};
__device__ const double g_log_tab[256] = {DEQ_LOG_TAB_INIT};

// Tableau coefficients live in the constant bank: they reach the FP64 instructions through uniform registers loaded
// with one LDCU.128 per two doubles, whereas 64-bit immediates cost two UMOV issue slots each (r1 profile: ~100 UMOV
// per step attempt).  Not `const` on purpose, so the front end does not fold the values back into immediates.
template <class TB>
__constant__ deqt::Data<TB::S> c_tab = TB::data();
// Controller constants with non-zero low words (as immediates they cost two UMOV issue slots each, every attempt).
__constant__ double c_ctl[4] = {0.8, 0.2, 0.01, -0x1.c8ff7c79a9a20p-3 /* ln 0.8 */};
// per tableau: 1/(order_min + 1) (the controller exponent, problem.rs:830) and -1/(2 (order_min + 1)) (FAST: applied to the
// squared norm)
template <class TB>
__constant__ double c_pw[2] = {1. / (double)(TB::ORDER_MIN + 1), -0.5 / (double)(TB::ORDER_MIN + 1)};
#define DEQ_C08 (deqk::c_ctl[0])
#define DEQ_C02 (deqk::c_ctl[1])
#define DEQ_C001 (deqk::c_ctl[2])

// FAST mode: combined error weights b - b* (one FMA chain for the error estimate instead of two for p and q).
template <int S_>
struct ErrWeights { double e[S_]; };
template <class TB>
constexpr ErrWeights<TB::S> make_err_weights() {
    ErrWeights<TB::S> w{};
    constexpr auto d = TB::data();
    for (int s = 0; s < TB::S; ++s) w.e[s] = d.b[s][0] - d.b[s][1];
    return w;
}
template <class TB>
__constant__ ErrWeights<TB::S> c_errw = make_err_weights<TB>();
#ifndef DEQ_TAB_CONST
#define DEQ_TAB_CONST 1
#endif
#if DEQ_TAB_CONST
#define DEQ_TAB(TB) const auto& tb = c_tab<TB>
#else
#define DEQ_TAB(TB) constexpr auto tb = TB::data()
#endif
#ifndef DEQ_MIN_BLOCKS
#define DEQ_MIN_BLOCKS 4
#endif

template <class F, bool PT>
__device__ __forceinline__ void load_params(double* prm, const double* params, const SharedParams& sh,
                                            unsigned long long traj, unsigned long long ntraj) {
#pragma unroll
    for (int j = 0; j < F::NP; ++j) prm[j] = PT ? params[(unsigned long long)j * ntraj + traj] : sh.v[j];
}

// calc_coefficients (problem.rs:702-737): k[0] holds k0; fills k[1..S-1].
// FAST = false: the reference's operation order, one rounding per * and +, zero coefficients multiplied through.
// FAST = true : fused multiply-adds, structurally-zero coefficients skipped (results differ in the last bits).
// `ylast` (may be nullptr) receives the argument of the last stage, y + dt * sum_j a[S-1][j] k_j: for a first-same-as-last
// tableau that IS the new solution (a[S-1][j] == b[j]), so FAST mode does not evaluate the weighted sum a second time.
// Loops over the state components are fully unrolled while the stage vectors fit the register file (N <= 16); for larger
// systems they stay rolled — `#pragma unroll (N <= 16 ? N : 1)` — which puts the per-thread arrays (k[S][N], stage argument,
// trial state) into local memory and keeps the code size independent of N.
// Time-only terms of the right-hand side ("forcing": cos(Omega t) of the driven pendulum).  A functor that declares
//   static constexpr int NG;  forcing(t, g[NG], p);  eval_g(t, y, dy, p, g)     with eval(t, y, dy, p) == eval_g(.., forcing(t))
// has them evaluated once per DISTINCT stage time: rows with equal c see the same bits of t + c dt, hence the same forcing
// (Fehlberg 7(8): c3 = c7 = 1/6, c10 = c12 = 1 — and the extra right-hand side of an accepted step sits at t + dt as well),
// and a row with c = 0 sees the forcing at t, which the previous accepted step left behind (`gcur`; CARRY flavours only).
// 13 sin + 9 cos per attempt of config 4 instead of 13 + 13; same bits, because nothing is evaluated at different operands.
template <class F, class = void>
struct ng_of { static constexpr int value = 0; };
template <class F>
struct ng_of<F, decltype((void)F::NG)> { static constexpr int value = F::NG; };

template <int S_>
constexpr bool has_c_row(const deqt::Data<S_>& d, double c) {
    for (int r = 1; r < S_; ++r)
        if (d.c[r] == c) return true;
    return false;
}
// carry the forcing at the current time across attempts: only when the tableau re-visits t (a c = 0 row) and an accepted
// step's end (a c = 1 row) comes for free
template <class TB, class F>
constexpr bool carries_forcing() { return ng_of<F>::value > 0 && has_c_row(TB::data(), 0.0) && has_c_row(TB::data(), 1.0); }

template <class TB, class F, bool FAST, bool SKIPZ = false, bool CARRY = false>
__device__ __forceinline__ void stages(double t, const double (&y)[F::N], double dt, double (&k)[TB::S][F::N],
                                       const double* prm, double* ylast = nullptr, const double* gcur = nullptr,
                                       double* gend = nullptr) {
    DEQ_TAB(TB);
    constexpr auto tbc = TB::data();
    constexpr int NG = ng_of<F>::value;
    double g[TB::S][NG ? NG : 1];
#pragma unroll
    for (int row = 1; row < TB::S; ++row) {
        // w_col = a[row][col] * dt first (one product per entry, like the reference), then every component accumulates its
        // stage argument over col in order: component-outer, so that the large-system flavour (rolled component loop, stage
        // vectors in local memory) touches each element once per row
        double w[TB::S];
#pragma unroll
        for (int col = 0; col < row; ++col) w[col] = tb.a[row][col] * dt;
        double yi[F::N];
#pragma unroll (F::N <= 16 ? F::N : 1)
        for (int d = 0; d < F::N; ++d) {
            double acc = y[d];
#pragma unroll
            for (int col = 0; col < row; ++col) {
                if (FAST) { if (tbc.a[row][col] != 0.0) acc = __fma_rn(k[col][d], w[col], acc); }
                else if (!SKIPZ || tbc.a[row][col] != 0.0) acc = acc + k[col][d] * w[col];
            }
            yi[d] = acc;
        }
        const double tn = FAST ? __fma_rn(tb.c[row], dt, t) : t + tb.c[row] * dt;
        if constexpr (NG > 0) {
            int src = row;   // the first row with this c (resolved at compile time once the row loop is unrolled)
#pragma unroll
            for (int r2 = row - 1; r2 >= 1; --r2)
                if (tbc.c[r2] == tbc.c[row]) src = r2;
            if (CARRY && tbc.c[row] == 0.0) {
#pragma unroll
                for (int j = 0; j < NG; ++j) g[row][j] = gcur[j];
            } else if (src == row) {
                F::forcing(tn, g[row], prm);
            } else {
#pragma unroll
                for (int j = 0; j < NG; ++j) g[row][j] = g[src][j];
            }
            F::eval_g(tn, yi, k[row], prm, g[row]);
            if (gend && tbc.c[row] == 1.0) {
#pragma unroll
                for (int j = 0; j < NG; ++j) gend[j] = g[row][j];
            }
        } else {
            F::eval(tn, yi, k[row], prm);
        }
        if (row == TB::S - 1 && ylast) {
#pragma unroll (F::N <= 16 ? F::N : 1)
            for (int d = 0; d < F::N; ++d) ylast[d] = yi[d];
        }
    }
}

// The stages exactly as the reference writes them (every zero coefficient multiplied through), behind a call: used by the rare
// attempts in which the zero-skipping main path is not provably bit-neutral.  kbuf is [S][N] row-major, kbuf[0..N) = k0 on entry.
template <class TB, class F>
__device__ __noinline__ void stages_exact_cold(double t, const double* ybuf, double dt, double* kbuf, const double* pbuf) {
    double y[F::N], k[TB::S][F::N], prm[F::NP ? F::NP : 1];
#pragma unroll (F::N <= 16 ? F::N : 1)
    for (int d = 0; d < F::N; ++d) { y[d] = ybuf[d]; k[0][d] = kbuf[d]; }
#pragma unroll
    for (int j = 0; j < (F::NP ? F::NP : 1); ++j) prm[j] = pbuf[j];
    stages<TB, F, false, false>(t, y, dt, k, prm);
#pragma unroll
    for (int s = 1; s < TB::S; ++s) {
#pragma unroll (F::N <= 16 ? F::N : 1)
        for (int d = 0; d < F::N; ++d) kbuf[s * F::N + d] = k[s][d];
    }
}

// Keeps a rarely taken block out of line: an empty volatile asm cannot be speculated, so the compiler does not hoist the
// block's (cheap-looking, but FP64-pipe) compares above the branch that guards it.
#define DEQ_COLD_PATH() asm volatile("")

// The operand of larger magnitude, for operands that are known not to be NaN at the call site (a NaN trial state takes the
// reject branch before the error scale is used; accepted states are never NaN): |absmax_operand(a, b)| == max(|a|, |b|)
// of the reference (problem.rs:827).  Written this way the absolute values are operand modifiers of the compare and of the
// consuming multiply — fabs() of a double on its own is a DADD on sm_100a and holds the FP64 issue port for two cycles;
// fmax() would add ~7 integer/select instructions of NaN plumbing.
__device__ __forceinline__ double absmax_operand(double a, double b) { return fabs(a) > fabs(b) ? a : b; }
// sel_max(c, x) == f64::max(c, x) when c is not NaN: a NaN x yields c, like the NaN-ignoring reference call
// (problem.rs:833,836).  PTX setp/selp on purpose: from `x > c ? x : c` the compiler builds its fmax idiom (DSETP.MAX plus
// NaN-quieting selects, ~10 instructions).
__device__ __forceinline__ double sel_max(double c, double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("{ .reg .pred p; setp.gt.f64 p, %1, %2; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(x), "d"(c));
    return r;
#else
    return x > c ? x : c;
#endif
}
__device__ __forceinline__ double sel_min(double c, double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("{ .reg .pred p; setp.lt.f64 p, %1, %2; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(x), "d"(c));
    return r;
#else
    return x < c ? x : c;
#endif
}
// `if timeout > 0 { new_dt = new_dt.min(dt) }` (problem.rs:835-837) as one predicated compare-select
__device__ __forceinline__ double sel_min_if(double c, double x, unsigned cond) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("{ .reg .pred p, q; setp.ne.u32 q, %3, 0; setp.lt.and.f64 p, %1, %2, q; selp.f64 %0, %1, %2, p; }"
        : "=d"(r) : "d"(x), "d"(c), "r"(cond));
    return r;
#else
    return (cond && x < c) ? x : c;
#endif
}
__device__ __forceinline__ uint32_t hi32(double x) { return (uint32_t)__double2hiint(x); }

// dt / 100. (problem.rs:337) — bit-identical to the IEEE division, in three FP64 instructions instead of the generic
// division sequence (MUFU seed + Newton steps + range checks, ~10 FP64 + 6 other instructions per accepted step).
// q = RN(x * R), rem = x - 100 q (exact in an FMA), q' = RN(q + rem * R) with R = RN(1/100): the value before the last
// rounding is within 2^-52 ulp of x/100, while x/100 itself can never be closer than 0.02 ulp to a rounding midpoint
// (x - 100 (m + 1/2) ulp(q) is a non-zero multiple of 2 ulp(q)), so the rounding is the correct one.  Checked against
// the hardware division on 1.1e9 random operands (0 mismatches).  Outside the guarded exponent range (2^-900 <= |x| <
// 2^900, tested on the high word with integer instructions) it divides.
__device__ __forceinline__ double div100_exact(double x) {
    const double R = DEQ_C001;
    const double q = x * R;
    double r = __fma_rn(__fma_rn(-100.0, q, x), R, q);
    if (((hi32(x) & 0x7fffffffu) - 0x07b00000u) >= (0x78300000u - 0x07b00000u)) { DEQ_COLD_PATH(); r = x / 100.; }
    return r;
}

template <int S_>
constexpr int zero_entries_of_a(const deqt::Data<S_>& d) {
    int z = 0;
    for (int r = 1; r < S_; ++r)
        for (int c = 0; c < r; ++c) z += d.a[r][c] == 0.0 ? 1 : 0;
    return z;
}

template <int S_>
constexpr bool has_zero_weight(const deqt::Data<S_>& d) {
    for (int s = 0; s < S_; ++s)
        if (d.b[s][0] == 0.0 || d.b[s][1] == 0.0) return true;
    return false;
}

// embedded_step (problem.rs:740-780): p = sum_s k_s b[s][0], q = sum_s k_s b[s][1] accumulated from 0.0 in stage order,
// yerr = (p - q) dt, ytrial = y + p dt — every product and sum rounded once, like the reference.
// The reference also adds the terms whose weight is zero.  Such a term is k * 0.0 = +-0 whenever k is finite, and adding
// +-0 to p leaves every bit of p alone: p starts as +0.0 and a sum is -0.0 only if both operands are, so p is never -0.0
// (the one value that x + 0.0 would change).  So the zero-weight terms are skipped when all the stage values they would
// multiply are finite, which is decided without FP64 work: the high words of those doubles are added as floats (a double
// that is inf, NaN or >= 2^1017 has a high word that reads as a float inf / NaN, and a float sum that overflows errs on
// the safe side).  Otherwise — the trajectory is blowing up — the sums are formed exactly as written.
template <class TB, int N>
__device__ __forceinline__ void embedded_parity(const double (&k)[TB::S][N], const double (&ys)[N], double dt,
                                                double (&ytrial)[N], double (&yerr)[N]) {
    DEQ_TAB(TB);
    constexpr auto tbc = TB::data();
    constexpr int S = TB::S;
    bool skip_ok = true;
    if (has_zero_weight(tbc)) {
        float m = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (tbc.b[s][0] == 0.0 || tbc.b[s][1] == 0.0) {
#pragma unroll (N <= 16 ? N : 1)
                for (int d = 0; d < N; ++d) m = m + fabsf(__int_as_float(__double2hiint(k[s][d])));
            }
        }
        skip_ok = m < 3.0e38f;
    }
    double p[N], q[N];
    if (skip_ok) {
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) {
            p[d] = 0.0; q[d] = 0.0;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                if (tbc.b[s][0] != 0.0) p[d] = p[d] + k[s][d] * tb.b[s][0];
                if (tbc.b[s][1] != 0.0) q[d] = q[d] + k[s][d] * tb.b[s][1];
            }
        }
    } else {
        DEQ_COLD_PATH();
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) {
            p[d] = 0.0; q[d] = 0.0;
#pragma unroll
            for (int s = 0; s < S; ++s) { p[d] = p[d] + k[s][d] * tb.b[s][0]; q[d] = q[d] + k[s][d] * tb.b[s][1]; }
        }
    }
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) {
        yerr[d] = (p[d] - q[d]) * dt;
        ytrial[d] = ys[d] + (p[d] * dt);
    }
}

#ifndef DEQ_BURST
#define DEQ_BURST 8
#endif
constexpr int BURST_ATTEMPTS = DEQ_BURST;

// 1/x to ~2^-40 relative accuracy without the IEEE corner-case handling (x is a positive, normal error scale):
// MUFU.RCP64H seed + one Newton step.  FAST mode only.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = __fma_rn(-x, r, 1.0);   // seed is good to ~2^-20
    r = __fma_rn(r, e, r);             // ~2^-40
#ifdef DEQ_RCP_TWO_STEPS                // 2^-40 is far below what the step-size controller can resolve; the second step is off
    e = __fma_rn(-x, r, 1.0);
    r = __fma_rn(r, e, r);             // ~2^-80 -> rounded double
#endif
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// oderk_fixed (problem.rs:361-406): every trajectory walks the same tspan grid, dt = tspan[i+1] - tspan[i].
// ---------------------------------------------------------------------------------------------------------------
template <class TB, class F, bool PT, bool ALL>
__global__ void __launch_bounds__(BLOCK) fixed_kernel(const FixedParams P) {
    constexpr int N = F::N, S = TB::S;
    DEQ_TAB(TB);
    if (F::USES_SINCOS) { deqm::sincos_stage(); __syncthreads(); }
    const unsigned long long traj = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x;
    if (traj >= P.ntraj) return;
    double prm[F::NP ? F::NP : 1];
    load_params<F, PT>(prm, P.params, P.shared, traj, P.ntraj);
    double y[N];
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) y[d] = P.y0[(unsigned long long)d * P.ntraj + traj];
    if (ALL) {
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) P.yall[((unsigned long long)d * P.nt) * P.ntraj + traj] = y[d];
    }
    // A single problem (BASELINE config 1: one Lorenz trajectory, 10^5 steps) is ONE thread walking a dependent chain, so every
    // cycle of latency in the loop is paid in full: the grid point is loaded one step ahead (and its cache line pulled into L2
    // 64 steps ahead), and the stage terms of structurally-zero coefficients are skipped under the exactness guard of the
    // adaptive kernel (attempt_math) — RK4: 21 of 117 FP64 instructions per step.
    constexpr bool SKIPZ = zero_entries_of_a(TB::data()) >= 3;
    double tcur = P.nt ? __ldg(P.tspan) : 0.0;
    double tpre = P.nt > 1 ? __ldg(P.tspan + 1) : 0.0;
    for (unsigned long long i = 0; i + 1 < P.nt; ++i) {
        const double tnext = tpre;
        if (i + 2 < P.nt) tpre = __ldg(P.tspan + i + 2);
#if defined(__CUDA_ARCH__)
        if ((i & 15ull) == 0ull && i + 64 < P.nt) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tspan + i + 64));
#endif
        const double dt = tnext - tcur;
        double k[S][N];
        F::eval(tcur, y, k[0], prm);
        stages<TB, F, false, SKIPZ>(tcur, y, dt, k, prm);
        if (SKIPZ) {   // (see attempt_math: a state component that is -0.0, or a stage value / step that is not finite)
            float m = fabsf(__int_as_float(__double2hiint(dt)));
            bool negzero = false;
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) {
                negzero = negzero || (((unsigned)__double2hiint(y[d]) ^ 0x80000000u) | (unsigned)__double2loint(y[d])) == 0u;
#pragma unroll
                for (int s2 = 0; s2 < S; ++s2) m = m + fabsf(__int_as_float(__double2hiint(k[s2][d])));
            }
            if (!(m < 3.0e38f) || negzero) {
                DEQ_COLD_PATH();
                double kbuf[S * N], ybuf[N], pbuf[F::NP ? F::NP : 1];
#pragma unroll (N <= 16 ? N : 1)
                for (int d = 0; d < N; ++d) { ybuf[d] = y[d]; kbuf[d] = k[0][d]; }
#pragma unroll
                for (int j = 0; j < (F::NP ? F::NP : 1); ++j) pbuf[j] = prm[j];
                stages_exact_cold<TB, F>(tcur, ybuf, dt, kbuf, pbuf);
#pragma unroll
                for (int s2 = 1; s2 < S; ++s2) {
#pragma unroll (N <= 16 ? N : 1)
                    for (int d = 0; d < N; ++d) k[s2][d] = kbuf[s2 * N + d];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) y[d] = y[d] + (k[s][d] * tb.b[s][0]) * dt;
        }
        if (ALL) {
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) P.yall[((unsigned long long)d * P.nt + (i + 1)) * P.ntraj + traj] = y[d];
        }
        tcur = tnext;
    }
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) P.yfinal[(unsigned long long)d * P.ntraj + traj] = y[d];
}

// ---------------------------------------------------------------------------------------------------------------
// hinit (problem.rs:850-899), one thread per trajectory; writes f0 = f(t0,y0) and the first step h.
// ---------------------------------------------------------------------------------------------------------------
template <class F, bool PT>
__global__ void __launch_bounds__(BLOCK) hinit_kernel(const HinitParams P) {
    constexpr int N = F::N;
    __shared__ uint64_t s_exp[256];
    __shared__ double s_powlog[384];
    __shared__ double s_log[256];
    for (int i = threadIdx.x; i < 256; i += BLOCK) { s_exp[i] = g_exp_tab[i]; s_log[i] = g_log_tab[i]; }
    for (int i = threadIdx.x; i < 384; i += BLOCK) s_powlog[i] = g_powlog_tab[i];
    if (F::USES_SINCOS) deqm::sincos_stage();
    __syncthreads();
    const deqm::Tables T{s_exp, s_powlog, s_log};
    const unsigned long long traj = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x;
    if (traj >= P.ntraj) return;
    double prm[F::NP ? F::NP : 1];
    load_params<F, PT>(prm, P.params, P.shared, traj, P.ntraj);
    double x0[N], f0[N];
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) x0[d] = P.y0[(unsigned long long)d * P.ntraj + traj];
    const double t0 = P.t0, tend = P.tend, reltol = P.reltol, abstol = P.abstol;
    const double diff = tend - t0;
    const double tdir = (diff != diff) ? diff : copysign(1.0, diff);  // f64::signum
    double norm = 0.0;
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) { const double a = fabs(x0[d]); if (a > norm) norm = a; }
    const double tau = fmax(norm * reltol, 1.0 * abstol);
    const double d0 = norm / tau;
    F::eval(t0, x0, f0, prm);
    double nf0 = 0.0;
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) { const double a = fabs(f0[d]); if (a > nf0) nf0 = a; }
    const double d1 = nf0 / tau;
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1.0e-6 : 0.01 * (d0 / d1);
    double x1[N], f1[N];
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) x1[d] = x0[d] + (f0[d] * h0) * tdir;
    F::eval(t0 + tdir * h0, x1, f1, prm);
    double nf1 = 0.0;
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) { const double a = fabs(f1[d] - f0[d]); if (a > nf1) nf1 = a; }
    const double d2 = nf1 / (tau * h0);
    const double m = fmax(d1, d2);
    double h1;
    if (m < 1e-15) {
        h1 = fmax(1.0e-6, 1.0e-3 * h0);
    } else {
        const double pw = (-(2. + deqm::log10_glibc(m, T))) / (double)(P.order + 1);
        h1 = P.libm_folded ? deqm::exp10_glibc(pw, T) : deqm::pow_glibc(10.0, pw, T);
    }
    const double h = tdir * fmin(fmin(h1, 100. * h0), tdir * (tend - t0));
    P.h0[traj] = h;
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) P.f0[(unsigned long long)d * P.ntraj + traj] = f0[d];
}

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory staging of dense output (trajectory-major layout, REC_TSPAN_TM).
//
// In the SoA layout [n][ntspan][ntraj] a warp's stores coalesce only while its lanes sit at the same output row — true for
// a parameter-sorted sweep (8.4 sectors per request measured on config 3), false for an arbitrary ensemble, where every
// lane writes its own row and a store instruction degenerates into 32 single-sector transactions (measured 6.4x slower
// on the permuted sweep).  Here every lane appends its rows to a private 32-double ring in shared memory
// (ring[col][lane ^ col]: conflict-free for the owner's appends and for the cooperative reads) and the lanes that are
// currently converged drain full 128-byte lines together: one coalesced store per 16 doubles instead of 16 scattered
// ones.  Output layout: out[traj][row][component], i.e. one contiguous `Vec<Y>`-like stream per trajectory.
// ---------------------------------------------------------------------------------------------------------------
#ifndef DEQ_STAGE_COLS
#define DEQ_STAGE_COLS 32
#endif
constexpr int STAGE_COLS = DEQ_STAGE_COLS;       // doubles per lane in the ring (a power of two)
constexpr int STAGE_LINE = DEQ_STAGE_COLS / 2;   // doubles drained at a time (16 doubles = one 128-byte line)
constexpr unsigned STAGE_BYTES = STAGE_COLS * BLOCK * sizeof(double);   // dynamic shared memory, requested only by
                                                                        // launches that use the staged layout
extern __shared__ double s_stage_dyn[];

struct StageLane {
    unsigned head = 0, cnt = 0;      // ring position of the oldest element, number of buffered doubles
    unsigned long long gpos = 0;     // global index (in doubles) the oldest element goes to
};

__device__ __forceinline__ double& stage_ref(double* stage, unsigned col, unsigned lane) {
    return stage[col * BLOCK + (threadIdx.x & ~31u) + (lane ^ (col & 31u))];
}

// Drain one STAGE_LINE of every participating lane whose ring holds at least that much (all == false), or everything that
// is left (all == true, used when a trajectory finishes).  The lanes that are converged here split the work items evenly:
// item i = element (i % LINE) of the (i / LINE)-th pending lane, so 32 lanes write two full lines per iteration with three
// shuffles and no vote inside the loop.
__device__ __forceinline__ void stage_drain(double* stage, double* out, StageLane& s, bool all) {
    const unsigned am = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp(am);  // the owners' appends are visible to the helpers
    const unsigned rank = __popc(am & ((1u << lane) - 1u)), m = __popc(am);
    if (!all) {
        unsigned pending = __ballot_sync(am, s.cnt >= (unsigned)STAGE_LINE);
        if (pending == 0u) return;
        if (am == 0xffffffffu && STAGE_LINE == 16) {
            // the whole warp is here (the steady state of the persistent loop): two pending lanes per iteration, one per half
            // warp, picked with warp-uniform bit scans — lane e of a half writes element e of its pending lane's line.  (The
            // general form below finds the n-th pending lane per element with __fns: ~40 instructions per 32 elements, against
            // ~15 here; ncu on config 3 showed the drain as the larger part of the 500 extra instructions per warp-step of the
            // trajectory-major layout.)
            const unsigned all_pending = pending;
            const unsigned e = lane & 15u;
            while (pending) {
                const int src0 = __ffs(pending) - 1;
                pending &= pending - 1u;
                const int src1 = pending ? __ffs(pending) - 1 : -1;
                if (pending) pending &= pending - 1u;
                const int src = lane < 16u ? src0 : src1;
                const unsigned h = __shfl_sync(0xffffffffu, s.head, src < 0 ? 0 : src);
                const unsigned long long g = __shfl_sync(0xffffffffu, s.gpos, src < 0 ? 0 : src);
                if (src >= 0) out[g + e] = stage_ref(stage, (h + e) & (STAGE_COLS - 1), (unsigned)src);
            }
            if ((all_pending >> lane) & 1u) { s.head = (s.head + STAGE_LINE) & (STAGE_COLS - 1); s.cnt -= STAGE_LINE; s.gpos += STAGE_LINE; }
            __syncwarp(am);
            return;
        }
        const unsigned W = (unsigned)STAGE_LINE * __popc(pending);
        for (unsigned base = 0; base < W; base += m) {  // uniform trip count over the participating lanes
            const unsigned i = base + rank;
            const bool has = i < W;
            const unsigned src = has ? __fns(pending, 0, (int)(i / STAGE_LINE) + 1) : lane;
            const unsigned h = __shfl_sync(am, s.head, src);
            const unsigned long long g = __shfl_sync(am, s.gpos, src);
            const unsigned e = i % STAGE_LINE;
            if (has) out[g + e] = stage_ref(stage, (h + e) & (STAGE_COLS - 1), src);
        }
        if ((pending >> lane) & 1u) { s.head = (s.head + STAGE_LINE) & (STAGE_COLS - 1); s.cnt -= STAGE_LINE; s.gpos += STAGE_LINE; }
    } else {
        unsigned pending = __ballot_sync(am, s.cnt > 0u);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1u;
            const unsigned h = __shfl_sync(am, s.head, src);
            const unsigned c = __shfl_sync(am, s.cnt, src);
            const unsigned long long g = __shfl_sync(am, s.gpos, src);
            for (unsigned e = rank; e < c; e += m) out[g + e] = stage_ref(stage, (h + e) & (STAGE_COLS - 1), (unsigned)src);
            if ((int)lane == src) { s.head = (s.head + c) & (STAGE_COLS - 1); s.cnt = 0; s.gpos += c; }
        }
    }
    __syncwarp(am);  // helpers are done reading before the owners append again
}

__device__ __forceinline__ void stage_append(double* stage, double* out, StageLane& s, double v) {
    const unsigned lane = threadIdx.x & 31u;
    if (s.cnt == (unsigned)STAGE_COLS) {  // ring full inside one step (many tspan points per step): the owner drains a line alone
        for (unsigned e = 0; e < (unsigned)STAGE_LINE; ++e) out[s.gpos + e] = stage_ref(stage, (s.head + e) & (STAGE_COLS - 1), lane);
        s.head = (s.head + STAGE_LINE) & (STAGE_COLS - 1); s.cnt -= STAGE_LINE; s.gpos += STAGE_LINE;
    }
    stage_ref(stage, (s.head + s.cnt) & (STAGE_COLS - 1), lane) = v;
    ++s.cnt;
}

// ---------------------------------------------------------------------------------------------------------------
// The arithmetic of ONE step attempt (problem.rs:261-283: calc_coefficients, embedded_step, stepsize_hw92), shared by the
// persistent-lane kernel and the row-synchronous dense kernel: from (t, dt, cy, ck) to the stage values k, the trial state,
// the proposed next step `ndt` and the accept decision `accu` (1 / 0).  `yl` is read by the Points::Specified flavour only.
// ---------------------------------------------------------------------------------------------------------------
template <class TB, class F, bool FAST, bool SPEC, class TT>
__device__ __forceinline__ void attempt_math(const AdaptParams& P, const TT& T, double t, double dt, const double (&cy)[F::N],
                                             const double (&ck)[F::N], const double (&yl)[F::N], const double* prm, unsigned timeout,
                                             double tdir, double reltol, double abstol, double (&k)[TB::S][F::N],
                                             double (&ytrial)[F::N], double& ndt, unsigned& accu,
                                             const double* cg = nullptr, double* gend = nullptr) {
    constexpr int N = F::N, S = TB::S;
    constexpr bool FSAL = deqt::fsal_of(TB::data());
    constexpr bool CARRY = carries_forcing<TB, F>();   // cg: forcing at t (in), gend: forcing at t + dt (out)
    DEQ_TAB(TB);
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) k[0][d] = ck[d];
    double yerr[N];
    if (!FAST) {
        // The reference also adds the stage terms whose coefficient a[row][col] is zero: k * (0 * dt) = +-0 for finite
        // k and dt, and x + (+-0) == x bit for bit unless x is -0.0 — which a partial sum y + ... can only be if y
        // itself is -0.0 (a sum is -0.0 only when both operands are).  So for tableaus with many structural zeros
        // (Fehlberg 7(8): 32 of 78 entries, 160 FP64 instructions per attempt for n = 2) the stages run without those
        // terms, and are re-run exactly as written in the rare attempt where a state component is -0.0 or a stage
        // value / the step is not finite (decided on high words: float adds and integer compares, no FP64 work).
        constexpr bool SKIPZ = zero_entries_of_a(TB::data()) >= 4;
        // (the carried forcing stands for the one at t + 0 * dt: the same bits unless dt is not finite or t is -0.0, which the
        // guard below sends to the exact path as well — hence carried in the guarded flavour only)
        stages<TB, F, false, SKIPZ, SKIPZ && CARRY>(t, cy, dt, k, prm, nullptr, cg, gend);
        if (SKIPZ) {
            float m = fabsf(__int_as_float(__double2hiint(dt)));
            bool negzero = CARRY && (((unsigned)__double2hiint(t) ^ 0x80000000u) | (unsigned)__double2loint(t)) == 0u;
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) {
                negzero = negzero || (((unsigned)__double2hiint(cy[d]) ^ 0x80000000u) | (unsigned)__double2loint(cy[d])) == 0u;
#pragma unroll
                for (int s2 = 0; s2 < S; ++s2) m = m + fabsf(__int_as_float(__double2hiint(k[s2][d])));
            }
            if (!(m < 3.0e38f) || negzero) {   // out of line: a second inlined copy of the stage code costs registers on the main path
                DEQ_COLD_PATH();
                double kbuf[S * N], ybuf[N], pbuf[F::NP ? F::NP : 1];
#pragma unroll (N <= 16 ? N : 1)
                for (int d = 0; d < N; ++d) { ybuf[d] = cy[d]; kbuf[d] = k[0][d]; }
#pragma unroll
                for (int j = 0; j < (F::NP ? F::NP : 1); ++j) pbuf[j] = prm[j];
                stages_exact_cold<TB, F>(t, ybuf, dt, kbuf, pbuf);
#pragma unroll
                for (int s2 = 1; s2 < S; ++s2) {
#pragma unroll (N <= 16 ? N : 1)
                    for (int d = 0; d < N; ++d) k[s2][d] = kbuf[s2 * N + d];
                }
            }
        }
        // embedded_step :761-771  (y == cy under Points::All; the last output row under Points::Specified)
        double ys[N];
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) ys[d] = SPEC ? yl[d] : cy[d];
        embedded_parity<TB, N>(k, ys, dt, ytrial, yerr);
        // stepsize_hw92 :801-845.  The reference first scans the trial state for NaN (-> err = 10, dt *= 0.2,
        // timeout = 5).  A NaN trial component makes the scaled error sum NaN as well, so the scan only has to
        // run when the sum is NaN — one test on the hot path instead of N.
        // Main path: every division, the reciprocal and both pow calls by their straight-line instruction sequences
        // (deq_math.cuh: div_main, rcp_main, pow_glibc_main), the out-of-range conditions OR-ed into ONE flag; if it
        // is set — or the error sum is inf / NaN — the block below re-evaluates the controller with the general
        // operators and routines.  Same bits either way; eight branch regions fewer per attempt on the main path.
        bool redo = false;
        double ssum = 0.0;
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) {
            const double e = deqm::div_main(yerr[d], fabs(absmax_operand(ys[d], ytrial[d])) * reltol + abstol, redo);
            const double a = fabs(e);
            ssum = ssum + a * a;
        }
        double err = P.libm_folded ? sqrt(ssum) : deqm::pow_glibc_main(ssum, 0.5, T, redo);
        double g = deqm::pow_glibc_main(deqm::rcp_main(err, redo), c_pw<TB>[0], T, redo) * DEQ_C08;
        redo = redo || hi32(ssum) >= 0x7ff00000u;
        bool isnan_trial = false;
        if (redo) {
            DEQ_COLD_PATH();
            ssum = 0.0;
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) {
                const double e = yerr[d] / (fabs(absmax_operand(ys[d], ytrial[d])) * reltol + abstol);
                const double a = fabs(e);
                ssum = ssum + a * a;
            }
            // the reference first scans the trial state for NaN (stepsize_hw92 :814-822): err = 10, dt *= 0.2, timeout = 5
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) isnan_trial = isnan_trial || (ytrial[d] != ytrial[d]);
            err = P.libm_folded ? sqrt(ssum) : deqm::pow_glibc_cold(ssum, 0.5, T);
            g = deqm::pow_glibc_cold(1.0 / err, c_pw<TB>[0], T) * DEQ_C08;
        }
        double nd = sel_min(P.maxstep, (sel_max(DEQ_C02, g) * tdir) * dt);
        nd = sel_min_if(nd, dt, timeout);
        ndt = tdir * nd;
        if (isnan_trial) { DEQ_COLD_PATH(); err = 10.; ndt = 0.2 * dt; }   // (timeout = 5 follows from the reject)
        accu = err < 1. ? 1u : 0u;
    } else {
        // FAST: same formulas with fused multiply-adds; zero weights skipped; error weights b - b* combined;
        // err < 1 decided on the squared norm; one pow for the step factor: 0.8 (1/err)^PW = exp(ln 0.8 - PW/2 ln ssum).
        // For a first-same-as-last tableau the argument of the last stage is the new solution itself.
        constexpr auto tbc = TB::data();
        stages<TB, F, true, false, CARRY>(t, cy, dt, k, prm, FSAL ? ytrial : nullptr, cg, gend);
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) {
            double p = 0.0, e = 0.0;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                if (!FSAL && tbc.b[s][0] != 0.0) p = __fma_rn(k[s][d], tb.b[s][0], p);
                if (tbc.b[s][0] - tbc.b[s][1] != 0.0) e = __fma_rn(k[s][d], c_errw<TB>.e[s], e);
            }
            yerr[d] = e * dt;
            if (!FSAL) ytrial[d] = __fma_rn(p, dt, cy[d]);
        }
        double ssum = 0.0;
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) {
            const double e = yerr[d] * rcp_fast(__fma_rn(fabs(absmax_operand(cy[d], ytrial[d])), reltol, abstol));
            ssum = __fma_rn(e, e, ssum);
        }
        const double g = deqm::pow_fast_pos(ssum, c_pw<TB>[1], T, c_ctl[3]);
        double nd = sel_min(P.maxstep, (sel_max(DEQ_C02, g) * tdir) * dt);
        nd = sel_min_if(nd, dt, timeout);
        ndt = tdir * nd;
        if (hi32(ssum) >= 0x7ff00000u) { DEQ_COLD_PATH(); if (ssum != ssum) ndt = 0.2 * dt; }
        accu = ssum < 1. ? 1u : 0u;   // sqrt(ssum) < 1  <=>  ssum < 1; false for NaN
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tail compaction: census of a CTA's live trajectories and, when they fit into fewer warps than currently hold them, a
// repack through shared memory into the lowest lanes of the CTA.  Out of line on purpose: with the barriers inlined into the
// persistent loop, ptxas hoists tableau constants into general registers and moves them back with R2UR in every attempt.
// All warps of the CTA call it together (converged, at the top of the persistent loop).  `cnt` is one of two census
// buffers, used alternately, so that a warp that is already in the next census cannot overwrite counts that a slower warp
// is still reading.
// ---------------------------------------------------------------------------------------------------------------
template <int SLOTS>
struct TailPack { double v[SLOTS]; unsigned active, total; };

template <int SLOTS>
__device__ __noinline__ TailPack<SLOTS> tail_census(TailPack<SLOTS> pk, double (*s_pack)[BLOCK], unsigned* cnt) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const unsigned am = __ballot_sync(0xffffffffu, pk.active != 0u);
    if (lane == 0) cnt[wid] = __popc(am);
    __syncthreads();
    unsigned total = 0u, below = 0u, holders = 0u;
#pragma unroll
    for (unsigned w = 0; w < BLOCK / 32; ++w) {
        const unsigned c = cnt[w];
        total += c; holders += c ? 1u : 0u; below += w < wid ? c : 0u;
    }
    pk.total = total;
    if (total != 0u && (total + 31u) / 32u < holders) {   // uniform over the CTA
        if (pk.active) {
            const unsigned slot = below + __popc(am & ((1u << lane) - 1u));
#pragma unroll
            for (int f = 0; f < SLOTS; ++f) s_pack[f][slot] = pk.v[f];
        }
        __syncthreads();
        pk.active = threadIdx.x < total ? 1u : 0u;
        if (pk.active) {
#pragma unroll
            for (int f = 0; f < SLOTS; ++f) pk.v[f] = s_pack[f][threadIdx.x];
        }
        __syncthreads();   // every slot is read before a later repack overwrites it
    }
    return pk;
}

// ---------------------------------------------------------------------------------------------------------------
// oderk_adapt (problem.rs:193-358), Points::All semantics.
// ---------------------------------------------------------------------------------------------------------------
// SPEC = Points::Specified (problem.rs:284-300), reproduced with its defect: the step starts from `y` re-read from the OUTPUT
// vector (:262) — the last row pushed, which only moves when a tspan point is emitted — while the stages start from the true
// state (:261).  Instantiated for MODE_REC only (every row of the stream is a tspan point; the counting pass alone serves
// a final-state request), PARITY arithmetic only.
template <class TB, class F, bool PT, int MODE, bool FAST, bool SPEC = false>
__global__ void __launch_bounds__(BLOCK, DEQ_MIN_BLOCKS) adapt_kernel(const AdaptParams P) {
    static_assert(!SPEC || (MODE == MODE_REC && !FAST), "Points::Specified: MODE_REC, parity arithmetic");
    constexpr int N = F::N, S = TB::S;
    constexpr bool FSAL = deqt::fsal_of(TB::data());
    // what this instantiation records besides the final state (compile time, so the hot loop carries no mode tests):
    constexpr bool SOA = MODE == MODE_DENSE;      // tspan rows, [N][nt][ntraj]
    constexpr bool TM = MODE == MODE_DENSE_TM;    // tspan rows, [ntraj][nt][N], staged through shared memory
    constexpr bool REC = MODE == MODE_REC;        // ragged Points::All stream, counting pass or writing pass (P.rec_mode)
    constexpr bool ROWS = MODE != MODE_FINAL;     // the tspan cursor `it` is tracked

    __shared__ uint64_t s_exp[256];
    __shared__ double s_powlog[384];
    for (int i = threadIdx.x; i < 256; i += BLOCK) s_exp[i] = g_exp_tab[i];
    for (int i = threadIdx.x; i < 384; i += BLOCK) s_powlog[i] = g_powlog_tab[i];
    if (F::USES_SINCOS) deqm::sincos_stage();
    // ---- tail compaction (the "periodic trajectory compaction" of the design) ----
    // Once the work pool is empty a warp can only lose lanes: it keeps issuing full-width instructions for ever fewer live
    // trajectories until its slowest lane is done (measured: 3.4 % of all warp-steps at 2^20 trajectories, and most of the run
    // when the ensemble is only a wave or two).  From then on the CTA takes a census every TAIL_BURSTS bursts and, whenever
    // its live trajectories fit into fewer warps than currently hold them, moves their state through shared memory into
    // the lowest lanes of the CTA; emptied warps then only wait at the census barrier, and the issue slots they
    // occupied go to the warps that still have work.  Scheduling only: a trajectory's arithmetic does not depend on the lane
    // that runs it.  Final-state and SoA-dense flavours (the staged layouts keep per-lane rings in shared memory).
    constexpr int NG = ng_of<F>::value;
    constexpr bool CARRY = carries_forcing<TB, F>();   // forcing at the current time, kept across attempts (see stages)
    constexpr bool GEND = NG > 0 && has_c_row(TB::data(), 1.0);   // the stages leave the forcing at t + dt behind
    constexpr int PACK_SLOTS = 2 + 2 * N + (PT ? F::NP : 0) + 3 + (ROWS ? 1 : 0) + (CARRY ? NG : 0);
    constexpr bool CAN_COMPACT = (MODE == MODE_FINAL || MODE == MODE_DENSE) && !SPEC && PACK_SLOTS <= 24;
    constexpr int TAIL_BURSTS = 4;
    // attempts between two refill checks: a check costs ~25 issue slots, an attempt of a 13-stage tableau with a transcendental
    // right-hand side ~6000 — and such runs are short (config 4: ~75 attempts per trajectory), so idling up to BURST - 1 attempts
    // after a trajectory ends would cost several per cent there
    constexpr int BURST = S >= 10 ? 2 : BURST_ATTEMPTS;
    __shared__ double s_pack[CAN_COMPACT ? PACK_SLOTS : 1][CAN_COMPACT ? BLOCK : 1];
    __shared__ unsigned s_cnt[2][BLOCK / 32];
    __shared__ unsigned s_tail;
    if (threadIdx.x == 0) s_tail = 0u;
    __syncthreads();
    const bool compacting = CAN_COMPACT && P.refill == 2;
    unsigned tail_mode = 0u, census = 0u;
    // Shared-window base addresses, computed once and made opaque (volatile asm) so the compiler keeps them in registers
    // instead of re-deriving them from SR_CgaCtaId in front of every table lookup.
    unsigned exp_base = (unsigned)__cvta_generic_to_shared(s_exp), powlog_base = (unsigned)__cvta_generic_to_shared(s_powlog);
    asm volatile("mov.u32 %0, %0;" : "+r"(exp_base));
    asm volatile("mov.u32 %0, %0;" : "+r"(powlog_base));
    const deqm::SmemTables T{exp_base, powlog_base, 0u};

    const unsigned lane = threadIdx.x & 31u;
    const double tdir = P.tdir, tend = P.tend, reltol = P.reltol, abstol = P.abstol;
    const uint32_t max_steps = P.max_steps == 0ull || P.max_steps > 0xffffffffull ? 0xffffffffu : (uint32_t)P.max_steps;

    // 32-bit flags on purpose: the compiler packs `bool` locals into byte lanes of one register and pays a PRMT for every
    // read and write of them inside the attempt loop
    unsigned active = 0u, pool_empty = 0u;
    unsigned long long traj = 0;
    double t = 0.0, dt = 0.0;
    double cy[N], ck[N], prm[F::NP ? F::NP : 1];
    double yl[N];   // SPEC only: ys[ys.len() - 1], the last row of the output vector (dead code otherwise)
    double cg[NG ? NG : 1];   // CARRY only: forcing at t
#pragma unroll
    for (int j = 0; j < (NG ? NG : 1); ++j) cg[j] = 0.0;
    unsigned timeout = 0;
    unsigned last_step = 0u;
    uint32_t acc = 0, rej = 0, attempts = 0;
    unsigned long long it = 1, wpos = 0, wbase = 0;
    double* const stage = s_stage_dyn;
    StageLane sl;
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) { cy[d] = 0.0; ck[d] = 0.0; }
#pragma unroll (N <= 16 ? N : 1)
    for (int d = 0; d < N; ++d) yl[d] = 0.0;
#pragma unroll
    for (int j = 0; j < (F::NP ? F::NP : 1); ++j) prm[j] = 0.0;

    // ---- a burst of step attempts without any warp-level synchronisation: the refill check above costs ~20
    // issue slots, which at one check per attempt was ~3 % of the kernel (FP64 instructions hold the issue port
    // for two cycles on B200 and nothing co-issues, so every non-FP64 instruction is paid in full).  A lane that
    // finishes inside a burst idles for at most BURST-1 attempts out of the ~10^3 its trajectory took. ----
    // (a lambda, expanded once in the persistent loop and once in the tail loop below)
    auto run_bursts = [&](const int nburst) {
#pragma unroll 1
        for (int burst = 0; burst < nburst && active; ++burst) {
            if (TM) stage_drain(stage, P.dense, sl, false);
            if (REC && P.rec_mode == REC_WRITE) stage_drain(stage, P.rec_rows, sl, false);
            // one step attempt (loop body of problem.rs:260-352).  Everything after the stage evaluations is straight-line,
            // predicated code: the lanes of a warp accept and reject independently, so a branchy accept / reject tail is executed
            // on both sides by practically every warp anyway and only adds branch and convergence-barrier instructions.
            // The max_steps guard is tested after the attempt (a trajectory that has used its budget and did not stop would
            // be stopped at the top of the next iteration with exactly this state).
            uint8_t fin = 0xff;  // 0xff: keep going
            {
                ++attempts;
                double k[S][N], ytrial[N];
                double ndt;
                unsigned accu;   // the accept decision as an opaque integer: ONE FP64 compare, every later test is an integer one
                const unsigned timeout_next = timeout - (timeout > 0u ? 1u : 0u);   // stepsize_hw92 :835-838
                double gend[NG ? NG : 1];
                attempt_math<TB, F, FAST, SPEC>(P, T, t, dt, cy, ck, yl, prm, timeout, tdir, reltol, abstol, k, ytrial, ndt, accu,
                                                cg, GEND ? gend : nullptr);
                asm("" : "+r"(accu));
                const bool accept = accu != 0u;
                const double tn = t + dt;   // (the row-emitting flavours need it before the state update)
                if (!FSAL || ROWS) {
                    // the parts of an accepted step that are real work: the extra right-hand side of a non-FSAL tableau
                    // and the output rows.  Branchy on purpose; absent from the FSAL / final-state instantiations.
                    if (accept) {
                        if (!FSAL) {   // f1 overwrites the last stage (dead after the error estimate)
                            if constexpr (GEND) F::eval_g(tn, ytrial, k[S - 1], prm, gend);
                            else F::eval(tn, ytrial, k[S - 1], prm);
                        }
                        double (&f1)[N] = k[S - 1];
                        if (SPEC) {
                            // Points::Specified :284-300 — every tspan point before t+dt, all remaining ones on the last step;
                            // Hermite from the step's (lagging) y, each emitted row becomes the new last output row
                            double y0s[N];
#pragma unroll (N <= 16 ? N : 1)
                            for (int d = 0; d < N; ++d) y0s[d] = yl[d];
                            while (it < P.nt) {
                                const double tq = __ldg(P.tspan + it);
                                if (!(tdir * tq < tdir * tn || last_step != 0u)) break;
                                const double theta = (tq - t) / dt;
                                if (P.rec_mode == REC_WRITE) stage_append(stage, P.rec_rows, sl, tq);
#pragma unroll (N <= 16 ? N : 1)
                                for (int d = 0; d < N; ++d) {
                                    const double v = (y0s[d] * (1. - theta) + ytrial[d] * theta) +
                                                     ((ytrial[d] - y0s[d]) * (1. - 2. * theta) + k[0][d] * (theta - 1.) * dt +
                                                      f1[d] * theta * dt) * theta * (theta - 1.);
                                    if (P.rec_mode == REC_WRITE) stage_append(stage, P.rec_rows, sl, v);
                                    yl[d] = v;
                                }
                                ++wpos;
                                ++it;
                            }
                        } else if (ROWS) {
                            // Points::All :303-319 — Hermite values at tspan points strictly inside (t, t+dt)
                            while (it < P.nt) {
                                const double tq = __ldg(P.tspan + it);
                                if (!(tdir * t < tdir * tq && tdir * tq < tdir * tn)) break;
                                const double theta = (tq - t) / dt;
                                if (REC && P.rec_mode == REC_WRITE) stage_append(stage, P.rec_rows, sl, tq);
#pragma unroll (N <= 16 ? N : 1)
                                for (int d = 0; d < N; ++d) {
                                    const double v = (cy[d] * (1. - theta) + ytrial[d] * theta) +
                                                     ((ytrial[d] - cy[d]) * (1. - 2. * theta) + k[0][d] * (theta - 1.) * dt +
                                                      f1[d] * theta * dt) * theta * (theta - 1.);
                                    if (SOA) P.dense[((unsigned long long)d * P.nt + it) * P.ntraj + traj] = v;
                                    else if (TM) stage_append(stage, P.dense, sl, v);
                                    else if (REC && P.rec_mode == REC_WRITE) stage_append(stage, P.rec_rows, sl, v);
                                }
                                if (REC) ++wpos;
                                ++it;
                            }
                            if (REC) {  // "also store every step taken" (problem.rs:320-322)
                                if (P.rec_mode == REC_WRITE) {
                                    stage_append(stage, P.rec_rows, sl, tn);
#pragma unroll (N <= 16 ? N : 1)
                                    for (int d = 0; d < N; ++d) stage_append(stage, P.rec_rows, sl, ytrial[d]);
                                }
                                ++wpos;
                            }
                        }
                    }
                }
                // ---- state update (problem.rs:324-351), predicated ----
                if (accept) {
#pragma unroll (N <= 16 ? N : 1)
                    for (int d = 0; d < N; ++d) { ck[d] = k[S - 1][d]; cy[d] = ytrial[d]; }
                    if constexpr (CARRY) {
#pragma unroll
                        for (int j = 0; j < NG; ++j) cg[j] = gend[j];
                    }
                }
                // end snap of an accepted, not-yet-last step: `if tdir*((t+dt) + dt/100) >= tdir*tend { dt = tend - t; last = true }`
                // with the NEW t and dt (:336-340); evaluated for every lane, used by the accepting ones.  A trajectory stops
                // on an accepted last step (DONE) or on a rejected step whose successor is below minstep (:342-344); whatever
                // else is written to its step-size state in that attempt is dead.
                const double dt100 = FAST ? ndt * DEQ_C001 : div100_exact(ndt);
                if (accept) t = tn;
                const bool snapped = accept && (tdir * ((t + ndt) + dt100) >= P.tdir_tend);
                const bool stop = accept ? last_step != 0u : (fabs(ndt) < P.minstep);
                dt = snapped ? tend - t : ndt;
                last_step = snapped ? 1u : 0u;
                timeout = accept ? timeout_next : 5u;
                acc += accept ? 1u : 0u;
                rej += (accept || stop) ? 0u : 1u;
                if (stop || attempts >= max_steps) fin = !stop ? (uint8_t)ST_MAX_STEPS : accept ? (uint8_t)ST_DONE : (uint8_t)ST_MINSTEP_BREAK;
            }
            if (fin != 0xff) {
                const uint32_t nfe = 2u + attempts * (uint32_t)(S - 1) + (FSAL ? 0u : acc);
                unsigned long long dst = traj;   // where this work item's results go (AdaptParams::order); its SoA dense rows stay in column `traj`
                if ((SOA || TM) && P.order) dst = P.order[traj];
#pragma unroll (N <= 16 ? N : 1)
                for (int d = 0; d < N; ++d) P.yfinal[(unsigned long long)d * P.ntraj + dst] = SPEC ? yl[d] : cy[d];
                P.accepted[dst] = acc; P.rejected[dst] = rej; P.nfe[dst] = nfe; P.status[dst] = fin; P.t_reached[dst] = t;
                if (ROWS) {
                    if (SOA) {
#pragma unroll (N <= 16 ? N : 1)
                        for (int d = 0; d < N; ++d) P.dense[((unsigned long long)d * P.nt + (P.nt - 1)) * P.ntraj + traj] = cy[d];
                    } else if (TM) {
                        stage_drain(stage, P.dense, sl, true);   // ends with a __syncwarp: the helpers' stores precede the row below
#pragma unroll (N <= 16 ? N : 1)
                        for (int d = 0; d < N; ++d) P.dense[(dst * P.nt + (P.nt - 1)) * (unsigned long long)N + d] = cy[d];
                    } else if (REC && P.rec_mode == REC_COUNT) {
                        P.rec_counts[traj] = (uint32_t)(wpos - wbase);
                    } else if (REC) {
                        stage_drain(stage, P.rec_rows, sl, true);
                    }
                    P.nvalid[dst] = (uint32_t)it;
                }
                active = 0u;
            }
        }
    };

    for (;;) {
        // ---- refill idle lanes from the global work counter (warp-aggregated) ----
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !pool_empty && (P.refill || idle == 0xffffffffu)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.work_counter, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned long long)__popc(idle) >= P.ntraj) pool_empty = 1u;
            if (!active) {
                const unsigned long long my = base + (unsigned long long)__popc(idle & ((1u << lane) - 1u));
                if (my < P.ntraj) {
                    traj = my;
                    active = 1u;
                    unsigned long long src = my;   // where this work item's inputs live (AdaptParams::order)
                    if ((SOA || TM) && P.order) src = P.order[my];
                    load_params<F, PT>(prm, P.params, P.shared, src, P.ntraj);
#pragma unroll (N <= 16 ? N : 1)
                    for (int d = 0; d < N; ++d) {
                        cy[d] = P.y0[(unsigned long long)d * P.ntraj + src];
                        ck[d] = P.f0[(unsigned long long)d * P.ntraj + src];
                        if (SPEC) yl[d] = cy[d];
                    }
                    t = P.t0;
                    if constexpr (CARRY) F::forcing(t, cg, prm);
                    dt = P.use_initstep ? P.initstep : P.h0[src];
                    timeout = 0;
                    last_step = fabs((t + dt) - tend) <= DBL_EPSILON ? 1u : 0u;
                    acc = 0; rej = 0; attempts = 0; it = 1;
                    if (ROWS) {
                        if (SOA) {
#pragma unroll (N <= 16 ? N : 1)
                            for (int d = 0; d < N; ++d) P.dense[((unsigned long long)d * P.nt) * P.ntraj + traj] = cy[d];
                        } else if (TM) {  // staged trajectory-major stream: row 0 = y0
                            sl.head = 0; sl.cnt = 0; sl.gpos = src * P.nt * (unsigned long long)N;   // (the stream of trajectory order[s]: no column permutation afterwards)
#pragma unroll (N <= 16 ? N : 1)
                            for (int d = 0; d < N; ++d) stage_append(stage, P.dense, sl, cy[d]);
                        } else {  // ragged Points::All stream, rows [t, y...] interleaved and staged: row 0 = (t0, y0)
                            wpos = 0ull; wbase = 0ull;
                            if (P.rec_mode == REC_WRITE) {
                                sl.head = 0; sl.cnt = 0; sl.gpos = P.rec_offsets[traj] * (unsigned long long)(N + 1);
                                stage_append(stage, P.rec_rows, sl, t);
#pragma unroll (N <= 16 ? N : 1)
                                for (int d = 0; d < N; ++d) stage_append(stage, P.rec_rows, sl, cy[d]);
                            }
                            ++wpos;
                        }
                    }
                }
            }
        }
        if (compacting) {
            // the pool is known to be empty by some warp of this CTA: every warp moves on to the tail loop
            if (pool_empty) *(volatile unsigned*)&s_tail = 1u;
            if (*(volatile unsigned*)&s_tail) { tail_mode = 1u; break; }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) break;
        run_bursts(BURST);
    }
    if constexpr (CAN_COMPACT) {
      if (tail_mode) {
        // a lane that never ran a trajectory has not loaded the ensemble-wide parameters yet; it may receive one now
        if (!PT) load_params<F, PT>(prm, P.params, P.shared, 0ull, P.ntraj);
        for (;;) {
            // ---- census of the CTA's live trajectories; repack when they fit into fewer warps (tail_census, out of line) ----
            TailPack<PACK_SLOTS> pk;
            {
                int f = 0;
                pk.v[f++] = t; pk.v[f++] = dt;
#pragma unroll (N <= 16 ? N : 1)
                for (int d = 0; d < N; ++d) { pk.v[f++] = cy[d]; pk.v[f++] = ck[d]; }
                if (PT) {
#pragma unroll
                    for (int j = 0; j < F::NP; ++j) pk.v[f++] = prm[j];
                }
                pk.v[f++] = __longlong_as_double((long long)traj);
                pk.v[f++] = __hiloint2double((int)acc, (int)rej);
                pk.v[f++] = __hiloint2double((int)attempts, (int)(timeout | (last_step << 16)));
                if (ROWS) pk.v[f++] = __longlong_as_double((long long)it);
                if (CARRY) {
#pragma unroll
                    for (int j = 0; j < NG; ++j) pk.v[f++] = cg[j];
                }
                pk.active = active;
            }
            pk = tail_census(pk, s_pack, s_cnt[census & 1u]);
            ++census;
            if (pk.total == 0u) break;                    // uniform over the CTA: every warp leaves together
            {
                int f = 0;
                t = pk.v[f++]; dt = pk.v[f++];
#pragma unroll (N <= 16 ? N : 1)
                for (int d = 0; d < N; ++d) { cy[d] = pk.v[f++]; ck[d] = pk.v[f++]; }
                if (PT) {
#pragma unroll
                    for (int j = 0; j < F::NP; ++j) prm[j] = pk.v[f++];
                }
                traj = (unsigned long long)__double_as_longlong(pk.v[f++]);
                const double ar = pk.v[f++], at = pk.v[f++];
                acc = (uint32_t)__double2hiint(ar); rej = (uint32_t)__double2loint(ar);
                attempts = (uint32_t)__double2hiint(at);
                timeout = (unsigned)__double2loint(at) & 0xffffu; last_step = ((unsigned)__double2loint(at) >> 16) & 1u;
                if (ROWS) it = (unsigned long long)__double_as_longlong(pk.v[f++]);
                if (CARRY) {
#pragma unroll
                    for (int j = 0; j < NG; ++j) cg[j] = pk.v[f++];
                }
                active = pk.active;
            }
            run_bursts(BURST * TAIL_BURSTS);
        }
      }
    }
    // (ensemble totals are summed from the per-trajectory counters afterwards: sum_counts_kernel)
}

// ---------------------------------------------------------------------------------------------------------------
// Row-synchronous dense output: the tspan rows of oderk_adapt (problem.rs:303-319) in the SoA layout [N][nt][ntraj] with
// fully coalesced stores for ensembles in ANY order.
//
// In the persistent-lane kernel every lane emits its rows inside its own accepted steps; a warp's stores land in one row
// only while its lanes advance together (a parameter-sorted sweep).  On a permuted sweep every lane sits in a different
// row: 32 single-sector stores per instruction, a partial-sector read-modify-write in L2 for each of them (ncu: 65 GB
// written + 61 GB read for 17 GB of rows, 84 ms instead of 12 ms at 2^20 trajectories), and the emission loop diverges
// (20 of 32 lanes active).  Here a warp owns 32 CONSECUTIVE trajectories and walks the output rows together: row r is
// emitted by the whole warp at once — one 256-byte store per component — as soon as every lane's last accepted step
// reaches past tspan[r]; lanes that are not there yet take step attempts while the others wait, holding the Hermite data
// of their last accepted step (its start state and slope; the end state and slope are the current ones).  The warp thus
// advances in TIME, not in step count; the number of attempts it executes is that of its slowest lane, like the static
// assignment of the other kernel.  Same arithmetic per trajectory (attempt_math, the reference's Hermite formula and
// row conditions, `nvalid` quirk included): results are bit-identical to MODE_DENSE.
// ---------------------------------------------------------------------------------------------------------------
template <class TB, class F, bool PT, bool FAST>
__global__ void __launch_bounds__(BLOCK, DEQ_MIN_BLOCKS) adapt_rows_kernel(const AdaptParams P) {
    constexpr int N = F::N, S = TB::S;
    constexpr bool FSAL = deqt::fsal_of(TB::data());
    __shared__ uint64_t s_exp[256];
    __shared__ double s_powlog[384];
    for (int i = threadIdx.x; i < 256; i += BLOCK) s_exp[i] = g_exp_tab[i];
    for (int i = threadIdx.x; i < 384; i += BLOCK) s_powlog[i] = g_powlog_tab[i];
    if (F::USES_SINCOS) deqm::sincos_stage();
    __syncthreads();
    unsigned exp_base = (unsigned)__cvta_generic_to_shared(s_exp), powlog_base = (unsigned)__cvta_generic_to_shared(s_powlog);
    asm volatile("mov.u32 %0, %0;" : "+r"(exp_base));
    asm volatile("mov.u32 %0, %0;" : "+r"(powlog_base));
    const deqm::SmemTables T{exp_base, powlog_base, 0u};

    const unsigned lane = threadIdx.x & 31u;
    const double tdir = P.tdir, tend = P.tend, reltol = P.reltol, abstol = P.abstol;
    const uint32_t max_steps = P.max_steps == 0ull || P.max_steps > 0xffffffffull ? 0xffffffffu : (uint32_t)P.max_steps;
    const unsigned long long nt = P.nt;

    for (;;) {
        // ---- the next block of 32 consecutive trajectories ----
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.work_counter, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= P.ntraj) break;
        const unsigned long long traj = base + lane;
        const bool live = traj < P.ntraj;
        double cy[N], ck[N], yl[N], prm[F::NP ? F::NP : 1];
        double y0v[N], f0v[N];           // start state and slope of the last accepted step (its end: cy, ck)
        constexpr int NG = ng_of<F>::value;
        constexpr bool CARRY = carries_forcing<TB, F>();
        constexpr bool GEND = NG > 0 && has_c_row(TB::data(), 1.0);
        double cg[NG ? NG : 1];
#pragma unroll
        for (int j = 0; j < (NG ? NG : 1); ++j) cg[j] = 0.0;
        double t = P.t0, dt = 0.0, t0v = P.t0, dt0v = 0.0;
        unsigned timeout = 0u, last_step = 0u, done = live ? 0u : 1u, have_iv = 0u;
        uint32_t acc = 0, rej = 0, attempts = 0;
        unsigned long long it = 1;
        uint8_t status = ST_DONE;
#pragma unroll (N <= 16 ? N : 1)
        for (int d = 0; d < N; ++d) { cy[d] = 0.0; ck[d] = 0.0; yl[d] = 0.0; y0v[d] = 0.0; f0v[d] = 0.0; }
#pragma unroll
        for (int j = 0; j < (F::NP ? F::NP : 1); ++j) prm[j] = 0.0;
        if (live) {
            load_params<F, PT>(prm, P.params, P.shared, traj, P.ntraj);
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) {
                cy[d] = P.y0[(unsigned long long)d * P.ntraj + traj];
                ck[d] = P.f0[(unsigned long long)d * P.ntraj + traj];
                P.dense[((unsigned long long)d * nt) * P.ntraj + traj] = cy[d];   // row 0 = y0
            }
            if constexpr (CARRY) F::forcing(t, cg, prm);
            dt = P.use_initstep ? P.initstep : P.h0[traj];
            last_step = fabs((t + dt) - tend) <= DBL_EPSILON ? 1u : 0u;
        }
        unsigned long long r = 1;   // warp-uniform row cursor; rows 1 .. nt-2 are Hermite rows, row nt-1 is the final state (counted only)
        for (;;) {
            // ---- emit every row the whole warp is ready for ----
            // (row nt-1 takes part as well: its VALUE is the final state, written below, but the reference's emission loop
            // (problem.rs:303-319) also counts tspan[nt-1] when the last step ends a rounding error past it — `t + (tend - t)` need
            // not be `tend` — or when a user initstep overshoots the end: `nvalid` then reads nt)
            while (r < nt) {
                const double tq = __ldg(P.tspan + r);
                bool emit = false, wait = false;
                if (live && it == r) {   // (it < r: this lane has stopped emitting — a step end hit a tspan point exactly — or its run ended early)
                    const bool after_start = have_iv && tdir * t0v < tdir * tq;
                    if (after_start && tdir * tq < tdir * t) emit = true;          // strictly inside the last accepted step
                    else if (!done && (!have_iv || after_start)) wait = true;     // not reached yet: step first
                }
                if (__any_sync(0xffffffffu, wait)) break;
                if (emit) {
                    if (r + 1 < nt) {
                        const double theta = (tq - t0v) / dt0v;
#pragma unroll (N <= 16 ? N : 1)
                        for (int d = 0; d < N; ++d) {
                            const double v = (y0v[d] * (1. - theta) + cy[d] * theta) +
                                             ((cy[d] - y0v[d]) * (1. - 2. * theta) + f0v[d] * (theta - 1.) * dt0v +
                                              ck[d] * theta * dt0v) * theta * (theta - 1.);
                            P.dense[((unsigned long long)d * nt + r) * P.ntraj + traj] = v;
                        }
                    }
                    ++it;
                }
                ++r;
            }
            // ---- one attempt for the lanes that have not reached row r (or emit nothing any more) and are not finished ----
            bool need = live && !done;
            if (need && it == r && r < nt && have_iv) {
                const double tq = __ldg(P.tspan + r);
                if (tdir * t0v < tdir * tq && tdir * tq < tdir * t) need = false;   // holds a row the warp has not emitted yet
            }
            if (!__any_sync(0xffffffffu, need)) break;
            if (need) {
                ++attempts;
                double k[S][N], ytrial[N];
                double ndt;
                unsigned accu;
                const unsigned timeout_next = timeout - (timeout > 0u ? 1u : 0u);
                double gend[NG ? NG : 1];
                attempt_math<TB, F, FAST, false>(P, T, t, dt, cy, ck, yl, prm, timeout, tdir, reltol, abstol, k, ytrial, ndt, accu,
                                                 cg, GEND ? gend : nullptr);
                const bool accept = accu != 0u;
                const double tn = t + dt;
                if (accept) {
                    if (!FSAL) {
                        if constexpr (GEND) F::eval_g(tn, ytrial, k[S - 1], prm, gend);
                        else F::eval(tn, ytrial, k[S - 1], prm);
                    }
                    if constexpr (CARRY) {
#pragma unroll
                        for (int j = 0; j < NG; ++j) cg[j] = gend[j];
                    }
#pragma unroll (N <= 16 ? N : 1)
                    for (int d = 0; d < N; ++d) { y0v[d] = cy[d]; f0v[d] = ck[d]; cy[d] = ytrial[d]; ck[d] = k[S - 1][d]; }
                    t0v = t; dt0v = dt; have_iv = 1u;
                    t = tn;
                }
                const double dt100 = FAST ? ndt * DEQ_C001 : div100_exact(ndt);
                const bool snapped = accept && (tdir * ((t + ndt) + dt100) >= P.tdir_tend);
                const bool stop = accept ? last_step != 0u : (fabs(ndt) < P.minstep);
                dt = snapped ? tend - t : ndt;
                last_step = snapped ? 1u : 0u;
                timeout = accept ? timeout_next : 5u;
                acc += accept ? 1u : 0u;
                rej += (accept || stop) ? 0u : 1u;
                if (stop || attempts >= max_steps) {
                    status = !stop ? (uint8_t)ST_MAX_STEPS : accept ? (uint8_t)ST_DONE : (uint8_t)ST_MINSTEP_BREAK;
                    done = 1u;
                }
            }
        }
        // ---- the block is finished: final states, counters, the end point as the last row (problem.rs:320-322) ----
        if (live) {
            const uint32_t nfe = 2u + attempts * (uint32_t)(S - 1) + (FSAL ? 0u : acc);
#pragma unroll (N <= 16 ? N : 1)
            for (int d = 0; d < N; ++d) {
                P.yfinal[(unsigned long long)d * P.ntraj + traj] = cy[d];
                P.dense[((unsigned long long)d * nt + (nt - 1)) * P.ntraj + traj] = cy[d];
            }
            P.accepted[traj] = acc; P.rejected[traj] = rej; P.nfe[traj] = nfe; P.status[traj] = status; P.t_reached[traj] = t;
            P.nvalid[traj] = (uint32_t)it;
        }
    }
}

}  // namespace deqk
