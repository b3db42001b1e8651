// stiff.cuh — the stiff solvers of the `Ode` enum, one trajectory per thread:
//   ode23s_kernel      modified Rosenbrock triple, adaptive      /root/reference/src/ode/problem.rs:409-557
//   rosenbrock_kernel  4-stage Rosenbrock on the tspan grid      /root/reference/src/ode/problem.rs:560-640
//                      (coefficient sets rosenbrock.rs:16-68 kr4, :71-117 s4)
//   fdjacobian         forward-difference Jacobian               /root/reference/src/ode/problem.rs:903-926
//
// Same parity contract as stepper.cuh: IEEE f64, -fmad=false, the reference's operation order.  The n x n linear
// algebra of the reference goes through nalgebra 0.20 (DMatrix operators, lu(), try_inverse()); it is restated here for
// n <= 32, every matrix per thread: in registers up to n = 8 (loops fully unrolled, all indices compile-time), in thread-local
// memory beyond (`#pragma unroll (N <= 8 ? N : 1)`: the loops stay rolled, like the explicit path does beyond n = 16):
//   * M * v             column-ordered accumulation  y = M[:,0] v0;  y = M[:,j] v_j + y
//   * lu()              partial pivoting on the first largest |.|, multipliers a_ji * (1 / pivot), exactly-zero pivot
//                       columns skipped; only the packed factors survive (`lu_internal()`), the permutation is dropped
//   * try_inverse()     adjugate formulas for n = 1, 2, 3, the 4x4 cofactor routine (do_inverse4) for n = 4, lu::try_invert_to
//                       for n >= 5; a zero determinant / pivot is the reference's OdeError::InvalidMatrix (per-trajectory
//                       status ST_INVALID_MATRIX here)
// The reference inverts the PACKED LU matrix of W = I - h d J rather than W (problem.rs:469,481); that is what the
// REFERENCE set reproduces.  P.true_inverse (TEXTBOOK set) inverts W itself — Shampine & Reichelt's method.
#pragma once
#include "stepper.cuh"

namespace deqk {

template <int N>
__device__ __forceinline__ void na_gemv(const double (&m)[N][N], const double (&v)[N], double (&y)[N]) {
#pragma unroll (N <= 8 ? N : 1)
    for (int r = 0; r < N; ++r) y[r] = m[r][0] * v[0];
#pragma unroll (N <= 8 ? N : 1)
    for (int j = 1; j < N; ++j) {
#pragma unroll (N <= 8 ? N : 1)
        for (int r = 0; r < N; ++r) y[r] = m[r][j] * v[j] + y[r];
    }
}

template <int N>
__device__ __forceinline__ void na_lu_packed(double (&m)[N][N]) {
#pragma unroll (N <= 8 ? N : 1)
    for (int i = 0; i < N; ++i) {
        int piv = i;
        double best = fabs(m[i][i]);
#pragma unroll (N <= 8 ? N : 1)
        for (int r = i + 1; r < N; ++r) { const double v = fabs(m[r][i]); if (v > best) { best = v; piv = r; } }
        // swap rows i and piv (all columns: the ones left of i by swap_rows, the rest inside gauss_step_swap)
#pragma unroll (N <= 8 ? N : 1)
        for (int r = i + 1; r < N; ++r) {
            if (piv == r) {
#pragma unroll (N <= 8 ? N : 1)
                for (int c = 0; c < N; ++c) { const double tmp = m[i][c]; m[i][c] = m[r][c]; m[r][c] = tmp; }
            }
        }
        const double diag = m[i][i];
        if (diag != 0.0) {
            const double inv = 1.0 / diag;
#pragma unroll (N <= 8 ? N : 1)
            for (int r = i + 1; r < N; ++r) m[r][i] = m[r][i] * inv;
#pragma unroll (N <= 8 ? N : 1)
            for (int k = i + 1; k < N; ++k) {
                const double a = -m[i][k];
#pragma unroll (N <= 8 ? N : 1)
                for (int r = i + 1; r < N; ++r) m[r][k] = a * m[r][i] + m[r][k];
            }
        }
    }
}

template <int N>
__device__ __forceinline__ bool na_try_inverse(const double (&m)[N][N], double (&o)[N][N]);

// FAST arithmetic only: the same adjugate formulas with ONE reciprocal of the determinant instead of n^2 IEEE divisions.
template <int N>
__device__ __forceinline__ bool inverse_fast(const double (&m)[N][N], double (&o)[N][N]) {
    static_assert(N >= 1 && N <= 32, "the stiff path carries n <= 32");
    if constexpr (N >= 4) {
        return na_try_inverse<N>(m, o);   // same routines, contracted by the FAST translation unit's -fmad=true
    } else
    if constexpr (N == 1) {
        if (m[0][0] == 0.0) return false;
        o[0][0] = 1.0 / m[0][0];
    } else if constexpr (N == 2) {
        const double det = m[0][0] * m[1][1] - m[1][0] * m[0][1];
        if (det == 0.0) return false;
        const double r = 1.0 / det;
        o[0][0] = m[1][1] * r; o[0][1] = -m[0][1] * r; o[1][0] = -m[1][0] * r; o[1][1] = m[0][0] * r;
    } else {
        const double c00 = m[1][1] * m[2][2] - m[2][1] * m[1][2], c01 = m[1][0] * m[2][2] - m[2][0] * m[1][2],
                     c02 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
        const double det = m[0][0] * c00 - m[0][1] * c01 + m[0][2] * c02;
        if (det == 0.0) return false;
        const double r = 1.0 / det;
        o[0][0] = c00 * r; o[0][1] = (m[0][2] * m[2][1] - m[2][2] * m[0][1]) * r; o[0][2] = (m[0][1] * m[1][2] - m[1][1] * m[0][2]) * r;
        o[1][0] = -c01 * r; o[1][1] = (m[0][0] * m[2][2] - m[2][0] * m[0][2]) * r; o[1][2] = (m[0][2] * m[1][0] - m[1][2] * m[0][0]) * r;
        o[2][0] = c02 * r; o[2][1] = (m[0][1] * m[2][0] - m[2][1] * m[0][0]) * r; o[2][2] = (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * r;
    }
    return true;
}

template <int N>
__device__ __forceinline__ bool na_try_inverse(const double (&m)[N][N], double (&o)[N][N]) {
    static_assert(N >= 1 && N <= 32, "the stiff path carries n <= 32");
    if constexpr (N == 4) {
        // `do_inverse4` (nalgebra 0.20 linalg/inverse.rs; MESA's gluInvertMatrix cofactor expansion) on the column-major slice
        // f[k] = M[k % 4][k / 4]: products and sums left to right, det from the first row of cofactors, then every cofactor is
        // MULTIPLIED by 1/det (the n <= 3 formulas divide)
        double f[16], c[16];
        _Pragma("unroll") for (int k = 0; k < 16; ++k) f[k] = m[k % 4][k / 4];
        c[0] = f[5] * f[10] * f[15] - f[5] * f[11] * f[14] - f[9] * f[6] * f[15] + f[9] * f[7] * f[14] + f[13] * f[6] * f[11] - f[13] * f[7] * f[10];
        c[1] = -f[1] * f[10] * f[15] + f[1] * f[11] * f[14] + f[9] * f[2] * f[15] - f[9] * f[3] * f[14] - f[13] * f[2] * f[11] + f[13] * f[3] * f[10];
        c[2] = f[1] * f[6] * f[15] - f[1] * f[7] * f[14] - f[5] * f[2] * f[15] + f[5] * f[3] * f[14] + f[13] * f[2] * f[7] - f[13] * f[3] * f[6];
        c[3] = -f[1] * f[6] * f[11] + f[1] * f[7] * f[10] + f[5] * f[2] * f[11] - f[5] * f[3] * f[10] - f[9] * f[2] * f[7] + f[9] * f[3] * f[6];
        c[4] = -f[4] * f[10] * f[15] + f[4] * f[11] * f[14] + f[8] * f[6] * f[15] - f[8] * f[7] * f[14] - f[12] * f[6] * f[11] + f[12] * f[7] * f[10];
        c[5] = f[0] * f[10] * f[15] - f[0] * f[11] * f[14] - f[8] * f[2] * f[15] + f[8] * f[3] * f[14] + f[12] * f[2] * f[11] - f[12] * f[3] * f[10];
        c[6] = -f[0] * f[6] * f[15] + f[0] * f[7] * f[14] + f[4] * f[2] * f[15] - f[4] * f[3] * f[14] - f[12] * f[2] * f[7] + f[12] * f[3] * f[6];
        c[7] = f[0] * f[6] * f[11] - f[0] * f[7] * f[10] - f[4] * f[2] * f[11] + f[4] * f[3] * f[10] + f[8] * f[2] * f[7] - f[8] * f[3] * f[6];
        c[8] = f[4] * f[9] * f[15] - f[4] * f[11] * f[13] - f[8] * f[5] * f[15] + f[8] * f[7] * f[13] + f[12] * f[5] * f[11] - f[12] * f[7] * f[9];
        c[9] = -f[0] * f[9] * f[15] + f[0] * f[11] * f[13] + f[8] * f[1] * f[15] - f[8] * f[3] * f[13] - f[12] * f[1] * f[11] + f[12] * f[3] * f[9];
        c[10] = f[0] * f[5] * f[15] - f[0] * f[7] * f[13] - f[4] * f[1] * f[15] + f[4] * f[3] * f[13] + f[12] * f[1] * f[7] - f[12] * f[3] * f[5];
        c[11] = -f[0] * f[5] * f[11] + f[0] * f[7] * f[9] + f[4] * f[1] * f[11] - f[4] * f[3] * f[9] - f[8] * f[1] * f[7] + f[8] * f[3] * f[5];
        c[12] = -f[4] * f[9] * f[14] + f[4] * f[10] * f[13] + f[8] * f[5] * f[14] - f[8] * f[6] * f[13] - f[12] * f[5] * f[10] + f[12] * f[6] * f[9];
        c[13] = f[0] * f[9] * f[14] - f[0] * f[10] * f[13] - f[8] * f[1] * f[14] + f[8] * f[2] * f[13] + f[12] * f[1] * f[10] - f[12] * f[2] * f[9];
        c[14] = -f[0] * f[5] * f[14] + f[0] * f[6] * f[13] + f[4] * f[1] * f[14] - f[4] * f[2] * f[13] - f[12] * f[1] * f[6] + f[12] * f[2] * f[5];
        c[15] = f[0] * f[5] * f[10] - f[0] * f[6] * f[9] - f[4] * f[1] * f[10] + f[4] * f[2] * f[9] + f[8] * f[1] * f[6] - f[8] * f[2] * f[5];
        const double det = f[0] * c[0] + f[1] * c[4] + f[2] * c[8] + f[3] * c[12];
        if (det == 0.0) return false;
        const double inv_det = 1.0 / det;
        _Pragma("unroll") for (int k = 0; k < 16; ++k) o[k % 4][k / 4] = c[k] * inv_det;
    } else if constexpr (N >= 5) {
        // `lu::try_invert_to` (nalgebra 0.20 linalg/lu.rs), what try_inverse does for n >= 5: Gaussian elimination with partial
        // pivoting on a copy (first largest |.|, multipliers a_ji * (1/pivot), update a_jk = (-p_k) l_j + a_jk; a zero pivot fails),
        // the row swaps applied to an identity `o`; then o <- L^-1 o (unit diagonal, forward, column by column) and o <- U^-1 o
        // (backward: coeff = b_i / u_ii, b_j = (-coeff) u_ji + b_j for j < i)
        double a[N][N];
        _Pragma("unroll (N <= 8 ? N : 1)") for (int r = 0; r < N; ++r) {
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int cc = 0; cc < N; ++cc) { a[r][cc] = m[r][cc]; o[r][cc] = r == cc ? 1.0 : 0.0; }
        }
        bool ok = true;
        _Pragma("unroll (N <= 8 ? N : 1)") for (int i = 0; i < N; ++i) {
            int piv = i;
            double best = fabs(a[i][i]);
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int r = i + 1; r < N; ++r) { const double v = fabs(a[r][i]); if (v > best) { best = v; piv = r; } }
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int r = i + 1; r < N; ++r) {
                if (piv == r) {
        _Pragma("unroll (N <= 8 ? N : 1)")             for (int cc = 0; cc < N; ++cc) {
                        double tmp = a[i][cc]; a[i][cc] = a[r][cc]; a[r][cc] = tmp;
                        tmp = o[i][cc]; o[i][cc] = o[r][cc]; o[r][cc] = tmp;
                    }
                }
            }
            const double diag = a[i][i];
            if (diag == 0.0) ok = false;
            const double inv = 1.0 / diag;
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int r = i + 1; r < N; ++r) a[r][i] = a[r][i] * inv;
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int k = i + 1; k < N; ++k) {
                const double p = -a[i][k];
        _Pragma("unroll (N <= 8 ? N : 1)")         for (int r = i + 1; r < N; ++r) a[r][k] = p * a[r][i] + a[r][k];
            }
        }
        if (!ok) return false;   // (the reference returns at the first zero pivot; nothing computed after it is used)
        _Pragma("unroll (N <= 8 ? N : 1)") for (int k = 0; k < N; ++k) {
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int i = 0; i + 1 < N; ++i) {
                const double coeff = o[i][k] / 1.0;
        _Pragma("unroll (N <= 8 ? N : 1)")         for (int r = i + 1; r < N; ++r) o[r][k] = (-coeff) * a[r][i] + o[r][k];
            }
        }
        _Pragma("unroll (N <= 8 ? N : 1)") for (int k = 0; k < N; ++k) {
        _Pragma("unroll (N <= 8 ? N : 1)")     for (int i = N - 1; i >= 0; --i) {
                const double diag = a[i][i];
                if (diag == 0.0) ok = false;
                const double coeff = o[i][k] / diag;
                o[i][k] = coeff;
        _Pragma("unroll (N <= 8 ? N : 1)")         for (int r = 0; r < i; ++r) o[r][k] = (-coeff) * a[r][i] + o[r][k];
            }
        }
        if (!ok) return false;
    } else
    if constexpr (N == 1) {
        const double det = m[0][0];
        if (det == 0.0) return false;
        o[0][0] = 1.0 / det;
    } else if constexpr (N == 2) {
        const double m11 = m[0][0], m12 = m[0][1], m21 = m[1][0], m22 = m[1][1];
        const double det = m11 * m22 - m21 * m12;
        if (det == 0.0) return false;
        o[0][0] = m22 / det; o[0][1] = -m12 / det; o[1][0] = -m21 / det; o[1][1] = m11 / det;
    } else {
        const double m11 = m[0][0], m12 = m[0][1], m13 = m[0][2], m21 = m[1][0], m22 = m[1][1], m23 = m[1][2],
                     m31 = m[2][0], m32 = m[2][1], m33 = m[2][2];
        const double minor_m12_m23 = m22 * m33 - m32 * m23;
        const double minor_m11_m23 = m21 * m33 - m31 * m23;
        const double minor_m11_m22 = m21 * m32 - m31 * m22;
        const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        if (det == 0.0) return false;
        o[0][0] = minor_m12_m23 / det; o[0][1] = (m13 * m32 - m33 * m12) / det; o[0][2] = (m12 * m23 - m22 * m13) / det;
        o[1][0] = -minor_m11_m23 / det; o[1][1] = (m11 * m33 - m31 * m13) / det; o[1][2] = (m13 * m21 - m23 * m11) / det;
        o[2][0] = minor_m11_m22 / det; o[2][1] = (m12 * m31 - m32 * m11) / det; o[2][2] = (m11 * m22 - m21 * m12) / det;
    }
    return true;
}

// fdjacobian (problem.rs:903-926).  `ftx` = f(t, x): the reference evaluates it again at the top of fdjacobian; every
// caller here already holds that very value (same t, same x, a deterministic f), so it is passed in.
template <class F, bool FAST = false>
__device__ __forceinline__ void fdjacobian(double t, const double (&x)[F::N], const double (&ftx)[F::N], const double* prm,
                                           double (&J)[F::N][F::N]) {
    constexpr int N = F::N;
#pragma unroll (N <= 8 ? N : 1)
    for (int c = 0; c < N; ++c) {
        double xj = x[c];
        if (xj == 0.0) xj += 1.0;
        const double dxj = xj * 0.01;
        double tmp[N], yj[N];
#pragma unroll (N <= 8 ? N : 1)
        for (int d = 0; d < N; ++d) tmp[d] = x[d];
        tmp[c] += dxj;
        F::eval(t, tmp, yj, prm);
        if (FAST) {
            const double rdx = 1.0 / dxj;
#pragma unroll (N <= 8 ? N : 1)
            for (int r = 0; r < N; ++r) J[r][c] = (yj[r] - ftx[r]) * rdx;
        } else {
#pragma unroll (N <= 8 ? N : 1)
            for (int r = 0; r < N; ++r) J[r][c] = (yj[r] - ftx[r]) / dxj;
        }
    }
}

// FAST arithmetic only: 2-norm by sqrt (what an LLVM-optimised build of the crate evaluates anyway)
template <int N>
__device__ __forceinline__ double norm2_fast(const double (&v)[N]) {
    double s = 0.0;
#pragma unroll (N <= 8 ? N : 1)
    for (int d = 0; d < N; ++d) s = __fma_rn(v[d], v[d], s);
    return sqrt(s);
}

template <int N, class TT>
__device__ __forceinline__ double pnorm2(const double (&v)[N], int libm_folded, const TT& T) {   // types.rs:109-115, P(2)
    double s = 0.0;
#pragma unroll (N <= 8 ? N : 1)
    for (int d = 0; d < N; ++d) { const double a = fabs(v[d]); s = s + a * a; }
    return libm_folded ? sqrt(s) : deqm::pow_glibc(s, 0.5, T);
}

// ---------------------------------------------------------------------------------------------------------------
// ode23s (problem.rs:409-557).  Persistent lanes re-filled from a global counter like adapt_kernel.  MODE_FINAL: loop
// state at exit; MODE_DENSE: the rows the reference pushes for the tspan times, SoA [N][nt][ntraj]; MODE_REC: the full
// (tout, yout) stream, rows [t, y...], counting pass then writing pass.  The reference rescans the whole tspan on every
// accepted step (:519); for a monotone tspan a cursor yields the same rows while time moves with the direction of
// integration, and the kernel falls back to the full scan when it does not (capi.cu rejects a non-monotone tspan).
// ---------------------------------------------------------------------------------------------------------------
// FAST (compiled in the -fmad=true objects only, deq_options.mode = DEQ_MODE_FAST): fused multiply-adds, one reciprocal per
// inverse / Jacobian column instead of n^2 divisions, sqrt norms, the 17-instruction table pow for the step factor.  Same
// formulas, results within the solver tolerance of PARITY (tests/test_gpu_stiff.py); no bit-parity claim.
template <class F, bool PT, int MODE, bool FAST = false>
__global__ void __launch_bounds__(BLOCK) ode23s_kernel(const AdaptParams P) {
    constexpr int N = F::N;
    constexpr bool SOA = MODE == MODE_DENSE, REC = MODE == MODE_REC, ROWS = MODE != MODE_FINAL;
    static_assert(MODE != MODE_DENSE_TM, "the stiff path writes SoA rows");
    constexpr double SQ2 = 1.4142135623730951;      // 2f64.sqrt()
    constexpr double D = 1. / (2. + SQ2);
    constexpr double E32 = 6. + SQ2;

    __shared__ uint64_t s_exp[256];
    __shared__ double s_powlog[384];
    for (int i = threadIdx.x; i < 256; i += BLOCK) s_exp[i] = g_exp_tab[i];
    for (int i = threadIdx.x; i < 384; i += BLOCK) s_powlog[i] = g_powlog_tab[i];
    if (F::USES_SINCOS) deqm::sincos_stage();
    __syncthreads();
    const deqm::SmemTables T{(unsigned)__cvta_generic_to_shared(s_exp), (unsigned)__cvta_generic_to_shared(s_powlog), 0u};

    const unsigned lane = threadIdx.x & 31u;
    const double tdir = P.tdir, tfinal = P.tend, reltol = P.reltol, abstol = P.abstol;
    const uint32_t max_steps = P.max_steps == 0ull || P.max_steps > 0xffffffffull ? 0xffffffffu : (uint32_t)P.max_steps;

    bool active = false, pool_empty = false, disorder = false;
    unsigned long long traj = 0, it = 1, wpos = 0;
    double t = 0.0, h = 0.0, tlast = 0.0;
    double y[N], f0[N], jac[N][N], prm[F::NP ? F::NP : 1];
    uint32_t acc = 0, rej = 0, attempts = 0, nv = 1;
    unsigned long long tot_acc = 0, tot_rej = 0, tot_nfe = 0;
#pragma unroll (N <= 8 ? N : 1)
    for (int d = 0; d < N; ++d) {
        y[d] = 0.0; f0[d] = 0.0;
#pragma unroll (N <= 8 ? N : 1)
        for (int c = 0; c < N; ++c) jac[d][c] = 0.0;
    }
#pragma unroll (N <= 8 ? N : 1)
    for (int j = 0; j < (F::NP ? F::NP : 1); ++j) prm[j] = 0.0;

    auto emit = [&](double tq, const double (&v)[N]) {   // one REC row
        if (P.rec_mode == REC_WRITE) {
            double* row = P.rec_rows + (P.rec_offsets[traj] + wpos) * (unsigned long long)(N + 1);
            row[0] = tq;
#pragma unroll (N <= 8 ? N : 1)
            for (int d = 0; d < N; ++d) row[1 + d] = v[d];
        }
        ++wpos;
    };

    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !pool_empty && (P.refill || idle == 0xffffffffu)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.work_counter, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned long long)__popc(idle) >= P.ntraj) pool_empty = true;
            if (!active) {
                const unsigned long long my = base + (unsigned long long)__popc(idle & ((1u << lane) - 1u));
                if (my < P.ntraj) {
                    traj = my;
                    active = true;
                    load_params<F, PT>(prm, P.params, P.shared, traj, P.ntraj);
#pragma unroll (N <= 8 ? N : 1)
                    for (int d = 0; d < N; ++d) {
                        y[d] = P.y0[(unsigned long long)d * P.ntraj + traj];
                        f0[d] = P.f0[(unsigned long long)d * P.ntraj + traj];   // init.f0 (:435-441)
                    }
                    t = P.t0;
                    const double hh = P.use_initstep ? P.initstep : P.h0[traj];
                    h = tdir * fmin(fabs(hh), P.maxstep);                      // :443
                    acc = 0; rej = 0; attempts = 0; it = 1; wpos = 0; nv = 1; tlast = t; disorder = false;
                    if (SOA) {
#pragma unroll (N <= 8 ? N : 1)
                        for (int d = 0; d < N; ++d) P.dense[((unsigned long long)d * P.nt) * P.ntraj + traj] = y[d];
                    }
                    if (REC) emit(t, y);
                    fdjacobian<F, FAST>(t, y, f0, prm, jac);                   // :454
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) break;

#pragma unroll 1
        for (int burst = 0; burst < BURST_ATTEMPTS && active; ++burst) {
            uint8_t fin = 0xff;
            bool invalid = false;
            const double rem = fabs(t - tfinal);
            if (!(rem > 0.)) {
                fin = ST_DONE;
            } else if (!(P.minstep < fabs(h))) {
                fin = ST_MINSTEP_BREAK;           // the while condition :462 ends the loop before tfinal, silently
            } else if (attempts >= max_steps) {
                fin = ST_MAX_STEPS;
            } else {
                ++attempts;
                if (rem < fabs(h)) h = tfinal - t;
                const double hd = h * D;
                double w[N][N];
#pragma unroll (N <= 8 ? N : 1)
                for (int r = 0; r < N; ++r) {
#pragma unroll (N <= 8 ? N : 1)
                    for (int c = 0; c < N; ++c) w[r][c] = (r == c ? 1.0 : 0.0) - jac[r][c] * hd;
                }
                if (N != 1 && !P.true_inverse) na_lu_packed<N>(w);
                double fdt[N];
                F::eval(t + h / 100., y, fdt, prm);
                const double sc = hd / (h / 100.);
#pragma unroll (N <= 8 ? N : 1)
                for (int i = 0; i < N; ++i) fdt[i] = (fdt[i] - f0[i]) * sc;
                double winv[N][N];
                if (!(FAST ? inverse_fast<N>(w, winv) : na_try_inverse<N>(w, winv))) {
                    fin = ST_INVALID_MATRIX; invalid = true;
                } else {
                    double rhs[N], k1[N], k2[N], k3[N], f1[N], f2[N], ynew[N];
#pragma unroll (N <= 8 ? N : 1)
                    for (int i = 0; i < N; ++i) rhs[i] = f0[i] + fdt[i];
                    na_gemv<N>(winv, rhs, k1);
#pragma unroll (N <= 8 ? N : 1)
                    for (int i = 0; i < N; ++i) rhs[i] = y[i] + (k1[i] * 0.5) * h;
                    F::eval(t + 0.5 * h, rhs, f1, prm);
#pragma unroll (N <= 8 ? N : 1)
                    for (int i = 0; i < N; ++i) rhs[i] = f1[i] - k1[i];
                    na_gemv<N>(winv, rhs, k2);
#pragma unroll (N <= 8 ? N : 1)
                    for (int i = 0; i < N; ++i) { k2[i] = k2[i] + k1[i]; ynew[i] = y[i] + k2[i] * h; }
                    const double tn = t + h;
                    F::eval(tn, ynew, f2, prm);
#pragma unroll (N <= 8 ? N : 1)
                    for (int i = 0; i < N; ++i) rhs[i] = ((f2[i] - (k2[i] - f1[i]) * E32) - (k1[i] - f0[i]) * 2.) + fdt[i];
                    na_gemv<N>(winv, rhs, k3);
                    double kerr[N];
#pragma unroll (N <= 8 ? N : 1)
                    for (int i = 0; i < N; ++i) kerr[i] = (k1[i] - k2[i] * 2.) + k3[i];
                    const double err = (FAST ? norm2_fast<N>(kerr) : pnorm2<N>(kerr, P.libm_folded, T)) * (fabs(h) / 6.);
                    const double delta = FAST ? fmax(fmax(norm2_fast<N>(y), norm2_fast<N>(ynew)) * reltol, abstol)
                                              : fmax(fmax(pnorm2<N>(y, P.libm_folded, T), pnorm2<N>(ynew, P.libm_folded, T)) * reltol, abstol);
                    if (err <= delta) {
                        acc += 1;
                        if (ROWS) {
                            // :519-534 — the reference rescans the whole tspan after every accepted step for `toi > t && toi <=
                            // t + h`, which needs h > 0.  A forward run that has always moved forward visits the grid in order: a
                            // cursor.  Otherwise — the tiny correction step with h > 0 at the end of a BACKWARD run that
                            // overshot tfinal, or any step after time once moved against the direction — the full scan.
                            if (h > 0.) {
                                const bool scan_all = tdir < 0. || disorder;
                                unsigned long long q = scan_all ? 0ull : it;
                                while (q < P.nt) {
                                    const double tq = __ldg(P.tspan + q);
                                    if (!(tq > t)) { ++q; continue; }            // in cursor mode: can never satisfy `toi > t` again
                                    if (!(tq <= tn)) { if (scan_all) { ++q; continue; } break; }
                                    const double s = (tq - t) / h;
                                    const double c1 = (s * (1. - s)) / (1. - 2. * D);
                                    const double c2 = (s * (s - 2. * D)) / (1. - 2. * D);
                                    double ytmp[N];
#pragma unroll (N <= 8 ? N : 1)
                                    for (int i = 0; i < N; ++i) ytmp[i] = y[i] + (k1[i] * c1 + k2[i] * c2) * h;
                                    if (SOA) {
#pragma unroll (N <= 8 ? N : 1)
                                        for (int d = 0; d < N; ++d) P.dense[((unsigned long long)d * P.nt + q) * P.ntraj + traj] = ytmp[d];
                                        nv = (uint32_t)q + 1u;
                                    }
                                    if (REC) emit(tq, ytmp);
                                    tlast = tq;
                                    ++q;
                                }
                                if (!scan_all) it = q;
                            } else if (tdir > 0.) {
                                disorder = true;   // a forward run stepped backwards (overshoot correction): grid order is lost
                            }
                            if (REC && !P.points_specified && fabs(tlast - tn) > DBL_EPSILON) { emit(tn, ynew); tlast = tn; }   // :536-542
                        }
                        t = tn;
#pragma unroll (N <= 8 ? N : 1)
                        for (int i = 0; i < N; ++i) { y[i] = ynew[i]; f0[i] = f2[i]; }
                        fdjacobian<F, FAST>(t, y, f0, prm, jac);                 // :549
                    } else {
                        rej += 1;
                    }
                    const double r = delta / err;
                    const double g = FAST ? deqm::pow_fast_pos(r, 1. / 3., T) : deqm::pow_glibc(r, 1. / 3., T);
                    h = fmin(P.maxstep, (g * fabs(h)) * 0.8) * tdir;   // :552-553
                }
            }
            if (fin != 0xff) {
                // evaluations the reference performs: hinit 2 (1 with an explicit initstep), N+1 per Jacobian, 3 per attempt
                const uint32_t nfe = (P.use_initstep ? 1u : 2u) + (uint32_t)(N + 1) * (1u + acc) + 3u * (attempts - (invalid ? 1u : 0u)) + (invalid ? 1u : 0u);
#pragma unroll (N <= 8 ? N : 1)
                for (int d = 0; d < N; ++d) P.yfinal[(unsigned long long)d * P.ntraj + traj] = y[d];
                P.accepted[traj] = acc; P.rejected[traj] = rej; P.nfe[traj] = nfe; P.status[traj] = fin; P.t_reached[traj] = t;
                if (ROWS) P.nvalid[traj] = SOA ? nv : (uint32_t)it;
                if (REC && P.rec_mode == REC_COUNT) P.rec_counts[traj] = (uint32_t)wpos;
                tot_acc += acc; tot_rej += rej; tot_nfe += nfe;
                active = false;
            }
        }
    }
#pragma unroll (N <= 8 ? N : 1)
    for (int o = 16; o > 0; o >>= 1) {
        tot_acc += __shfl_down_sync(0xffffffffu, tot_acc, o);
        tot_rej += __shfl_down_sync(0xffffffffu, tot_rej, o);
        tot_nfe += __shfl_down_sync(0xffffffffu, tot_nfe, o);
    }
    if (lane == 0) {
        atomicAdd(P.totals + 0, tot_acc); atomicAdd(P.totals + 1, tot_rej); atomicAdd(P.totals + 2, tot_nfe);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// oderosenbrock (problem.rs:560-640): one step per tspan interval, hs = tspan[i+1] - tspan[i].  Reference quirks kept:
// `.take(i - 1)` (:614) leaves the newest stage out of the a/c sums, the c-term is added AFTER the multiplication with
// the inverse (:620-625), and `b` doubles as the time nodes (:593,623).  With these the scheme is not the textbook
// Kaps-Rentrop/Shampine method and diverges on ordinary problems; it is reproduced, not repaired.
// ---------------------------------------------------------------------------------------------------------------
struct RosenbrockCoeffs { double gamma, a[4][4], b[4], c[4][4]; };
struct Kr4 {   // rosenbrock.rs:16-68 (Matrix4::new takes its arguments row-major: a[i][j] = a[(i, j)])
    static constexpr RosenbrockCoeffs data() {
        return RosenbrockCoeffs{0.231,
            {{0., 0., 0., 0.}, {2., 0., 0., 0.}, {4.452470820736, 4.16352878860, 0., 0.}, {4.452470820736, 4.16352878860, 0., 0.}},
            {3.95750374663, 4.62489238836, 0.617477263873, 1.2826129456},
            {{0., 0., 0., 0.}, {-5.07167533877, 0., 0., 0.}, {6.02015272865, 0.1597500684673, 0., 0.}, {-1.856343618677, -8.50538085819, -2.08407513602, 0.0}}};
    }
};
struct S4 {    // rosenbrock.rs:71-117
    static constexpr RosenbrockCoeffs data() {
        return RosenbrockCoeffs{0.5,
            {{0., 0., 0., 0.}, {2., 0., 0., 0.}, {48. / 25., 6. / 250., 0., 0.}, {48. / 25., 6. / 250., 0., 0.}},
            {19. / 9., 1. / 2., 25. / 108., 125. / 108.},
            {{0., 0., 0., 0.}, {-8., 0., 0., 0.}, {372. / 25., 12. / 5., 0., 0.}, {-112. / 125., -54. / 125., -2. / 5., 0.}}};
    }
};

template <class RC, class F, bool PT, bool ALL>
__global__ void __launch_bounds__(BLOCK) rosenbrock_kernel(const FixedParams P) {
    constexpr int N = F::N;
    constexpr RosenbrockCoeffs co = RC::data();
    if (F::USES_SINCOS) { deqm::sincos_stage(); __syncthreads(); }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long traj = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x;
    unsigned long long steps = 0, nfe = 0;
    if (traj < P.ntraj) {
        double prm[F::NP ? F::NP : 1];
        load_params<F, PT>(prm, P.params, P.shared, traj, P.ntraj);
        double x[N];
#pragma unroll (N <= 8 ? N : 1)
        for (int d = 0; d < N; ++d) x[d] = P.y0[(unsigned long long)d * P.ntraj + traj];
        if (ALL) {
#pragma unroll (N <= 8 ? N : 1)
            for (int d = 0; d < N; ++d) P.yall[((unsigned long long)d * P.nt) * P.ntraj + traj] = x[d];
        }
        uint8_t st = ST_DONE;
        double ts = P.nt ? __ldg(P.tspan) : 0.0;
        for (unsigned long long i = 0; i + 1 < P.nt; ++i) {
            const double tnext = __ldg(P.tspan + i + 1);
            const double hs = tnext - ts;
            double ftx[N], dfdx[N][N], jac[N][N], jinv[N][N];
            F::eval(ts, x, ftx, prm);
            fdjacobian<F>(ts, x, ftx, prm, dfdx);
            nfe += N + 1;
            const double dg = 1. / (co.gamma * hs);
#pragma unroll (N <= 8 ? N : 1)
            for (int r = 0; r < N; ++r) {
#pragma unroll (N <= 8 ? N : 1)
                for (int c = 0; c < N; ++c) jac[r][c] = (r == c ? dg : 0.0) - dfdx[r][c];
            }
            if (!na_try_inverse<N>(jac, jinv)) { st = ST_INVALID_MATRIX; break; }
            double g[4][N], fx[N], nx[N];
            F::eval(ts + co.b[0] * hs, x, fx, prm);
            na_gemv<N>(jinv, fx, g[0]);
#pragma unroll (N <= 8 ? N : 1)
            for (int d = 0; d < N; ++d) nx[d] = x[d] + g[0][d] * co.b[0];
#pragma unroll (N <= 8 ? N : 1)
            for (int s = 1; s < 4; ++s) {
                double dx[N], df[N], xa[N];
#pragma unroll (N <= 8 ? N : 1)
                for (int d = 0; d < N; ++d) { dx[d] = 0.0; df[d] = 0.0; }
#pragma unroll (N <= 8 ? N : 1)
                for (int j = 0; j < s - 1; ++j) {
#pragma unroll (N <= 8 ? N : 1)
                    for (int d = 0; d < N; ++d) { dx[d] = dx[d] + g[j][d] * co.a[s][j]; df[d] = df[d] + g[j][d] * co.c[s][j]; }
                }
#pragma unroll (N <= 8 ? N : 1)
                for (int d = 0; d < N; ++d) xa[d] = x[d] + dx[d];
                F::eval(ts + co.b[s] * hs, xa, fx, prm);
                na_gemv<N>(jinv, fx, g[s]);
                const double ih = 1. / hs;
#pragma unroll (N <= 8 ? N : 1)
                for (int d = 0; d < N; ++d) { g[s][d] = g[s][d] + df[d] * ih; nx[d] = nx[d] + g[s][d] * co.b[s]; }
            }
            nfe += 4;
#pragma unroll (N <= 8 ? N : 1)
            for (int d = 0; d < N; ++d) x[d] = nx[d];
            if (ALL) {
#pragma unroll (N <= 8 ? N : 1)
                for (int d = 0; d < N; ++d) P.yall[((unsigned long long)d * P.nt + (i + 1)) * P.ntraj + traj] = x[d];
            }
            ++steps;
            ts = tnext;
        }
#pragma unroll (N <= 8 ? N : 1)
        for (int d = 0; d < N; ++d) P.yfinal[(unsigned long long)d * P.ntraj + traj] = x[d];
        P.accepted[traj] = (uint32_t)steps; P.rejected[traj] = 0u; P.nfe[traj] = (uint32_t)nfe; P.status[traj] = st;
        P.t_reached[traj] = ts; P.nvalid[traj] = (uint32_t)(P.nt ? steps + 1 : 0);
    }
#pragma unroll (N <= 8 ? N : 1)
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_down_sync(0xffffffffu, steps, o);
        nfe += __shfl_down_sync(0xffffffffu, nfe, o);
    }
    if (lane == 0) { atomicAdd(P.totals + 0, steps); atomicAdd(P.totals + 2, nfe); }
}

}  // namespace deqk
