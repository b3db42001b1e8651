// rhs.cuh — built-in right-hand sides as device functors.
//
// The reference takes the RHS as a closure `F: Fn(f64, &Y) -> Y` (/root/reference/src/ode/problem.rs:32-36) that
// captures its parameters; a Rust closure cannot run on the device, so the RHS is a device functor
// `eval(t, y[N], dy[N], p[NP])` with the captured constants passed as the parameter vector `p`
// (shared by the ensemble or one row per trajectory).  User systems come in as CUDA source through NVRTC with the
// same signature (see nvrtc_rhs.cpp).
#pragma once
#include "deq_rtc_compat.h"
#include "deq_sincos.cuh"
#if defined(__CUDACC__)
#define DEQ_RHS_HD __host__ __device__ __forceinline__
#else
#define DEQ_RHS_HD inline
#endif

// sin / cos with glibc's bits (what `f64::sin` / `f64::cos` in a reference closure evaluate to); user NVRTC sources
// may call these too when they need bit parity with a Rust right-hand side.
// In a FAST-mode translation unit (DEQ_FAST_TU) they are the platform's sin / cos instead.  Not inlined on the device:
// a 13-stage tableau would otherwise carry 26 copies of the four-way range dispatch (~300 KB of SASS) and thrash the
// instruction cache (measured 3x slower than one shared copy behind a call).
#if defined(__CUDACC__) && defined(DEQ_FAST_TU)
#define DEQ_TRIG_HD static __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define DEQ_TRIG_HD static __host__ __device__ __noinline__
#else
#define DEQ_TRIG_HD inline
#endif
#if defined(__CUDACC__) && defined(DEQ_FAST_TU)
// FAST arithmetic: sin / cos for |x| < 1e5 as a three-term Cody-Waite reduction by pi/2 (FMA, exact products) and the
// fdlibm kernel polynomials on |r| <= pi/4, both evaluated and selected by the quadrant — 20 FP64 instructions, no table, no
// Payne-Hanek slow path in line (CUDA's general sin/cos: ~40 FP64 plus the range dispatch).  Error < 2 ulp on that range (like CUDA's);
// larger arguments, inf and NaN go to the platform routines.  phase = 0: sin(x), 1: cos(x).
static __constant__ double c_trigk[16] = {
    6.36619772367581382433e-01 /* 2/pi */, 1.57079632673412561417e+00, 6.07710050630396597660e-11, 2.02226624879595063154e-21,
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,
    -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,
    2.08757232129817482790e-09, -1.13596475577881948265e-11};
// (out of line: 26 inlined copies of the platform routines' range dispatch and Payne-Hanek call would quadruple the code of a
// 13-stage kernel and spread its hot path over the instruction cache)
static __device__ __noinline__ double deq_trig_slow(double x, int phase) { return phase ? cos(x) : sin(x); }
static __device__ __forceinline__ double deq_trig_fast(double x, int phase) {
    const double* K = c_trigk;
    if (!(fabs(x) < 1.0e5)) return deq_trig_slow(x, phase);
    double q = __fma_rn(x, K[0], 6755399441055744.0);
    const int n = __double2loint(q) + phase;
    q -= 6755399441055744.0;
    double r = __fma_rn(q, -K[1], x);
    r = __fma_rn(q, -K[2], r);
    r = __fma_rn(q, -K[3], r);
    const double r2 = r * r;
    double ps = __fma_rn(r2, K[9], K[8]);
    double pc = __fma_rn(r2, K[15], K[14]);
    ps = __fma_rn(r2, ps, K[7]); pc = __fma_rn(r2, pc, K[13]);
    ps = __fma_rn(r2, ps, K[6]); pc = __fma_rn(r2, pc, K[12]);
    ps = __fma_rn(r2, ps, K[5]); pc = __fma_rn(r2, pc, K[11]);
    ps = __fma_rn(r2, ps, K[4]); pc = __fma_rn(r2, pc, K[10]);
    const double vs = __fma_rn(r * r2, ps, r);
    const double vc = __fma_rn(r2 * r2, pc, __fma_rn(r2, -0.5, 1.0));
    const double v = (n & 1) ? vc : vs;
    return __hiloint2double(__double2hiint(v) ^ ((n & 2) << 30), __double2loint(v));
}
#endif

DEQ_TRIG_HD double deq_sin(double x) {
#if defined(__CUDA_ARCH__) && defined(DEQ_FAST_TU)
    return deq_trig_fast(x, 0);
#elif defined(__CUDA_ARCH__)
    return deqm::sin_glibc(x, deqm::SinCosShared());
#else
    return deqm::sin_glibc(x, deqm::SinCosHost());
#endif
}
DEQ_TRIG_HD double deq_cos(double x) {
#if defined(__CUDA_ARCH__) && defined(DEQ_FAST_TU)
    return deq_trig_fast(x, 1);
#elif defined(__CUDA_ARCH__)
    return deqm::cos_glibc(x, deqm::SinCosShared());
#else
    return deqm::cos_glibc(x, deqm::SinCosHost());
#endif
}

// sin(xs) and cos(xc) behind ONE call: half the call overhead of the pair above, and the two independent evaluations
// interleave (the FP64 pipe has an 8-cycle dependent latency).  Same bits as deq_sin / deq_cos.
struct DeqSinCos { double s, c; };
DEQ_TRIG_HD DeqSinCos deq_sin_cos(double xs, double xc) {
#if defined(__CUDA_ARCH__) && defined(DEQ_FAST_TU)
    return DeqSinCos{deq_trig_fast(xs, 0), deq_trig_fast(xc, 1)};
#elif defined(__CUDA_ARCH__)
    DeqSinCos r;
    deqm::sin_cos_glibc(xs, xc, deqm::SinCosShared(), r.s, r.c);
    return r;
#else
    DeqSinCos r;
    deqm::sin_cos_glibc(xs, xc, deqm::SinCosHost(), r.s, r.c);
    return r;
#endif
}

namespace deqr {

// Lorenz-63, operation order of the reference's test fixture (problem.rs:989-1000); p = {sigma, rho, beta}.
struct Lorenz {
    static constexpr int N = 3, NP = 3;
    static constexpr bool USES_SINCOS = false;
    DEQ_RHS_HD static void eval(double, const double* v, double* d, const double* p) {
        const double x = v[0], y = v[1], z = v[2];
        d[0] = p[0] * (y - x);
        d[1] = x * (p[1] - z) - y;
        d[2] = x * y - p[2] * z;
    }
    template <class R>  // generic-precision form for the optional f32 mode
    DEQ_RHS_HD static void eval_r(R, const R* v, R* d, const R* p) {
        const R x = v[0], y = v[1], z = v[2];
        d[0] = p[0] * (y - x);
        d[1] = x * (p[1] - z) - y;
        d[2] = x * y - p[2] * z;
    }
};

// Van der Pol  x' = y, y' = mu (1 - x^2) y - x ; p = {mu}.
struct VanDerPol {
    static constexpr int N = 2, NP = 1;
    static constexpr bool USES_SINCOS = false;
    DEQ_RHS_HD static void eval(double, const double* v, double* d, const double* p) {
        const double x = v[0], y = v[1];
        d[0] = y;
        d[1] = p[0] * (1.0 - x * x) * y - x;
    }
    template <class R>
    DEQ_RHS_HD static void eval_r(R, const R* v, R* d, const R* p) {
        const R x = v[0], y = v[1];
        d[0] = y;
        d[1] = p[0] * (R(1) - x * x) * y - x;
    }
};

// Damped driven pendulum  th' = w, w' = -gamma w - sin th + A cos(Omega t) ; p = {gamma, A, Omega}.
struct Pendulum {
    static constexpr int N = 2, NP = 3;
    static constexpr bool USES_SINCOS = true;
    DEQ_RHS_HD static void eval(double t, const double* v, double* d, const double* p) {
        const double th = v[0], w = v[1];
        const DeqSinCos sc = deq_sin_cos(th, p[2] * t);
        d[0] = w;
        d[1] = -p[0] * w - sc.s + p[1] * sc.c;
    }
    // the time-only term on its own (stepper.cuh `stages`: evaluated once per distinct stage time)
    static constexpr int NG = 1;
    DEQ_RHS_HD static void forcing(double t, double* g, const double* p) { g[0] = deq_cos(p[2] * t); }
    DEQ_RHS_HD static void eval_g(double, const double* v, double* d, const double* p, const double* g) {
        const double w = v[1];
        d[0] = w;
        d[1] = -p[0] * w - deq_sin(v[0]) + p[1] * g[0];
    }
    DEQ_RHS_HD static void eval_r(float t, const float* v, float* d, const float* p) {
        const float th = v[0], w = v[1];
        d[0] = w;
        d[1] = -p[0] * w - sinf(th) + p[1] * cosf(p[2] * t);
    }
};

}  // namespace deqr
