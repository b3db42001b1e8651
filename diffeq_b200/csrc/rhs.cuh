// rhs.cuh — built-in right-hand sides as device functors.
//
// The reference takes the RHS as a closure `F: Fn(f64, &Y) -> Y` (/root/reference/src/ode/problem.rs:32-36) that
// captures its parameters; a Rust closure cannot run on the device, so the RHS is a device functor
// `eval(t, y[N], dy[N], p[NP])` with the captured constants passed as the parameter vector `p`
// (shared by the ensemble or one row per trajectory).  User systems come in as CUDA source through NVRTC with the
// same signature (see nvrtc_rhs.cpp).
#pragma once
#include "deq_rtc_compat.h"
#if defined(__CUDACC__)
#define DEQ_RHS_HD __host__ __device__ __forceinline__
#else
#define DEQ_RHS_HD inline
#endif

namespace deqr {

// Lorenz-63, operation order of the reference's test fixture (problem.rs:989-1000); p = {sigma, rho, beta}.
struct Lorenz {
    static constexpr int N = 3, NP = 3;
    DEQ_RHS_HD static void eval(double, const double* v, double* d, const double* p) {
        const double x = v[0], y = v[1], z = v[2];
        d[0] = p[0] * (y - x);
        d[1] = x * (p[1] - z) - y;
        d[2] = x * y - p[2] * z;
    }
};

// Van der Pol  x' = y, y' = mu (1 - x^2) y - x ; p = {mu}.
struct VanDerPol {
    static constexpr int N = 2, NP = 1;
    DEQ_RHS_HD static void eval(double, const double* v, double* d, const double* p) {
        const double x = v[0], y = v[1];
        d[0] = y;
        d[1] = p[0] * (1.0 - x * x) * y - x;
    }
};

// Damped driven pendulum  th' = w, w' = -gamma w - sin th + A cos(Omega t) ; p = {gamma, A, Omega}.
struct Pendulum {
    static constexpr int N = 2, NP = 3;
    DEQ_RHS_HD static void eval(double t, const double* v, double* d, const double* p) {
        const double th = v[0], w = v[1];
        d[0] = w;
        d[1] = -p[0] * w - sin(th) + p[1] * cos(p[2] * t);
    }
};

}  // namespace deqr
