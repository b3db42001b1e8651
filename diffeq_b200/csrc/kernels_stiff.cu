// kernels_stiff.cu — instantiates the stiff solvers (stiff.cuh: ode23s, oderosenbrock kr4 / s4) for the built-in systems
// and exposes host launchers.  Compiled with -fmad=false like every parity object.
#include "stiff.cuh"
#include "launch.h"

namespace deql {

template <class K>
static int persistent_grid(K kernel, unsigned long long ntraj) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, deqk::BLOCK, 0);
    if (per_sm < 1) per_sm = 1;
    const unsigned long long want = (ntraj + deqk::BLOCK - 1) / deqk::BLOCK, cap = (unsigned long long)sms * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

#ifdef DEQ_FAST_TU
constexpr bool kFast = true;     // this object is compiled with -fmad=true
#define DEQ_LAUNCH_ODE23S launch_ode23s_fast
#else
constexpr bool kFast = false;
#define DEQ_LAUNCH_ODE23S launch_ode23s
#endif

template <class F, bool PT, int MODE>
static cudaError_t launch_ode23s_t(const deqk::AdaptParams& P, cudaStream_t st) {
    auto kernel = deqk::ode23s_kernel<F, PT, MODE, kFast>;
    kernel<<<persistent_grid(kernel, P.ntraj), deqk::BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

template <class F>
static cudaError_t launch_ode23s_f(int pt, int mode, const deqk::AdaptParams& P, cudaStream_t st) {
    switch (mode * 2 + (pt ? 1 : 0)) {
        case deqk::MODE_FINAL * 2: return launch_ode23s_t<F, false, deqk::MODE_FINAL>(P, st);
        case deqk::MODE_FINAL * 2 + 1: return launch_ode23s_t<F, true, deqk::MODE_FINAL>(P, st);
        case deqk::MODE_DENSE * 2: return launch_ode23s_t<F, false, deqk::MODE_DENSE>(P, st);
        case deqk::MODE_DENSE * 2 + 1: return launch_ode23s_t<F, true, deqk::MODE_DENSE>(P, st);
        case deqk::MODE_REC * 2: return launch_ode23s_t<F, false, deqk::MODE_REC>(P, st);
        case deqk::MODE_REC * 2 + 1: return launch_ode23s_t<F, true, deqk::MODE_REC>(P, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t DEQ_LAUNCH_ODE23S(int sys, int pt, int mode, const deqk::AdaptParams& P, cudaStream_t st) {
    switch (sys) {
        case 0: return launch_ode23s_f<deqr::Lorenz>(pt, mode, P, st);
        case 1: return launch_ode23s_f<deqr::VanDerPol>(pt, mode, P, st);
        case 2: return launch_ode23s_f<deqr::Pendulum>(pt, mode, P, st);
        default: return cudaErrorInvalidValue;
    }
}

#ifndef DEQ_FAST_TU
template <class RC, class F>
static cudaError_t launch_rosenbrock_f(int pt, int all, const deqk::FixedParams& P, cudaStream_t st) {
    const unsigned grid = (unsigned)((P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK);
    if (pt) { if (all) deqk::rosenbrock_kernel<RC, F, true, true><<<grid, deqk::BLOCK, 0, st>>>(P); else deqk::rosenbrock_kernel<RC, F, true, false><<<grid, deqk::BLOCK, 0, st>>>(P); }
    else    { if (all) deqk::rosenbrock_kernel<RC, F, false, true><<<grid, deqk::BLOCK, 0, st>>>(P); else deqk::rosenbrock_kernel<RC, F, false, false><<<grid, deqk::BLOCK, 0, st>>>(P); }
    return cudaGetLastError();
}

template <class RC>
static cudaError_t launch_rosenbrock_c(int sys, int pt, int all, const deqk::FixedParams& P, cudaStream_t st) {
    switch (sys) {
        case 0: return launch_rosenbrock_f<RC, deqr::Lorenz>(pt, all, P, st);
        case 1: return launch_rosenbrock_f<RC, deqr::VanDerPol>(pt, all, P, st);
        case 2: return launch_rosenbrock_f<RC, deqr::Pendulum>(pt, all, P, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_rosenbrock(int which, int sys, int pt, int all, const deqk::FixedParams& P, cudaStream_t st) {
    return which == 0 ? launch_rosenbrock_c<deqk::Kr4>(sys, pt, all, P, st) : launch_rosenbrock_c<deqk::S4>(sys, pt, all, P, st);
}

void rosenbrock_info(int which, double* gamma, double* a, double* b, double* c) {
    auto fill = [&](auto rc) {
        constexpr deqk::RosenbrockCoeffs co = decltype(rc)::data();
        *gamma = co.gamma;
        for (int i = 0; i < 4; ++i) { b[i] = co.b[i]; for (int j = 0; j < 4; ++j) { a[4 * i + j] = co.a[i][j]; c[4 * i + j] = co.c[i][j]; } }
    };
    if (which == 0) fill(deqk::Kr4{}); else fill(deqk::S4{});
}
#endif  // !DEQ_FAST_TU

}  // namespace deql
