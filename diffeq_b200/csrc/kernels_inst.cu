// kernels_inst.cu — instantiates the steppers of ONE Butcher tableau (selected with -DDEQ_TB=<deql::TbId>) for the
// built-in systems and exposes host launchers.  Compiled with -fmad=false (parity contract, see stepper.cuh).
#include "stepper.cuh"
#ifdef DEQ_FAST_TU
#include "stepper_f32.cuh"
#endif
#include "launch.h"

#ifndef DEQ_TB
#error "compile with -DDEQ_TB=<tableau id>"
#endif

namespace deql {

#if DEQ_TB == 0
using TB = deqt::Feuler;
#elif DEQ_TB == 1
using TB = deqt::Midpoint;
#elif DEQ_TB == 2
using TB = deqt::Heun;
#elif DEQ_TB == 3
using TB = deqt::Rk4;
#elif DEQ_TB == 4
using TB = deqt::Rk21<false>;
#elif DEQ_TB == 5
using TB = deqt::Rk21<true>;
#elif DEQ_TB == 6
using TB = deqt::Rk23<false>;
#elif DEQ_TB == 7
using TB = deqt::Rk23<true>;
#elif DEQ_TB == 8
using TB = deqt::Rk45;
#elif DEQ_TB == 9
using TB = deqt::Dopri5<false>;
#elif DEQ_TB == 10
using TB = deqt::Dopri5<true>;
#elif DEQ_TB == 11
using TB = deqt::Feh78<false>;
#elif DEQ_TB == 12
using TB = deqt::Feh78<true>;
#endif

#define DEQ_CAT2(a, b) a##b
#define DEQ_CAT(a, b) DEQ_CAT2(a, b)

template <class K>
static int persistent_grid(K kernel, unsigned long long ntraj, size_t dyn_smem = 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, deqk::BLOCK, dyn_smem);
    if (per_sm < 1) per_sm = 1;
    unsigned long long want = (ntraj + deqk::BLOCK - 1) / deqk::BLOCK;
    unsigned long long cap = (unsigned long long)sms * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

#ifdef DEQ_FAST_TU
[[maybe_unused]] constexpr bool kFast = true;    // this object is compiled with -fmad=true: FMA contraction everywhere
#define DEQ_LAUNCH_ADAPT DEQ_CAT(launch_adapt_fast_, DEQ_TB)
#else
[[maybe_unused]] constexpr bool kFast = false;   // parity object, -fmad=false
#define DEQ_LAUNCH_ADAPT DEQ_CAT(launch_adapt_, DEQ_TB)
#endif

template <class F, bool PT, int MODE, bool SPEC = false>
static cudaError_t launch_adapt_t(const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    if constexpr (TB::ADAPTIVE && !(SPEC && kFast)) {
        auto kernel = deqk::adapt_kernel<TB, F, PT, MODE, kFast, SPEC>;
        const size_t smem = (MODE == deqk::MODE_DENSE_TM || (MODE == deqk::MODE_REC && P.rec_mode == deqk::REC_WRITE)) ? deqk::STAGE_BYTES : 0;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int grid = grid_blocks > 0 ? grid_blocks : persistent_grid(kernel, P.ntraj, smem);
        kernel<<<grid, deqk::BLOCK, smem, st>>>(P);
        return cudaGetLastError();
    } else {
        (void)P; (void)grid_blocks; (void)st;
        return cudaErrorInvalidValue;
    }
}

template <class F, bool PT>
static cudaError_t launch_adapt_rows_t(const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    if constexpr (TB::ADAPTIVE) {
        auto kernel = deqk::adapt_rows_kernel<TB, F, PT, kFast>;
        const int grid = grid_blocks > 0 ? grid_blocks : persistent_grid(kernel, P.ntraj);
        kernel<<<grid, deqk::BLOCK, 0, st>>>(P);
        return cudaGetLastError();
    } else {
        (void)P; (void)grid_blocks; (void)st;
        return cudaErrorInvalidValue;
    }
}

template <class F>
static cudaError_t launch_adapt_f(int pt, int mode, const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    if (mode == deqk::MODE_DENSE_ROWS) return pt ? launch_adapt_rows_t<F, true>(P, grid_blocks, st) : launch_adapt_rows_t<F, false>(P, grid_blocks, st);
    switch (mode * 2 + (pt ? 1 : 0)) {
        case deqk::MODE_FINAL * 2: return launch_adapt_t<F, false, deqk::MODE_FINAL>(P, grid_blocks, st);
        case deqk::MODE_FINAL * 2 + 1: return launch_adapt_t<F, true, deqk::MODE_FINAL>(P, grid_blocks, st);
        case deqk::MODE_DENSE * 2: return launch_adapt_t<F, false, deqk::MODE_DENSE>(P, grid_blocks, st);
        case deqk::MODE_DENSE * 2 + 1: return launch_adapt_t<F, true, deqk::MODE_DENSE>(P, grid_blocks, st);
        case deqk::MODE_DENSE_TM * 2: return launch_adapt_t<F, false, deqk::MODE_DENSE_TM>(P, grid_blocks, st);
        case deqk::MODE_DENSE_TM * 2 + 1: return launch_adapt_t<F, true, deqk::MODE_DENSE_TM>(P, grid_blocks, st);
        case deqk::MODE_REC * 2: return launch_adapt_t<F, false, deqk::MODE_REC>(P, grid_blocks, st);
        case deqk::MODE_REC * 2 + 1: return launch_adapt_t<F, true, deqk::MODE_REC>(P, grid_blocks, st);
        case deqk::MODE_REC_SPEC * 2: return launch_adapt_t<F, false, deqk::MODE_REC, true>(P, grid_blocks, st);       // Points::Specified
        case deqk::MODE_REC_SPEC * 2 + 1: return launch_adapt_t<F, true, deqk::MODE_REC, true>(P, grid_blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t DEQ_LAUNCH_ADAPT(int sys, int pt, int mode, const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    switch (sys) {
        case 0: return launch_adapt_f<deqr::Lorenz>(pt, mode, P, grid_blocks, st);
        case 1: return launch_adapt_f<deqr::VanDerPol>(pt, mode, P, grid_blocks, st);
        case 2: return launch_adapt_f<deqr::Pendulum>(pt, mode, P, grid_blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

#ifdef DEQ_FAST_TU
// optional single-precision variant (stepper_f32.cuh), built-in systems only
template <class F, bool PT, int MODE>
static cudaError_t launch_adapt_f32_t(const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    if constexpr (TB::ADAPTIVE) {
        auto kernel = deqk::adapt_kernel_f32<TB, F, PT, MODE>;
        const int grid = grid_blocks > 0 ? grid_blocks : persistent_grid(kernel, P.ntraj);
        kernel<<<grid, deqk::BLOCK, 0, st>>>(P);
        return cudaGetLastError();
    } else {
        (void)P; (void)grid_blocks; (void)st;
        return cudaErrorInvalidValue;
    }
}
template <class F>
static cudaError_t launch_adapt_f32_f(int pt, int mode, const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    if (pt) return mode == deqk::MODE_DENSE ? launch_adapt_f32_t<F, true, deqk::MODE_DENSE>(P, grid_blocks, st)
                                            : launch_adapt_f32_t<F, true, deqk::MODE_FINAL>(P, grid_blocks, st);
    return mode == deqk::MODE_DENSE ? launch_adapt_f32_t<F, false, deqk::MODE_DENSE>(P, grid_blocks, st)
                                    : launch_adapt_f32_t<F, false, deqk::MODE_FINAL>(P, grid_blocks, st);
}
cudaError_t DEQ_CAT(launch_adapt_f32_, DEQ_TB)(int sys, int pt, int mode, const deqk::AdaptParams& P, int grid_blocks, cudaStream_t st) {
    switch (sys) {
        case 0: return launch_adapt_f32_f<deqr::Lorenz>(pt, mode, P, grid_blocks, st);
        case 1: return launch_adapt_f32_f<deqr::VanDerPol>(pt, mode, P, grid_blocks, st);
        case 2: return launch_adapt_f32_f<deqr::Pendulum>(pt, mode, P, grid_blocks, st);
        default: return cudaErrorInvalidValue;
    }
}
#endif

#ifndef DEQ_FAST_TU
template <class F, bool PT, bool ALL>
static cudaError_t launch_fixed_t(const deqk::FixedParams& P, cudaStream_t st) {
    if constexpr (!TB::ADAPTIVE) {
        const unsigned long long grid = (P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK;
        deqk::fixed_kernel<TB, F, PT, ALL><<<(unsigned)grid, deqk::BLOCK, 0, st>>>(P);
        return cudaGetLastError();
    } else {
        (void)P; (void)st;
        return cudaErrorInvalidValue;
    }
}

template <class F>
static cudaError_t launch_fixed_f(int pt, int all, const deqk::FixedParams& P, cudaStream_t st) {
    if (pt) return all ? launch_fixed_t<F, true, true>(P, st) : launch_fixed_t<F, true, false>(P, st);
    return all ? launch_fixed_t<F, false, true>(P, st) : launch_fixed_t<F, false, false>(P, st);
}

cudaError_t DEQ_CAT(launch_fixed_, DEQ_TB)(int sys, int pt, int all, const deqk::FixedParams& P, cudaStream_t st) {
    switch (sys) {
        case 0: return launch_fixed_f<deqr::Lorenz>(pt, all, P, st);
        case 1: return launch_fixed_f<deqr::VanDerPol>(pt, all, P, st);
        case 2: return launch_fixed_f<deqr::Pendulum>(pt, all, P, st);
        default: return cudaErrorInvalidValue;
    }
}

void DEQ_CAT(info_, DEQ_TB)(TableauInfo* out) {
    constexpr auto tb = TB::data();
    out->S = TB::S; out->adaptive = TB::ADAPTIVE; out->order_min = TB::ORDER_MIN;
    out->fsal = deqt::fsal_of(tb);
    for (int r = 0; r < TB::S; ++r) {
        for (int c = 0; c < TB::S; ++c) out->a[r * TB::S + c] = tb.a[r][c];
        out->b[2 * r] = tb.b[r][0]; out->b[2 * r + 1] = tb.b[r][1]; out->c[r] = tb.c[r];
    }
}

#endif  // !DEQ_FAST_TU

}  // namespace deql
