// kernels_common.cu — tableau-independent kernels: hinit pre-pass, AoS<->SoA transposes, device self-tests for the
// libm replicas and the DFMA-rate microbenchmark that provides the FP64 roofline denominator.
// Compiled with -fmad=false like every parity translation unit.
#include "stepper.cuh"
#include "launch.h"

namespace deql {

template <class F>
static cudaError_t launch_hinit_f(int pt, const deqk::HinitParams& P, cudaStream_t st) {
    const unsigned long long grid = (P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK;
    if (pt) deqk::hinit_kernel<F, true><<<(unsigned)grid, deqk::BLOCK, 0, st>>>(P);
    else deqk::hinit_kernel<F, false><<<(unsigned)grid, deqk::BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_hinit(int sys, int pt, const deqk::HinitParams& P, cudaStream_t st) {
    switch (sys) {
        case 0: return launch_hinit_f<deqr::Lorenz>(pt, P, st);
        case 1: return launch_hinit_f<deqr::VanDerPol>(pt, P, st);
        case 2: return launch_hinit_f<deqr::Pendulum>(pt, P, st);
        default: return cudaErrorInvalidValue;
    }
}

// [ntraj][n] -> [n][ntraj].  n is small (2..16): each thread moves one trajectory; the strided side stays within
// a few cache lines per warp, so both sides run at full sector efficiency out of L1/L2.
__global__ void to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, unsigned long long ntraj, int n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntraj) return;
    for (int d = 0; d < n; ++d) soa[(unsigned long long)d * ntraj + i] = aos[i * n + d];
}
__global__ void to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, unsigned long long ntraj, int n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntraj) return;
    for (int d = 0; d < n; ++d) aos[i * n + d] = soa[(unsigned long long)d * ntraj + i];
}
cudaError_t launch_transpose_to_soa(const double* a, double* s, unsigned long long ntraj, int n, cudaStream_t st) {
    if (ntraj == 0) return cudaSuccess;
    to_soa_kernel<<<(unsigned)((ntraj + 255) / 256), 256, 0, st>>>(a, s, ntraj, n);
    return cudaGetLastError();
}
cudaError_t launch_transpose_to_aos(const double* s, double* a, unsigned long long ntraj, int n, cudaStream_t st) {
    if (ntraj == 0) return cudaSuccess;
    to_aos_kernel<<<(unsigned)((ntraj + 255) / 256), 256, 0, st>>>(s, a, ntraj, n);
    return cudaGetLastError();
}

__global__ void test_pow_kernel(const double* x, const double* y, double* out, unsigned long long n) {
    __shared__ uint64_t s_exp[256];
    __shared__ double s_powlog[384];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exp[i] = deqk::g_exp_tab[i];
    for (int i = threadIdx.x; i < 384; i += blockDim.x) s_powlog[i] = deqk::g_powlog_tab[i];
    __syncthreads();
    const deqm::Tables T{s_exp, s_powlog, nullptr};
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = deqm::pow_glibc(x[i], y[i], T);
}
__global__ void test_log10_kernel(const double* x, double* out, unsigned long long n) {
    const deqm::Tables T{deqk::g_exp_tab, deqk::g_powlog_tab, deqk::g_log_tab};
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = deqm::log10_glibc(x[i], T);
}
cudaError_t launch_test_pow(const double* x, const double* y, double* out, unsigned long long n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    test_pow_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(x, y, out, n);
    return cudaGetLastError();
}
cudaError_t launch_test_log10(const double* x, double* out, unsigned long long n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    test_log10_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(x, out, n);
    return cudaGetLastError();
}

// Ensemble totals of the per-trajectory counters (the Diagnostics the reference counts, problem.rs:957-970): a grid-stride
// sum with a warp-shuffle reduction and one atomic per warp.  Runs after the solve, on demand (deq_result_totals), so the
// adaptive kernel does not carry three 64-bit accumulators through its hot loop (six registers at the 128-register cap).
// HBM-bound: 12 bytes per trajectory, ~2 us per 2^20 trajectories.
__global__ void __launch_bounds__(256) sum_counts_kernel(const uint32_t* __restrict__ acc, const uint32_t* __restrict__ rej,
                                                         const uint32_t* __restrict__ nfe, unsigned long long n,
                                                         unsigned long long* __restrict__ totals) {
    unsigned long long a = 0, r = 0, f = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        a += acc[i]; r += rej[i]; f += nfe[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o); r += __shfl_down_sync(0xffffffffu, r, o); f += __shfl_down_sync(0xffffffffu, f, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(totals + 0, a); atomicAdd(totals + 1, r); atomicAdd(totals + 2, f); }
}
cudaError_t launch_sum_counts(const uint32_t* acc, const uint32_t* rej, const uint32_t* nfe, unsigned long long n,
                              unsigned long long* totals, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long want = (n + 255) / 256, cap = (unsigned long long)sms * 8;
    sum_counts_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(acc, rej, nfe, n, totals);
    return cudaGetLastError();
}

// Self-test of div_main / rcp_main (deq_math.cuh) against the IEEE operators: operand pairs from a counter hash — random
// mantissas, exponents spread over the whole double range in a quarter of the cases and over the range the step-size
// controller produces in the rest — compared bit for bit wherever the main-path flag stays clear.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) test_div_kernel(unsigned long long seed, unsigned long long count, unsigned long long* out) {
    unsigned long long mism = 0, fast = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long a = mix64(seed + 2 * i), b = mix64(seed + 2 * i + 1), c = mix64(a ^ b);
        unsigned long long na = a, nb = b;
        if (c & 3ull) {   // controller-like magnitudes: numerator 2^-60 .. 2^20, denominator 2^-40 .. 2^20
            na = (a & 0x800fffffffffffffull) | ((unsigned long long)(1023 - 60 + (int)((c >> 8) % 80)) << 52);
            nb = (b & 0x000fffffffffffffull) | ((unsigned long long)(1023 - 40 + (int)((c >> 20) % 60)) << 52);
        }
        const double n = __longlong_as_double((long long)na), d = __longlong_as_double((long long)nb);
        bool bad = false;
        const double q = deqm::div_main(n, d, bad);
        if (!bad) { ++fast; if (__double_as_longlong(q) != __double_as_longlong(n / d)) ++mism; }
        bool bad2 = false;
        const double r = deqm::rcp_main(d, bad2);
        if (!bad2) { ++fast; if (__double_as_longlong(r) != __double_as_longlong(1.0 / d)) ++mism; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mism += __shfl_down_sync(0xffffffffu, mism, o); fast += __shfl_down_sync(0xffffffffu, fast, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out + 0, mism); atomicAdd(out + 1, fast); }
}
cudaError_t launch_test_div(unsigned long long seed, unsigned long long count, unsigned long long* d_out, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    test_div_kernel<<<sms * 8, 256, 0, st>>>(seed, count, d_out);
    return cudaGetLastError();
}

// How alike are neighbouring trajectories?  Fraction-of-pairs statistic on the first step sizes hinit produced: pairs (i, i + 1)
// inside a 32-block whose h0 differ by more than 1/16.  A parameter-sorted sweep has (almost) none; an ensemble in arbitrary
// order has most.  deq_solve uses it to pick the SoA dense kernel (lane-refill for coherent neighbours, row-synchronous
// otherwise) when the caller leaves the choice open.  HBM-bound: reads 8 bytes per trajectory.
__global__ void __launch_bounds__(256) coherence_kernel(const double* __restrict__ h0, unsigned long long n, unsigned long long* __restrict__ out) {
    unsigned long long differ = 0, pairs = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        if ((i & 31ull) == 31ull) continue;
        const double a = h0[i], b = h0[i + 1];
        ++pairs;
        if (!(fabs(a - b) <= 0.0625 * fabs(a))) ++differ;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { differ += __shfl_down_sync(0xffffffffu, differ, o); pairs += __shfl_down_sync(0xffffffffu, pairs, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out + 0, differ); atomicAdd(out + 1, pairs); }
}
cudaError_t launch_coherence(const double* h0, unsigned long long n, unsigned long long* d_out, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long want = (n + 255) / 256, cap = (unsigned long long)sms * 8;
    coherence_kernel<<<(unsigned)(want < cap ? (want ? want : 1) : cap), 256, 0, st>>>(h0, n, d_out);
    return cudaGetLastError();
}

// DFMA stream: 8 independent accumulator chains per thread, 2 flop per FMA.  The result is written so the loop
// cannot be removed.  Used to measure the chip's sustained FP64 rate (MEASURED_PEAKS.json has no FP64 figure).
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
cudaError_t launch_dfma_peak(double* sink, int iters, int blocks, cudaStream_t st) {
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(sink, iters);
    return cudaGetLastError();
}

}  // namespace deql
