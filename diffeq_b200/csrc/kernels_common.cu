// kernels_common.cu — tableau-independent kernels: hinit pre-pass, AoS<->SoA transposes, device self-tests for the
// libm replicas and the DFMA-rate microbenchmark that provides the FP64 roofline denominator.
// Compiled with -fmad=false like every parity translation unit.
#include "stepper.cuh"
#include "launch.h"
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

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

// ---- first-step ordering (the sorted SoA dense path of deq_solve) --------------------------------------------------------
// An ensemble in arbitrary order defeats the lane-refill kernel's row stores (every lane sits in its own output row).  Sorted by
// the size of the first step hinit chose — a cheap stand-in for "how fast does this trajectory move" — neighbours advance alike
// again (for a parameter sweep the order IS the sweep).  Keys are the bit patterns of |h0| (monotone for non-negative
// doubles; NaN sorts last), values the trajectory indices: one CUB radix sort, 12 bytes per trajectory.
__global__ void __launch_bounds__(256) sort_keys_kernel(const double* __restrict__ h0, unsigned long long n, unsigned long long* __restrict__ keys,
                                                        uint32_t* __restrict__ vals) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (unsigned long long)__double_as_longlong(h0[i]) & 0x7fffffffffffffffull;
    vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) invert_order_kernel(const uint32_t* __restrict__ order, unsigned long long n, uint32_t* __restrict__ inv) {
    const unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) inv[order[s]] = (uint32_t)s;
}
static size_t align256(size_t b) { return (b + 255) / 256 * 256; }
static size_t cub_sort_bytes(unsigned long long n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n);
    return b;
}
size_t sort_by_h0_scratch_bytes(unsigned long long n) { return align256(n * 8) + align256(n * 4) + align256(cub_sort_bytes(n)); }
cudaError_t launch_sort_by_h0(const double* h0, unsigned long long n, uint32_t* order, uint32_t* inv, double* h0_sorted,
                              void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n > 0x7fffffffull || scratch_bytes < sort_by_h0_scratch_bytes(n)) return cudaErrorInvalidValue;
    char* p = (char*)scratch;
    unsigned long long* keys = (unsigned long long*)p; p += align256(n * 8);
    uint32_t* vals = (uint32_t*)p; p += align256(n * 4);
    size_t cb = cub_sort_bytes(n);
    const unsigned grid = (unsigned)((n + 255) / 256);
    sort_keys_kernel<<<grid, 256, 0, st>>>(h0, n, keys, vals);
    // (bits 24 .. 62: the sign is masked off, and the lowest 24 mantissa bits — a relative 2^-28 — cannot matter for "neighbours
    // start alike"; ties keep the caller's order, the sort is stable.  Five 8-bit passes instead of eight.)
    cudaError_t e = cub::DeviceRadixSort::SortPairs((void*)p, cb, (const unsigned long long*)keys, (unsigned long long*)h0_sorted,
                                                    (const uint32_t*)vals, order, (int)n, 24, 63, st);
    if (e != cudaSuccess) return e;
    invert_order_kernel<<<grid, 256, 0, st>>>(order, n, inv);
    return cudaGetLastError();
}

// Columns back into caller order, IN PLACE: rows[q][j] <- rows[q][inv[j]] for nrows rows of ntraj doubles.  A row cannot be
// permuted in place directly (a gather reads what a neighbour overwrites), and a second copy of the rows may not fit (config 3
// is 67 GB), so the rows travel in batches through a small scratch that stays in L2: launch i gathers batch i out of scratch
// (i & 1) back into its rows — coalesced writes, scattered reads that hit L2 — and copies batch i + 1 into the other scratch
// (coalesced both ways).  DRAM sees each byte about once in each direction; the scratch is 2 x batch_rows x ntraj doubles.
// One persistent launch: phase i gathers batch i and copies batch i + 1; the phases are separated by a grid-wide barrier (all
// CTAs are co-resident: the grid is sized from the occupancy query), which costs ~2 us against ~5 us of launch latency and the
// ramp-up / drain of 668 separate launches at 2^20 trajectories.  Scratch reads bypass L1 (__ldcg): other SMs wrote the data
// earlier in this same launch.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned nblocks, unsigned& generation) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++generation;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < generation * nblocks) {}
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256) permute_persistent_kernel(double* rows, const uint32_t* __restrict__ inv, unsigned long long ntraj,
                                                                 unsigned long long chunks, unsigned long long nrows, unsigned long long B,
                                                                 double* scratch, unsigned* barrier) {
    const unsigned long long nb = (nrows + B - 1) / B;
    unsigned generation = 0;
    for (unsigned long long ph = 0; ph <= nb; ++ph) {   // phase ph: gather batch ph - 1, copy batch ph
        const unsigned long long g_row0 = ph ? (ph - 1) * B : 0, g_nrows = ph ? (ph < nb ? B : nrows - (nb - 1) * B) : 0;
        const unsigned long long c_row0 = ph * B, c_nrows = ph < nb ? (ph + 1 < nb ? B : nrows - ph * B) : 0;
        const double* g_scratch = scratch + ((ph + 1) & 1ull) * B * ntraj;   // batch ph - 1 was copied into half (ph - 1) & 1
        double* c_scratch = scratch + (ph & 1ull) * B * ntraj;
        const unsigned long long g_items = g_nrows * chunks, c_items = c_nrows * chunks;
        for (unsigned long long b = blockIdx.x; b < g_items + c_items; b += gridDim.x) {
            const bool gather = b < g_items;
            const unsigned long long bb = gather ? b : b - g_items;
            const unsigned long long q = bb / chunks, j0 = (bb % chunks) * 1024ull;
            if (gather) {
                const double* s = g_scratch + q * ntraj;
                double* d = rows + (g_row0 + q) * ntraj;
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long j = j0 + (unsigned long long)u * 256 + threadIdx.x;
                    if (j < ntraj) v[u] = __ldcg(s + inv[j]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {   // streaming stores: the finished rows must not push the scratch out of L2
                    const unsigned long long j = j0 + (unsigned long long)u * 256 + threadIdx.x;
                    if (j < ntraj) __stcs(d + j, v[u]);
                }
            } else {
                const double* s = rows + (c_row0 + q) * ntraj;
                double* d = c_scratch + q * ntraj;
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long j = j0 + (unsigned long long)u * 256 + threadIdx.x;
                    if (j < ntraj) v[u] = __ldcs(s + j);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long j = j0 + (unsigned long long)u * 256 + threadIdx.x;
                    if (j < ntraj) d[j] = v[u];
                }
            }
        }
        if (ph < nb) grid_barrier(barrier, gridDim.x, generation);
    }
}
unsigned long long permute_rows_batch(unsigned long long ntraj) {
    if (const char* e = getenv("DEQ_PERM_BATCH")) return strtoull(e, nullptr, 10) ? strtoull(e, nullptr, 10) : 1;
    const unsigned long long b = (48ull << 20) / (ntraj * 8ull ? ntraj * 8ull : 1ull);   // ~48 MB per scratch half (measured best of 16 .. 96 MB at 2^20 trajectories: fewer barriers against L2 pressure)
    return b < 1 ? 1 : (b > 64 ? 64 : b);
}
// scratch: 2 * permute_rows_batch(ntraj) * ntraj doubles, followed by one zeroed 32-bit word (the barrier counter)
cudaError_t launch_permute_rows_inplace(double* rows, const uint32_t* inv, unsigned long long nrows, unsigned long long ntraj, double* scratch,
                                        cudaStream_t st) {
    if (nrows == 0 || ntraj == 0) return cudaSuccess;
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, permute_persistent_kernel, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const unsigned long long B = permute_rows_batch(ntraj), chunks = (ntraj + 1023) / 1024;
    unsigned* barrier = (unsigned*)(scratch + 2 * B * ntraj);
    e = cudaMemsetAsync(barrier, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    const unsigned long long cap = (unsigned long long)sms * (unsigned long long)per_sm, want = 2 * B * chunks;
    permute_persistent_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(rows, inv, ntraj, chunks, nrows, B, scratch, barrier);
    return cudaGetLastError();
}

// ---- the same in-place permutation with every global access coalesced (ensembles of 2^16 trajectories and more) ------------
// The gather above moves 8 bytes per 32-byte sector it touches and is bound by the rate of scattered requests.  One permutation
// serves ALL rows, so it is factored once per solve:  slot s --(stage 1)--> intermediate position m(s) --(stage 2)--> column
// order[s], where m is the stable rank of s under the key order[s] / PERM_BK.  `order` being a permutation, destination bucket b
// (columns [PERM_BK b, PERM_BK (b + 1))) owns exactly the intermediate positions with the same numbers, so
//   stage 2 is a permutation INSIDE each PERM_BK-element block: read contiguous, shuffle in shared memory, write contiguous;
//   stage 1 moves a chunk of 8192 consecutive slots (64 KB, staged in shared memory) to n / PERM_BK runs of consecutive intermediate
//           positions (32 doubles on average at 2^20 trajectories), written in ascending order of m (list M1; L1 = the slot
//           inside the chunk that each position takes its value from).
// Plan: two stable radix sorts and three index kernels, 8 bytes of indices per trajectory (L2-resident while the rows stream
// through).  Measured in isolation (tools/micro/perm_bench.cu, profiles/r2_microbench_column_permutation.txt): 3.4 ms per 500
// rows of 2^20 columns against 4.5 ms for the gather form; batching, scratch and barrier as above.
constexpr int PERM_CH = 8192, PERM_BK = 4096, PERM_NT = 512;   // (bucket size: 1024 / 2048 / 4096 measured 3.48 / 3.33 / 3.27 ms per 500 rows — longer runs in stage 1)
__global__ void __launch_bounds__(256) perm_keys_a_kernel(const uint32_t* __restrict__ order, unsigned long long n, uint32_t* __restrict__ keys,
                                                          uint32_t* __restrict__ vals) {
    const unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) { keys[s] = order[s] / PERM_BK; vals[s] = (uint32_t)s; }
}
__global__ void __launch_bounds__(256) perm_keys_b_kernel(const uint32_t* __restrict__ s1, const uint32_t* __restrict__ order, unsigned long long n,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint16_t* __restrict__ d2) {
    const unsigned long long m = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n) { keys[m] = s1[m] / PERM_CH; vals[m] = (uint32_t)m; d2[m] = (uint16_t)(order[s1[m]] % PERM_BK); }
}
__global__ void __launch_bounds__(256) perm_l1_kernel(const uint32_t* __restrict__ s1, const uint32_t* __restrict__ m1, unsigned long long n,
                                                      uint16_t* __restrict__ l1) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) l1[i] = (uint16_t)(s1[m1[i]] % PERM_CH);
}
static size_t cub_sort32_bytes(unsigned long long n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return b;
}
// plan = M1 [n] u32 | L1 [n] u16 | D2 [n] u16 ; scratch for building it: 4 u32 arrays + CUB's
size_t permute_plan_bytes(unsigned long long n) { return align256(n * 4) + 2 * align256(n * 2); }
size_t permute_plan_scratch_bytes(unsigned long long n) { return 4 * align256(n * 4) + align256(cub_sort32_bytes(n)); }
cudaError_t launch_permute_plan(const uint32_t* order, unsigned long long n, void* plan, void* scratch, cudaStream_t st) {
    if (n == 0 || n > 0x7fffffffull) return cudaErrorInvalidValue;
    char* p = (char*)scratch;
    uint32_t* keys = (uint32_t*)p; p += align256(n * 4);
    uint32_t* vals = (uint32_t*)p; p += align256(n * 4);
    uint32_t* keys_out = (uint32_t*)p; p += align256(n * 4);
    uint32_t* s1 = (uint32_t*)p; p += align256(n * 4);
    size_t cb = cub_sort32_bytes(n);
    char* q = (char*)plan;
    uint32_t* m1 = (uint32_t*)q; q += align256(n * 4);
    uint16_t* l1 = (uint16_t*)q; q += align256(n * 2);
    uint16_t* d2 = (uint16_t*)q;
    const unsigned grid = (unsigned)((n + 255) / 256);
    auto bits_for = [](unsigned long long count) { int b = 1; while ((1ull << b) < count) ++b; return b; };   // keys are 0 .. count - 1
    perm_keys_a_kernel<<<grid, 256, 0, st>>>(order, n, keys, vals);
    cudaError_t e = cub::DeviceRadixSort::SortPairs((void*)p, cb, (const uint32_t*)keys, keys_out, (const uint32_t*)vals, s1, (int)n, 0,
                                                    bits_for((n + PERM_BK - 1) / PERM_BK), st);
    if (e != cudaSuccess) return e;
    perm_keys_b_kernel<<<grid, 256, 0, st>>>(s1, order, n, keys, vals, d2);
    e = cub::DeviceRadixSort::SortPairs((void*)p, cb, (const uint32_t*)keys, keys_out, (const uint32_t*)vals, m1, (int)n, 0,
                                        bits_for((n + PERM_CH - 1) / PERM_CH), st);
    if (e != cudaSuccess) return e;
    perm_l1_kernel<<<grid, 256, 0, st>>>(s1, m1, n, l1);
    return cudaGetLastError();
}
// phase ph: stage 2 of batch ph - 1 (scratch -> rows, caller order) + stage 1 of batch ph (rows, slot order -> the other scratch half)
__global__ void __launch_bounds__(PERM_NT) permute_staged_kernel(double* rows, const uint32_t* __restrict__ m1, const uint16_t* __restrict__ l1,
                                                                 const uint16_t* __restrict__ d2, unsigned long long ntraj, unsigned long long nrows,
                                                                 unsigned long long B, double* scratch, unsigned* barrier) {
    extern __shared__ double perm_tile[];   // PERM_CH doubles
    const unsigned long long nb = (nrows + B - 1) / B, nch = (ntraj + PERM_CH - 1) / PERM_CH;
    unsigned generation = 0;
    for (unsigned long long ph = 0; ph <= nb; ++ph) {
        const unsigned long long g_row0 = ph ? (ph - 1) * B : 0, g_nrows = ph ? (ph < nb ? B : nrows - (nb - 1) * B) : 0;
        const unsigned long long c_row0 = ph * B, c_nrows = ph < nb ? (ph + 1 < nb ? B : nrows - ph * B) : 0;
        const double* g_scratch = scratch + ((ph + 1) & 1ull) * B * ntraj;
        double* c_scratch = scratch + (ph & 1ull) * B * ntraj;
        const unsigned long long g_items = g_nrows * nch, c_items = c_nrows * nch;
        // (a fixed round-robin: handing the items out from an atomic counter, to fill the last round of a phase, cost 2.6 ms — one
        // more barrier and an atomic round trip per item)
        for (unsigned long long w = blockIdx.x; w < g_items + c_items; w += gridDim.x) {
            if (w < g_items) {   // stage 2 on the blocks that make up one chunk's worth of columns
                const unsigned long long q = w / nch, c = w % nch, lo = c * PERM_CH, hi = lo + PERM_CH < ntraj ? lo + PERM_CH : ntraj;
                const double* s = g_scratch + q * ntraj;
                double* d = rows + (g_row0 + q) * ntraj;
#pragma unroll 8
                for (unsigned long long m = lo + threadIdx.x; m < hi; m += PERM_NT)
                    perm_tile[((m - lo) & ~(unsigned long long)(PERM_BK - 1)) + d2[m]] = __ldcg(s + m);
                __syncthreads();
#pragma unroll 8
                for (unsigned long long j = lo + threadIdx.x; j < hi; j += PERM_NT) __stcs(d + j, perm_tile[j - lo]);
                // this part of the scratch is dead now (it is rewritten before it is read again): drop its lines from L2 instead of
                // letting them be written back to DRAM (ncu: 33 GB written for 17 GB of rows without this).  Only 128-byte lines
                // that lie wholly inside this item's range — no other CTA reads them.
                {
                    const unsigned long long a0 = ((unsigned long long)(s + lo) + 127ull) & ~127ull, a1 = (unsigned long long)(s + hi) & ~127ull;
                    for (unsigned long long a = a0 + (unsigned long long)threadIdx.x * 128ull; a < a1; a += (unsigned long long)PERM_NT * 128ull)
                        asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
                }
                __syncthreads();
            } else {
                const unsigned long long ww = w - g_items, q = ww / nch, c = ww % nch, lo = c * PERM_CH, hi = lo + PERM_CH < ntraj ? lo + PERM_CH : ntraj;
                const double* s = rows + (c_row0 + q) * ntraj;
                double* d = c_scratch + q * ntraj;
#pragma unroll 8
                for (unsigned long long j = lo + threadIdx.x; j < hi; j += PERM_NT) perm_tile[j - lo] = __ldcs(s + j);
                __syncthreads();
#pragma unroll 8
                for (unsigned long long i = lo + threadIdx.x; i < hi; i += PERM_NT) d[m1[i]] = perm_tile[l1[i]];
                __syncthreads();
            }
        }
        if (ph < nb) grid_barrier(barrier, gridDim.x, generation);
    }
}
// scratch: 2 * permute_rows_batch(ntraj) * ntraj + 3 doubles (the batches, then the barrier word and two item counters)
cudaError_t launch_permute_rows_staged(double* rows, const void* plan, unsigned long long nrows, unsigned long long ntraj, double* scratch,
                                       cudaStream_t st) {
    if (nrows == 0 || ntraj == 0) return cudaSuccess;
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = PERM_CH * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(permute_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, permute_staged_kernel, PERM_NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (per_sm > 2) per_sm = 2;   // measured: a third CTA per SM (192 KB of shared memory, little L1 left) halves stage 1's rate
    const char* q = (const char*)plan;
    const uint32_t* m1 = (const uint32_t*)q; q += align256(ntraj * 4);
    const uint16_t* l1 = (const uint16_t*)q; q += align256(ntraj * 2);
    const uint16_t* d2 = (const uint16_t*)q;
    const unsigned long long B = permute_rows_batch(ntraj), nch = (ntraj + PERM_CH - 1) / PERM_CH;
    unsigned* barrier = (unsigned*)(scratch + 2 * B * ntraj);   // barrier word, pad, two 64-bit item counters
    e = cudaMemsetAsync(barrier, 0, 3 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const unsigned long long cap = (unsigned long long)sms * (unsigned long long)per_sm, want = 2 * B * nch;
    permute_staged_kernel<<<(unsigned)(want < cap ? want : cap), PERM_NT, smem, st>>>(rows, m1, l1, d2, ntraj, nrows, B, scratch, barrier);
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
