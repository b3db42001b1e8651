// perm_bench.cu — where does the column permutation of the sorted dense path spend its time?  (development probe)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/perm_bench.cu -o tools/micro/perm_bench ; ./perm_bench [log2_ntraj] [nrows]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef unsigned long long u64;

__global__ void __launch_bounds__(256) copy_k(const double* __restrict__ s, double* __restrict__ d, u64 n) {
    for (u64 i = ((u64)blockIdx.x * 256 + threadIdx.x) * 2; i < n; i += (u64)gridDim.x * 512) {
        const double2 v = *(const double2*)(s + i);
        *(double2*)(d + i) = v;
    }
}
// out-of-place gather, rows in order
__global__ void __launch_bounds__(256) gather_k(const double* __restrict__ src, double* __restrict__ dst, const unsigned* __restrict__ inv, u64 nrows, u64 ntraj, u64 chunks) {
    for (u64 b = blockIdx.x; b < nrows * chunks; b += gridDim.x) {
        const u64 row = b / chunks, j0 = (b % chunks) * 1024ull;
        const double* s = src + row * ntraj; double* d = dst + row * ntraj;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u64 j = j0 + u * 256 + threadIdx.x; if (j < ntraj) v[u] = s[inv[j]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u64 j = j0 + u * 256 + threadIdx.x; if (j < ntraj) d[j] = v[u]; }
    }
}
// out-of-place scatter
__global__ void __launch_bounds__(256) scatter_k(const double* __restrict__ src, double* __restrict__ dst, const unsigned* __restrict__ order, u64 nrows, u64 ntraj, u64 chunks) {
    for (u64 b = blockIdx.x; b < nrows * chunks; b += gridDim.x) {
        const u64 row = b / chunks, j0 = (b % chunks) * 1024ull;
        const double* s = src + row * ntraj; double* d = dst + row * ntraj;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u64 j = j0 + u * 256 + threadIdx.x; if (j < ntraj) v[u] = s[j]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u64 j = j0 + u * 256 + threadIdx.x; if (j < ntraj) d[order[j]] = v[u]; }
    }
}
// in-place through a per-CTA-group... variant: row-at-a-time with shared memory halves is impossible (row = 8 MB); instead
// gather where the source row is first pulled into L2 by a coalesced prefetch pass of the SAME kernel one row ahead
__global__ void __launch_bounds__(256) gather_pf_k(const double* __restrict__ src, double* __restrict__ dst, const unsigned* __restrict__ inv, u64 nrows, u64 ntraj, u64 chunks, int ahead) {
    for (u64 b = blockIdx.x; b < nrows * chunks; b += gridDim.x) {
        const u64 row = b / chunks, j0 = (b % chunks) * 1024ull;
        if (row + ahead < nrows && threadIdx.x < 64) {   // 1024 doubles = 64 lines of 128 bytes
            const double* pf = src + (row + ahead) * ntraj + j0 + threadIdx.x * 16;
            if (j0 + threadIdx.x * 16 < ntraj) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        }
        const double* s = src + row * ntraj; double* d = dst + row * ntraj;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u64 j = j0 + u * 256 + threadIdx.x; if (j < ntraj) v[u] = __ldcg(s + inv[j]); }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u64 j = j0 + u * 256 + threadIdx.x; if (j < ntraj) __stcs(d + j, v[u]); }
    }
}

// ---- two coalesced stages (plan built on the host here) ----
#ifndef PERM_BK
#define PERM_BK 1024
#endif
constexpr int CH = 8192, BK = PERM_BK;
// stage 1: chunk c of 8192 consecutive slots -> intermediate positions m1[i] (ascending within the chunk), value from slot l1[i] of the chunk
__global__ void __launch_bounds__(256) stage1_k(const double* __restrict__ src, double* __restrict__ mid, const unsigned* __restrict__ m1,
                                                const unsigned short* __restrict__ l1, u64 nrows, u64 ntraj, u64 nch) {
    extern __shared__ double tile[];
    for (u64 w = blockIdx.x; w < nrows * nch; w += gridDim.x) {
        const u64 q = w / nch, c = w % nch, lo = c * CH, hi = lo + CH < ntraj ? lo + CH : ntraj;
        const double* s = src + q * ntraj; double* d = mid + q * ntraj;
#pragma unroll 8
        for (u64 j = lo + threadIdx.x; j < hi; j += 256) tile[j - lo] = __ldcs(s + j);
        __syncthreads();
#pragma unroll 8
        for (u64 i = lo + threadIdx.x; i < hi; i += 256) d[m1[i]] = tile[l1[i]];
        __syncthreads();
    }
}
// stage 2: inside each BK-block, element m goes to column BK b + d2[m]
__global__ void __launch_bounds__(256) stage2_k(const double* __restrict__ mid, double* __restrict__ dst, const unsigned short* __restrict__ d2,
                                                u64 nrows, u64 ntraj, u64 nbk) {
    extern __shared__ double tile[];
    for (u64 w = blockIdx.x; w < nrows * nbk; w += gridDim.x) {
        const u64 q = w / nbk, b = w % nbk, lo = b * BK, hi = lo + BK < ntraj ? lo + BK : ntraj;
        const double* s = mid + q * ntraj; double* d = dst + q * ntraj;
#pragma unroll 4
        for (u64 m = lo + threadIdx.x; m < hi; m += 256) tile[d2[m]] = __ldcg(s + m);
        __syncthreads();
#pragma unroll 4
        for (u64 j = lo + threadIdx.x; j < hi; j += 256) __stcs(d + j, tile[j - lo]);
        __syncthreads();
    }
}
// ---- in place: persistent kernel, phase ph = stage 2 of batch ph - 1 (scratch -> rows) + stage 1 of batch ph (rows -> other scratch half) ----
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
template <int NT>
__global__ void __launch_bounds__(NT) perm2_persistent_k(double* rows, const unsigned* __restrict__ m1, const unsigned short* __restrict__ l1,
                                                         const unsigned short* __restrict__ d2, u64 ntraj, u64 nrows, u64 B, double* scratch, unsigned* barrier) {
    extern __shared__ double tile[];
    const u64 nb = (nrows + B - 1) / B, nch = (ntraj + CH - 1) / CH;
    unsigned generation = 0;
    for (u64 ph = 0; ph <= nb; ++ph) {
        const u64 g_row0 = ph ? (ph - 1) * B : 0, g_nrows = ph ? (ph < nb ? B : nrows - (nb - 1) * B) : 0;
        const u64 c_row0 = ph * B, c_nrows = ph < nb ? (ph + 1 < nb ? B : nrows - ph * B) : 0;
        const double* g_scratch = scratch + ((ph + 1) & 1ull) * B * ntraj;
        double* c_scratch = scratch + (ph & 1ull) * B * ntraj;
        const u64 g_items = g_nrows * nch, c_items = c_nrows * nch;
        for (u64 w = blockIdx.x; w < g_items + c_items; w += gridDim.x) {
            if (w < g_items) {   // stage 2 on a group of 8 blocks (one chunk's worth of columns)
                const u64 q = w / nch, c = w % nch, lo = c * CH, hi = lo + CH < ntraj ? lo + CH : ntraj;
                const double* s = g_scratch + q * ntraj; double* d = rows + (g_row0 + q) * ntraj;
#pragma unroll 8
                for (u64 m = lo + threadIdx.x; m < hi; m += NT) tile[((m - lo) & ~(u64)(BK - 1)) + d2[m]] = __ldcg(s + m);
                __syncthreads();
#pragma unroll 8
                for (u64 j = lo + threadIdx.x; j < hi; j += NT) __stcs(d + j, tile[j - lo]);
                __syncthreads();
            } else {
                const u64 ww = w - g_items, q = ww / nch, c = ww % nch, lo = c * CH, hi = lo + CH < ntraj ? lo + CH : ntraj;
                const double* s = rows + (c_row0 + q) * ntraj; double* d = c_scratch + q * ntraj;
#pragma unroll 8
                for (u64 j = lo + threadIdx.x; j < hi; j += NT) tile[j - lo] = __ldcs(s + j);
                __syncthreads();
#pragma unroll 8
                for (u64 i = lo + threadIdx.x; i < hi; i += NT) d[m1[i]] = tile[l1[i]];
                __syncthreads();
            }
        }
        if (ph < nb) grid_barrier(barrier, gridDim.x, generation);
    }
}

// variant: chunk size CHX (plan arrays m1x / l1x built for it), stage 1 either staged in shared memory or read straight from the (L1-resident) chunk
template <int NT, int CHX, bool DIRECT>
__global__ void __launch_bounds__(NT) perm2x_k(double* rows, const unsigned* __restrict__ m1, const unsigned short* __restrict__ l1,
                                               const unsigned short* __restrict__ d2, u64 ntraj, u64 nrows, u64 B, double* scratch, unsigned* barrier) {
    extern __shared__ double tile[];
    const u64 nb = (nrows + B - 1) / B, nch = (ntraj + CHX - 1) / CHX;
    unsigned generation = 0;
    for (u64 ph = 0; ph <= nb; ++ph) {
        const u64 g_row0 = ph ? (ph - 1) * B : 0, g_nrows = ph ? (ph < nb ? B : nrows - (nb - 1) * B) : 0;
        const u64 c_row0 = ph * B, c_nrows = ph < nb ? (ph + 1 < nb ? B : nrows - ph * B) : 0;
        const double* g_scratch = scratch + ((ph + 1) & 1ull) * B * ntraj;
        double* c_scratch = scratch + (ph & 1ull) * B * ntraj;
        const u64 g_items = g_nrows * nch, c_items = c_nrows * nch;
        for (u64 w = blockIdx.x; w < g_items + c_items; w += gridDim.x) {
            if (w < g_items) {
                const u64 q = w / nch, c = w % nch, lo = c * CHX, hi = lo + CHX < ntraj ? lo + CHX : ntraj;
                const double* s = g_scratch + q * ntraj; double* d = rows + (g_row0 + q) * ntraj;
#pragma unroll 4
                for (u64 m = lo + threadIdx.x; m < hi; m += NT) tile[((m - lo) & ~(u64)(BK - 1)) + d2[m]] = __ldcg(s + m);
                __syncthreads();
#pragma unroll 4
                for (u64 j = lo + threadIdx.x; j < hi; j += NT) __stcs(d + j, tile[j - lo]);
                __syncthreads();
            } else {
                const u64 ww = w - g_items, q = ww / nch, c = ww % nch, lo = c * CHX, hi = lo + CHX < ntraj ? lo + CHX : ntraj;
                const double* s = rows + (c_row0 + q) * ntraj; double* d = c_scratch + q * ntraj;
                if (DIRECT) {
#pragma unroll 4
                    for (u64 i = lo + threadIdx.x; i < hi; i += NT) d[m1[i]] = s[lo + l1[i]];
                } else {
#pragma unroll 4
                    for (u64 j = lo + threadIdx.x; j < hi; j += NT) tile[j - lo] = __ldcs(s + j);
                    __syncthreads();
#pragma unroll 4
                    for (u64 i = lo + threadIdx.x; i < hi; i += NT) d[m1[i]] = tile[l1[i]];
                    __syncthreads();
                }
            }
        }
        if (ph < nb) grid_barrier(barrier, gridDim.x, generation);
    }
}
int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 20;
    const u64 ntraj = 1ull << lg, nrows = argc > 2 ? atoll(argv[2]) : 1000, n = ntraj * nrows, chunks = (ntraj + 1023) / 1024;
    double *a, *b; unsigned *inv, *order;
    CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&inv, ntraj * 4)); CK(cudaMalloc(&order, ntraj * 4));
    CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8));
    std::vector<unsigned> h(ntraj), hi(ntraj);
    for (u64 i = 0; i < ntraj; ++i) h[i] = (unsigned)i;
    std::mt19937_64 g(1); std::shuffle(h.begin(), h.end(), g);
    for (u64 i = 0; i < ntraj; ++i) hi[h[i]] = (unsigned)i;
    CK(cudaMemcpy(order, h.data(), ntraj * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(inv, hi.data(), ntraj * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn) {
        fn(); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); fn(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %8.3f ms  %7.1f GB/s (read + write of the rows)\n", name, ms, 2.0 * n * 8 / ms / 1e6);
    };
    const unsigned grid = 148 * 8;
    timeit("copy", [&] { copy_k<<<grid, 256>>>(a, b, n); });
    timeit("gather out-of-place", [&] { gather_k<<<grid, 256>>>(a, b, inv, nrows, ntraj, chunks); });
    timeit("scatter out-of-place", [&] { scatter_k<<<grid, 256>>>(a, b, order, nrows, ntraj, chunks); });
    for (int ahead : {1, 2, 4}) { char nm[64]; snprintf(nm, 64, "gather + L2 prefetch +%d", ahead); timeit(nm, [&] { gather_pf_k<<<grid, 256>>>(a, b, inv, nrows, ntraj, chunks, ahead); }); }
    for (unsigned gm : {2u, 4u, 16u}) { char nm[64]; snprintf(nm, 64, "gather pf+2, %u CTAs/SM", gm); timeit(nm, [&] { gather_pf_k<<<148 * gm, 256>>>(a, b, inv, nrows, ntraj, chunks, 2); }); }
    // plan: S1 = slots stably sorted by order[s] / 1024; M1 = intermediate positions stably sorted by S1[m] / 8192
    {
        std::vector<unsigned> S1(ntraj), M1(ntraj); std::vector<unsigned short> L1(ntraj), D2(ntraj);
        for (u64 i = 0; i < ntraj; ++i) S1[i] = (unsigned)i;
        std::stable_sort(S1.begin(), S1.end(), [&](unsigned x, unsigned y) { return h[x] / BK < h[y] / BK; });
        for (u64 i = 0; i < ntraj; ++i) M1[i] = (unsigned)i;
        std::stable_sort(M1.begin(), M1.end(), [&](unsigned x, unsigned y) { return S1[x] / CH < S1[y] / CH; });
        for (u64 i = 0; i < ntraj; ++i) { L1[i] = (unsigned short)(S1[M1[i]] % CH); D2[i] = (unsigned short)(h[S1[i]] % BK); }
        unsigned* m1; unsigned short *l1, *d2; double* mid;
        CK(cudaMalloc(&m1, ntraj * 4)); CK(cudaMalloc(&l1, ntraj * 2)); CK(cudaMalloc(&d2, ntraj * 2)); CK(cudaMalloc(&mid, n * 8));
        CK(cudaMemcpy(m1, M1.data(), ntraj * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(l1, L1.data(), ntraj * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d2, D2.data(), ntraj * 2, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(stage1_k, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * 8));
        CK(cudaFuncSetAttribute(stage2_k, cudaFuncAttributeMaxDynamicSharedMemorySize, BK * 8));
        const u64 nch = (ntraj + CH - 1) / CH, nbk = (ntraj + BK - 1) / BK;
        // correctness: fill a with column ids, permute, compare with the scatter
        {
            std::vector<double> col(ntraj); for (u64 j = 0; j < ntraj; ++j) col[j] = (double)j;
            CK(cudaMemcpy(a, col.data(), ntraj * 8, cudaMemcpyHostToDevice));
            stage1_k<<<148 * 3, 256, CH * 8>>>(a, mid, m1, l1, 1, ntraj, nch); stage2_k<<<148 * 2, 256, BK * 8>>>(mid, b, d2, 1, ntraj, nbk);
            std::vector<double> out(ntraj); CK(cudaMemcpy(out.data(), b, ntraj * 8, cudaMemcpyDeviceToHost));
            u64 bad = 0; for (u64 sidx = 0; sidx < ntraj; ++sidx) if (out[h[sidx]] != (double)sidx) ++bad;
            printf("two-stage permutation check: %llu wrong of %llu\n", bad, ntraj);
        }
        // in place, persistent: batch sizes and CTA shapes
        {
            double* scratch; unsigned* bar; CK(cudaMalloc(&scratch, 2ull * 16 * ntraj * 8)); CK(cudaMalloc(&bar, 4));
            CK(cudaFuncSetAttribute(perm2_persistent_k<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * 8));
            CK(cudaFuncSetAttribute(perm2_persistent_k<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * 8));
            {   // correctness of the in-place form on 7 rows, B = 3
                std::vector<double> col(7 * ntraj); for (u64 q = 0; q < 7; ++q) for (u64 j = 0; j < ntraj; ++j) col[q * ntraj + j] = (double)(j + q);
                CK(cudaMemcpy(a, col.data(), 7 * ntraj * 8, cudaMemcpyHostToDevice)); CK(cudaMemset(bar, 0, 4));
                perm2_persistent_k<256><<<148 * 2, 256, CH * 8>>>(a, m1, l1, d2, ntraj, 7, 3, scratch, bar);
                std::vector<double> out(7 * ntraj); CK(cudaMemcpy(out.data(), a, 7 * ntraj * 8, cudaMemcpyDeviceToHost));
                u64 bad = 0; for (u64 q = 0; q < 7; ++q) for (u64 sidx = 0; sidx < ntraj; ++sidx) if (out[q * ntraj + h[sidx]] != (double)(sidx + q)) ++bad;
                printf("in-place persistent two-stage check: %llu wrong\n", bad);
            }
            {   // chunk 4096 plan
                constexpr int CX = 4096;
                std::vector<unsigned> M1x(ntraj); std::vector<unsigned short> L1x(ntraj);
                for (u64 i = 0; i < ntraj; ++i) M1x[i] = (unsigned)i;
                std::stable_sort(M1x.begin(), M1x.end(), [&](unsigned x, unsigned y) { return S1[x] / CX < S1[y] / CX; });
                for (u64 i = 0; i < ntraj; ++i) L1x[i] = (unsigned short)(S1[M1x[i]] % CX);
                unsigned* m1x; unsigned short* l1x; CK(cudaMalloc(&m1x, ntraj * 4)); CK(cudaMalloc(&l1x, ntraj * 2));
                CK(cudaMemcpy(m1x, M1x.data(), ntraj * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(l1x, L1x.data(), ntraj * 2, cudaMemcpyHostToDevice));
                auto run = [&](const char* nm, auto kfn, unsigned grid, unsigned nt, size_t sm, const unsigned* pm1, const unsigned short* pl1, u64 B) {
                    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                    int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, (int)nt, sm));
                    if (grid > 148u * per_sm) { printf("%s: %u CTAs/SM asked, %d fit\n", nm, grid / 148, per_sm); grid = 148u * per_sm; }
                    char name[96]; snprintf(name, 96, "%s B=%llu grid=%ux%u", nm, B, grid / 148, nt);
                    timeit(name, [&] { cudaMemsetAsync(bar, 0, 4); kfn<<<grid, nt, sm>>>(a, pm1, pl1, d2, ntraj, nrows, B, scratch, bar); });
                };
                for (u64 B : {3ull, 6ull}) {
                    run("CH4096 smem", perm2x_k<256, 4096, false>, 148 * 4, 256, 4096 * 8, m1x, l1x, B);
                    run("CH4096 smem", perm2x_k<256, 4096, false>, 148 * 6, 256, 4096 * 8, m1x, l1x, B);
                    run("CH4096 smem", perm2x_k<512, 4096, false>, 148 * 4, 512, 4096 * 8, m1x, l1x, B);
                    run("CH4096 direct", perm2x_k<256, 4096, true>, 148 * 6, 256, 4096 * 8, m1x, l1x, B);
                    run("CH8192 direct", perm2x_k<512, 8192, true>, 148 * 2, 512, 8192 * 8, m1, l1, B);
                    run("CH8192 smem", perm2x_k<1024, 8192, false>, 148 * 2, 1024, 8192 * 8, m1, l1, B);
                }
            }
            for (u64 B : {6ull})
                for (int shape = 0; shape < 3; ++shape) {
                    char nm[64]; snprintf(nm, 64, "in-place 2-stage B=%llu %s", B, shape == 0 ? "2x256" : shape == 1 ? "2x512" : "1x512");
                    timeit(nm, [&] {
                        cudaMemsetAsync(bar, 0, 4);
                        if (shape == 0) perm2_persistent_k<256><<<148 * 2, 256, CH * 8>>>(a, m1, l1, d2, ntraj, nrows, B, scratch, bar);
                        else if (shape == 1) perm2_persistent_k<512><<<148 * 2, 512, CH * 8>>>(a, m1, l1, d2, ntraj, nrows, B, scratch, bar);
                        else perm2_persistent_k<512><<<148, 512, CH * 8>>>(a, m1, l1, d2, ntraj, nrows, B, scratch, bar);
                    });
                }
        }
        for (unsigned g1 : {1u, 2u, 3u}) { char nm[64]; snprintf(nm, 64, "stage 1 (%u CTAs/SM)", g1); timeit(nm, [&] { stage1_k<<<148 * g1, 256, CH * 8>>>(a, mid, m1, l1, nrows, ntraj, nch); }); }
        for (unsigned g2 : {2u, 3u}) { char nm[64]; snprintf(nm, 64, "stage 2 (%u CTAs/SM)", g2); timeit(nm, [&] { stage2_k<<<148 * g2, 256, BK * 8>>>(mid, b, d2, nrows, ntraj, nbk); }); }
    }
    return 0;
}
