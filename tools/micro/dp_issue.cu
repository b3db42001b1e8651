// Does an FP64 instruction occupy the SMSP issue port for 2 cycles?  Interleave K independent integer ops per DFMA.
// If time ~ 2*DP + INT (per warp instruction), FP64 blocks issue; if time ~ max(2*DP, DP+INT), it co-issues.
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void __launch_bounds__(256) k(double* out, int* iout, int iters, unsigned seed) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    unsigned x0 = threadIdx.x ^ seed, x1 = x0 * 3u + seed, x2 = x0 * 5u, x3 = x0 * 7u;
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fma_rn(a0, m, c);
            if (K >= 1) x0 = (x0 ^ seed) + 0x9e3779b9u;
            if (K >= 2) x1 = (x1 ^ seed) + 0x7f4a7c15u;
            a1 = __fma_rn(a1, m, c);
            if (K >= 3) x2 = (x2 ^ seed) + 0x85ebca6bu;
            if (K >= 4) x3 = (x3 ^ seed) + 0xc2b2ae35u;
            a2 = __fma_rn(a2, m, c);
            if (K >= 5) x0 = (x0 ^ seed) + 0x27d4eb2fu;
            if (K >= 6) x1 = (x1 ^ seed) + 0x165667b1u;
            a3 = __fma_rn(a3, m, c);
            if (K >= 7) x2 = (x2 ^ seed) + 0xd3a2646cu;
            if (K >= 8) x3 = (x3 ^ seed) + 0xfd7046c5u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (a0 + a1) + (a2 + a3);
    iout[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3;
}
template <int K>
void run() {
    double* out; int* iout; cudaMalloc(&out, 148 * 8 * 256 * 8); cudaMalloc(&iout, 148 * 8 * 256 * 4);
    int iters = 4000, blocks = 148 * 4;  // 32 warps/SM = 8 per SMSP
    k<K><<<blocks, 256>>>(out, iout, iters, 12345u);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<K><<<blocks, 256>>>(out, iout, iters, 12345u); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SMSP: 8 warps x iters x 8 x (4 DP + 2K int-ish [xor+add]) warp-instructions
    double cyc = ms * 1e-3 * 1.965e9;
    double dp = 8.0 * iters * 8 * 4, in = 8.0 * iters * 8 * 2 * K / 2.0 * 2;  // K ops each = LOP3 + IADD (maybe fused to 1-2 instrs)
    printf("K=%d int-ops per 4 DFMA: %.0f cycles/SMSP; DP warp-instr %.0f -> %.2f cyc per DFMA-group-of-4 (2*DP model: 8 + int)\n", K, cyc, dp,
           cyc / (8.0 * iters * 8));
}
int main() { run<0>(); run<2>(); run<4>(); run<6>(); run<8>(); return 0; }
