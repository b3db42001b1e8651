// FP64 pipe microbenchmark for B200: dependent-chain latency and throughput vs ILP / warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/micro/dp_latency.cu -o /tmp/dp_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, int iters, long long* cyc) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = __fma_rn(a[i], m, c);
                else if (OP == 1) a[i] = __dadd_rn(a[i], c);
                else if (OP == 2) a[i] = __dmul_rn(a[i], m);
                else { a[i] = __dmul_rn(a[i], m); a[i] = __dadd_rn(a[i], c); }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP>
void run(int warps_per_sm, const char* name) {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&cyc, 8);
    int iters = 2000;
    int threads = warps_per_sm * 32; int blocks = 148;
    if (threads > 1024) { blocks = 148 * (threads / 1024); threads = 1024; }
    k<ILP, OP><<<blocks, threads>>>(out, iters, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<ILP, OP><<<blocks, threads>>>(out, iters, cyc); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops_per_thread = (double)iters * 16 * ILP * (OP == 3 ? 2 : 1);
    double total = ops_per_thread * blocks * threads;
    printf("%-8s ILP=%d warps/SM=%2d: %.2f cyc per dependent op (1 warp view), %.2f T-op/s, %.1f%% of 18.6 Tinstr-lane/s\n", name, ILP,
           warps_per_sm, (double)h / (iters * 16.0 * (OP == 3 ? 2 : 1)), total / (ms * 1e-3) / 1e12, 100.0 * total / (ms * 1e-3) / (148 * 64 * 1.965e9));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>(1, "DFMA"); run<1, 1>(1, "DADD"); run<1, 2>(1, "DMUL");
    run<1, 0>(4, "DFMA"); run<1, 0>(8, "DFMA"); run<1, 0>(16, "DFMA"); run<1, 0>(32, "DFMA"); run<1, 0>(64, "DFMA");
    run<2, 0>(16, "DFMA"); run<3, 0>(16, "DFMA"); run<4, 0>(16, "DFMA"); run<8, 0>(16, "DFMA");
    run<3, 3>(16, "MUL+ADD"); run<3, 3>(24, "MUL+ADD"); run<3, 3>(32, "MUL+ADD"); run<6, 3>(16, "MUL+ADD");
    return 0;
}
