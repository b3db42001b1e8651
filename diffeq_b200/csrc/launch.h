// launch.h — host-side entry points of the per-tableau kernel translation units (kernels_inst.cu is compiled once
// per tableau with -DDEQ_TB=<id>; see build.py).  capi.cu dispatches a (method, tableau set) pair to one of these.
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace deql {

// Tableau ids (one object file each).
enum TbId : int {
    TB_FEULER = 0, TB_MIDPOINT, TB_HEUN, TB_RK4,                 // fixed step
    TB_RK21_REF, TB_RK21_TXT, TB_RK23_REF, TB_RK23_TXT, TB_RK45,  // adaptive
    TB_DOPRI5_REF, TB_DOPRI5_TXT, TB_FEH78_REF, TB_FEH78_TXT,
    TB_COUNT
};

struct TableauInfo { int S; int adaptive; int order_min; int fsal; double a[13 * 13]; double b[13 * 2]; double c[13]; };

// Launch the adaptive kernel of tableau `tb` for built-in system `sys` (deq_system).  `grid_blocks` <= 0 lets the
// launcher size a persistent grid (SM count x resident CTAs per SM).  Returns the CUDA error of the launch.
typedef cudaError_t (*AdaptLaunchFn)(int sys, int per_traj_params, int mode, const deqk::AdaptParams& P,
                                     int grid_blocks, cudaStream_t st);
typedef cudaError_t (*FixedLaunchFn)(int sys, int per_traj_params, int all, const deqk::FixedParams& P,
                                     cudaStream_t st);
typedef void (*InfoFn)(TableauInfo* out);

struct TbEntry { AdaptLaunchFn adapt; AdaptLaunchFn adapt_fast; AdaptLaunchFn adapt_f32; FixedLaunchFn fixed; InfoFn info; };
const TbEntry& tb_entry(int tb);  // defined in capi.cu from the per-TU registrars below

#define DEQ_DECL_TB(ID) \
    cudaError_t launch_adapt_##ID(int, int, int, const deqk::AdaptParams&, int, cudaStream_t); \
    cudaError_t launch_adapt_fast_##ID(int, int, int, const deqk::AdaptParams&, int, cudaStream_t); \
    cudaError_t launch_adapt_f32_##ID(int, int, int, const deqk::AdaptParams&, int, cudaStream_t);  \
    cudaError_t launch_fixed_##ID(int, int, int, const deqk::FixedParams&, cudaStream_t);      \
    void info_##ID(TableauInfo*);
DEQ_DECL_TB(0) DEQ_DECL_TB(1) DEQ_DECL_TB(2) DEQ_DECL_TB(3) DEQ_DECL_TB(4) DEQ_DECL_TB(5) DEQ_DECL_TB(6)
DEQ_DECL_TB(7) DEQ_DECL_TB(8) DEQ_DECL_TB(9) DEQ_DECL_TB(10) DEQ_DECL_TB(11) DEQ_DECL_TB(12)
#undef DEQ_DECL_TB

// stiff solvers (kernels_stiff.cu): ode23s, and oderosenbrock with `which` = 0 kr4 (Ode4skr) / 1 s4 (Ode4ss)
cudaError_t launch_ode23s(int sys, int per_traj_params, int mode, const deqk::AdaptParams& P, cudaStream_t st);
cudaError_t launch_ode23s_fast(int sys, int per_traj_params, int mode, const deqk::AdaptParams& P, cudaStream_t st);   // -fmad=true object
cudaError_t launch_rosenbrock(int which, int sys, int per_traj_params, int all, const deqk::FixedParams& P, cudaStream_t st);
void rosenbrock_info(int which, double* gamma, double* a /*[16]*/, double* b /*[4]*/, double* c /*[16]*/);

// hinit pre-pass and small helpers live in kernels_common.cu
cudaError_t launch_hinit(int sys, int per_traj_params, const deqk::HinitParams& P, cudaStream_t st);
cudaError_t launch_transpose_to_soa(const double* d_aos, double* d_soa, unsigned long long ntraj, int n, cudaStream_t st);
cudaError_t launch_transpose_to_aos(const double* d_soa, double* d_aos, unsigned long long ntraj, int n, cudaStream_t st);
cudaError_t launch_sum_counts(const uint32_t* acc, const uint32_t* rej, const uint32_t* nfe, unsigned long long n,
                              unsigned long long* totals, cudaStream_t st);
cudaError_t launch_coherence(const double* h0, unsigned long long n, unsigned long long* d_out, cudaStream_t st);
// first-step ordering of an ensemble (deq_solve's sorted SoA dense path): order[s] = index of the trajectory with the s-th smallest
// |h0|, inv = its inverse, h0_sorted[s] = |h0[order[s]]|; scratch of sort_by_h0_scratch_bytes(n) bytes
size_t sort_by_h0_scratch_bytes(unsigned long long n);
cudaError_t launch_sort_by_h0(const double* h0, unsigned long long n, uint32_t* order, uint32_t* inv, double* h0_sorted,
                              void* scratch, size_t scratch_bytes, cudaStream_t st);
// rows[q][j] <- rows[q][inv[j]] in place for nrows rows of ntraj doubles (columns back from sorted to caller order), through a
// scratch of 2 * permute_rows_batch(ntraj) * ntraj + 1 doubles (the batches stay in L2; the last word is a barrier counter)
unsigned long long permute_rows_batch(unsigned long long ntraj);
cudaError_t launch_permute_rows_inplace(double* rows, const uint32_t* inv, unsigned long long nrows, unsigned long long ntraj, double* scratch,
                                        cudaStream_t st);
// the same permutation factored into two coalesced stages (kernels_common.cu): plan built once per solve from `order`
size_t permute_plan_bytes(unsigned long long n);
size_t permute_plan_scratch_bytes(unsigned long long n);
cudaError_t launch_permute_plan(const uint32_t* order, unsigned long long n, void* plan, void* scratch, cudaStream_t st);
cudaError_t launch_permute_rows_staged(double* rows, const void* plan, unsigned long long nrows, unsigned long long ntraj, double* scratch,
                                       cudaStream_t st);
cudaError_t launch_test_div(unsigned long long seed, unsigned long long count, unsigned long long* d_out, cudaStream_t st);
cudaError_t launch_test_pow(const double* x, const double* y, double* out, unsigned long long n, cudaStream_t st);
cudaError_t launch_test_log10(const double* x, double* out, unsigned long long n, cudaStream_t st);
cudaError_t launch_dfma_peak(double* sink, int iters, int blocks, cudaStream_t st);

}  // namespace deql
