// nvrtc_rhs.h — user right-hand sides compiled at run time (NVRTC, sm_100a) and launched through the driver API.
// The user supplies CUDA source for `__device__ void rhs(double t, const double* y, double* dy, const double* p)`
// — the device-side replacement of the reference's closure `F: Fn(f64,&Y)->Y` (problem.rs:32-36); the library splices
// it into the same stepper templates the built-in systems use (stepper.cuh), so numerics are identical.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "params.h"

namespace deqn {

struct Program;  // one compiled module per (user source, n, nparams); ref-counted

Program* create(const char* cuda_src, int n, int nparams, std::string* err);
void retain(Program* p);
void release(Program* p);

bool launch_hinit(Program* p, int per_traj_params, const deqk::HinitParams& P, cudaStream_t st, std::string* err);
bool launch_adapt(Program* p, int tb, int per_traj_params, int mode, bool fast, const deqk::AdaptParams& P,
                  cudaStream_t st, std::string* err);
bool launch_fixed(Program* p, int tb, int per_traj_params, int all, const deqk::FixedParams& P, cudaStream_t st,
                  std::string* err);

// stiff solvers (stiff.cuh), n <= 3: ode23s with the MODE_* output kinds, oderosenbrock with which = 0 kr4 / 1 s4
bool launch_ode23s(Program* p, int per_traj_params, int mode, bool fast, const deqk::AdaptParams& P, cudaStream_t st, std::string* err);
bool launch_rosenbrock(Program* p, int which, int per_traj_params, int all, const deqk::FixedParams& P, cudaStream_t st,
                       std::string* err);

}  // namespace deqn
