// Dev harness: instantiate only the headline kernels so SASS can be inspected quickly.
#include "../../diffeq_b200/csrc/stepper.cuh"
#ifndef HOT_FAST
#define HOT_FAST false
#endif
#ifndef HOT_TB
#define HOT_TB deqt::Dopri5<false>
#endif
#ifndef HOT_SYS
#define HOT_SYS deqr::Lorenz
#endif
#ifndef HOT_MODE
#define HOT_MODE deqk::MODE_FINAL
#endif
template __global__ void deqk::adapt_kernel<HOT_TB, HOT_SYS, false, HOT_MODE, HOT_FAST, false>(const deqk::AdaptParams);
#ifdef HOT_ROWS
template __global__ void deqk::adapt_rows_kernel<deqt::Rk45, deqr::VanDerPol, true, HOT_FAST>(const deqk::AdaptParams);
#endif
