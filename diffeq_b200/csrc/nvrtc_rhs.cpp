// nvrtc_rhs.cpp — placeholder until the NVRTC path lands: creation fails loudly with DEQ_ERR_NVRTC.
#include "nvrtc_rhs.h"
namespace deqn {
struct Program { int refs; };
Program* create(const char*, int, int, std::string* err) { if (err) *err = "NVRTC user right-hand sides are not built yet"; return nullptr; }
void retain(Program* p) { if (p) ++p->refs; }
void release(Program* p) { if (p && --p->refs == 0) delete p; }
bool launch_hinit(Program*, int, const deqk::HinitParams&, cudaStream_t, std::string* err) { if (err) *err = "no NVRTC"; return false; }
bool launch_adapt(Program*, int, int, int, const deqk::AdaptParams&, cudaStream_t, std::string* err) { if (err) *err = "no NVRTC"; return false; }
bool launch_fixed(Program*, int, int, int, const deqk::FixedParams&, cudaStream_t, std::string* err) { if (err) *err = "no NVRTC"; return false; }
}  // namespace deqn
