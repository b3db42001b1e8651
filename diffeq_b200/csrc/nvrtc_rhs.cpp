// nvrtc_rhs.cpp — user right-hand sides: CUDA source -> NVRTC (sm_100a, -fmad=false) -> cubin -> cudaLibrary kernel.
//
// The reference's RHS is a Rust closure `F: Fn(f64,&Y)->Y` (problem.rs:32-36), which cannot run on the device; the
// replacement contract is CUDA source defining
//     __device__ void rhs(double t, const double* y, double* dy, const double* p);
// It is spliced in front of the very same stepper templates the built-in systems are compiled from (stepper.cuh and
// its headers are embedded in the library as strings, nvrtc_sources.inc), so a user system gets the same parity
// arithmetic, persistent scheduling and output modes.  Kernels are compiled lazily, one (kind, tableau, params
// layout, output mode) at a time, and cached per Program.  libnvrtc is dlopen()ed on first use so that the library
// itself only depends on the (statically linked) CUDA runtime.
#include "nvrtc_rhs.h"

#include <dlfcn.h>

#include <atomic>
#include <map>
#include <mutex>
#include <sstream>
#include <tuple>
#include <vector>

namespace {

#include "nvrtc_sources.inc"  // static const char* const kSrcNames[]; kSrcTexts[]; kNumSrc

// ---- minimal NVRTC API surface, resolved at run time ----
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
    const char* (*GetErrorString)(nvrtcResult);
};

Nvrtc* nvrtc(std::string* err) {
    static Nvrtc api;
    static std::once_flag once;
    static std::string load_err;
    std::call_once(once, [] {
        const char* cands[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        for (const char* c : cands) {
            api.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (api.h) break;
        }
        if (!api.h) { load_err = std::string("cannot load libnvrtc: ") + dlerror(); return; }
#define L(name) *(void**)(&api.name) = dlsym(api.h, "nvrtc" #name); if (!api.name) { load_err = "libnvrtc lacks nvrtc" #name; api.h = nullptr; return; }
        L(CreateProgram) L(DestroyProgram) L(CompileProgram) L(GetCUBINSize) L(GetCUBIN) L(GetProgramLogSize)
        L(GetProgramLog) L(AddNameExpression) L(GetLoweredName) L(GetErrorString)
#undef L
    });
    if (!api.h) { if (err) *err = load_err; return nullptr; }
    return &api;
}

const char* tb_type(int tb) {
    static const char* names[] = {"deqt::Feuler", "deqt::Midpoint", "deqt::Heun", "deqt::Rk4", "deqt::Rk21<false>",
                                  "deqt::Rk21<true>", "deqt::Rk23<false>", "deqt::Rk23<true>", "deqt::Rk45",
                                  "deqt::Dopri5<false>", "deqt::Dopri5<true>", "deqt::Feh78<false>", "deqt::Feh78<true>"};
    return (tb >= 0 && tb < 13) ? names[tb] : nullptr;
}

struct Kernel { cudaLibrary_t lib = nullptr; cudaKernel_t fn = nullptr; int blocks_per_sm = 0; };

// Process-wide cache of compiled kernels keyed by (user source, n, nparams, kernel instantiation): creating the same
// right-hand side again — e.g. one deq_rhs per worker thread — does not recompile.  Entries live until process exit.
std::mutex g_cache_mu;
std::map<std::string, Kernel> g_cache;

}  // namespace

namespace deqn {

struct Program {
    std::atomic<int> refs{1};
    std::string user_src;
    int n = 0, np = 0;
    std::mutex mu;
    std::map<std::tuple<int, int, int, int>, Kernel> cache;  // (kind, tb, pt, mode)
    // compiled libraries are owned by the process-wide cache, not by the Program
};

static bool compile(Program* p, const std::string& name_expr, bool fmad, Kernel* out, std::string* err, bool stiff = false) {
    Nvrtc* rt = nvrtc(err);
    if (!rt) return false;
    std::ostringstream src;
    src << "#define DEQ_USER_N " << p->n << "\n#define DEQ_USER_NP " << p->np << "\n"
        << "__device__ void rhs(double t, const double* y, double* dy, const double* p);\n"
        << "#include \"rhs.cuh\"\n"
        << "#line 1 \"user_rhs.cu\"\n" << p->user_src << "\n#line 1 \"deq_user_glue.cu\"\n"
        << (stiff ? "#include \"stiff.cuh\"\n" : "#include \"stepper.cuh\"\n")
        << "struct DeqUserRhs {\n  static constexpr int N = DEQ_USER_N, NP = DEQ_USER_NP;\n"
        << "  static constexpr bool USES_SINCOS = true;  // user code may call deq_sin / deq_cos (3.5 KB of shared memory)\n"
        << "  __device__ __forceinline__ static void eval(double t, const double* y, double* dy, const double* p) { rhs(t, y, dy, p); }\n"
        // optional: time-only terms evaluated once per distinct stage time (stepper.cuh `stages`).  The user source defines
        // DEQ_USER_NG and two more functions, and promises rhs(t, y, dy, p) == rhs_g(t, y, dy, p, g) with g = forcing(t, ., p).
        << "#ifdef DEQ_USER_NG\n  static constexpr int NG = DEQ_USER_NG;\n"
        << "  __device__ __forceinline__ static void forcing(double t, double* g, const double* p) { ::forcing(t, g, p); }\n"
        << "  __device__ __forceinline__ static void eval_g(double t, const double* y, double* dy, const double* p, const double* g) { rhs_g(t, y, dy, p, g); }\n"
        << "#endif\n};\n";
    nvrtcProgram prog = nullptr;
    nvrtcResult rc = rt->CreateProgram(&prog, src.str().c_str(), "deq_user_program.cu", kNumSrc, kSrcTexts, kSrcNames);
    if (rc) { if (err) *err = std::string("nvrtcCreateProgram: ") + rt->GetErrorString(rc); return false; }
    rt->AddNameExpression(prog, name_expr.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", fmad ? "--fmad=true" : "--fmad=false", "-default-device",
                          "-lineinfo", "-DDEQ_NVRTC=1", fmad ? "-DDEQ_FAST_TU=1" : "-DDEQ_PARITY_TU=1"};
    rc = rt->CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
    if (rc) {
        size_t n = 0;
        rt->GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) rt->GetProgramLog(prog, &log[0]);
        if (err) *err = std::string("NVRTC compilation failed: ") + rt->GetErrorString(rc) + "\n" + log;
        rt->DestroyProgram(&prog);
        return false;
    }
    const char* lowered = nullptr;
    rc = rt->GetLoweredName(prog, name_expr.c_str(), &lowered);
    if (rc || !lowered) { if (err) *err = "nvrtcGetLoweredName failed for " + name_expr; rt->DestroyProgram(&prog); return false; }
    const std::string lname = lowered;
    size_t sz = 0;
    rt->GetCUBINSize(prog, &sz);
    std::vector<char> cubin(sz);
    rt->GetCUBIN(prog, cubin.data());
    rt->DestroyProgram(&prog);
    cudaError_t ce = cudaLibraryLoadData(&out->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce != cudaSuccess) { if (err) *err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce); return false; }
    ce = cudaLibraryGetKernel(&out->fn, out->lib, lname.c_str());
    if (ce != cudaSuccess) { if (err) *err = std::string("cudaLibraryGetKernel(") + lname + "): " + cudaGetErrorString(ce); return false; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)out->fn, deqk::BLOCK, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
    }
    out->blocks_per_sm = per_sm;
    return true;
}

static bool get_kernel(Program* p, int kind, int tb, int pt, int mode, Kernel* out, std::string* err) {
    std::lock_guard<std::mutex> g(p->mu);
    auto key = std::make_tuple(kind, tb, pt, mode);
    auto it = p->cache.find(key);
    if (it != p->cache.end()) { *out = it->second; return true; }
    std::ostringstream ne;
    const char* b = pt ? "true" : "false";
    if (kind == 0) ne << "&deqk::hinit_kernel<DeqUserRhs, " << b << ">";
    else if (kind == 1) ne << "&deqk::adapt_kernel<" << tb_type(tb) << ", DeqUserRhs, " << b << ", " << (mode & 3) << ", "
                           << ((mode & 4) ? "true" : "false") << ", " << ((mode & 8) ? "true" : "false") << ">";
    else if (kind == 2) ne << "&deqk::fixed_kernel<" << tb_type(tb) << ", DeqUserRhs, " << b << ", " << (mode ? "true" : "false") << ">";
    else if (kind == 5) ne << "&deqk::adapt_rows_kernel<" << tb_type(tb) << ", DeqUserRhs, " << b << ", " << ((mode & 4) ? "true" : "false") << ">";
    else if (kind == 3) ne << "&deqk::ode23s_kernel<DeqUserRhs, " << b << ", " << (mode & 3) << ", " << ((mode & 4) ? "true" : "false") << ">";
    else ne << "&deqk::rosenbrock_kernel<" << (tb == 0 ? "deqk::Kr4" : "deqk::S4") << ", DeqUserRhs, " << b << ", " << (mode ? "true" : "false") << ">";
    Kernel k;
    const bool fmad = (kind == 1 || kind == 3 || kind == 5) && (mode & 4);
    const std::string gkey = std::to_string(p->n) + "|" + std::to_string(p->np) + "|" + ne.str() + (fmad ? "|fast|" : "|parity|") + p->user_src;
    {
        std::lock_guard<std::mutex> gg(g_cache_mu);
        auto git = g_cache.find(gkey);
        if (git != g_cache.end()) { p->cache[key] = git->second; *out = git->second; return true; }
    }
    if (!compile(p, ne.str(), fmad, &k, err, kind == 3 || kind == 4)) return false;
    {
        std::lock_guard<std::mutex> gg(g_cache_mu);
        g_cache[gkey] = k;
    }
    p->cache[key] = k;
    *out = k;
    return true;
}

Program* create(const char* cuda_src, int n, int nparams, std::string* err) {
    Program* p = new Program();
    p->user_src = cuda_src;
    p->n = n;
    p->np = nparams;
    // compile the hinit pre-pass right away so that syntax errors surface at creation, like a failing `fun` would
    Kernel k;
    if (!get_kernel(p, 0, 0, 0, 0, &k, err)) { delete p; return nullptr; }
    return p;
}
void retain(Program* p) { if (p) p->refs.fetch_add(1); }
void release(Program* p) { if (p && p->refs.fetch_sub(1) == 1) delete p; }

template <class P>
static bool launch(const Kernel& k, unsigned grid, const P& params, cudaStream_t st, std::string* err, size_t smem = 0) {
    void* args[] = {(void*)&params};
    cudaError_t ce = cudaLaunchKernel((const void*)k.fn, dim3(grid), dim3(deqk::BLOCK), args, smem, st);
    if (ce != cudaSuccess) { if (err) *err = std::string("cudaLaunchKernel: ") + cudaGetErrorString(ce); return false; }
    return true;
}

bool launch_hinit(Program* p, int pt, const deqk::HinitParams& P, cudaStream_t st, std::string* err) {
    Kernel k;
    if (!get_kernel(p, 0, 0, pt ? 1 : 0, 0, &k, err)) return false;
    return launch(k, (unsigned)((P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK), P, st, err);
}

static unsigned persistent_blocks(const Kernel& k, unsigned long long ntraj);

bool launch_adapt(Program* p, int tb, int pt, int mode, bool fast, const deqk::AdaptParams& P, cudaStream_t st,
                  std::string* err) {
    Kernel k;
    if (mode == deqk::MODE_DENSE_ROWS) {   // row-synchronous SoA dense kernel
        if (!get_kernel(p, 5, tb, pt ? 1 : 0, fast ? 4 : 0, &k, err)) return false;
        return launch(k, persistent_blocks(k, P.ntraj), P, st, err);
    }
    const bool spec = mode == deqk::MODE_REC_SPEC;   // Points::Specified flavour of MODE_REC
    if (spec) mode = deqk::MODE_REC;
    if (!get_kernel(p, 1, tb, pt ? 1 : 0, (mode & 3) | (fast ? 4 : 0) | (spec ? 8 : 0), &k, err)) return false;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (mode == deqk::MODE_DENSE_TM || (mode == deqk::MODE_REC && P.rec_mode == deqk::REC_WRITE)) ? 32u * deqk::BLOCK * sizeof(double) : 0;
    int per_sm = k.blocks_per_sm;
    if (smem) {  // the staged layout lowers the resident-CTA count
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k.fn, deqk::BLOCK, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 2; }
    }
    unsigned long long want = (P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK, cap = (unsigned long long)sms * per_sm;
    return launch(k, (unsigned)(want < cap ? (want ? want : 1) : cap), P, st, err, smem);
}

bool launch_fixed(Program* p, int tb, int pt, int all, const deqk::FixedParams& P, cudaStream_t st, std::string* err) {
    Kernel k;
    if (!get_kernel(p, 2, tb, pt ? 1 : 0, all ? 1 : 0, &k, err)) return false;
    return launch(k, (unsigned)((P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK), P, st, err);
}

static unsigned persistent_blocks(const Kernel& k, unsigned long long ntraj) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long want = (ntraj + deqk::BLOCK - 1) / deqk::BLOCK, cap = (unsigned long long)sms * k.blocks_per_sm;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

bool launch_ode23s(Program* p, int pt, int mode, bool fast, const deqk::AdaptParams& P, cudaStream_t st, std::string* err) {
    Kernel k;
    if (!get_kernel(p, 3, 0, pt ? 1 : 0, (mode & 3) | (fast ? 4 : 0), &k, err)) return false;
    return launch(k, persistent_blocks(k, P.ntraj), P, st, err);
}

bool launch_rosenbrock(Program* p, int which, int pt, int all, const deqk::FixedParams& P, cudaStream_t st, std::string* err) {
    Kernel k;
    if (!get_kernel(p, 4, which, pt ? 1 : 0, all ? 1 : 0, &k, err)) return false;
    return launch(k, (unsigned)((P.ntraj + deqk::BLOCK - 1) / deqk::BLOCK), P, st, err);
}

}  // namespace deqn
