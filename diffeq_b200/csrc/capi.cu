// capi.cu — implementation of the C ABI declared in include/diffeq_b200.h.
//
// Host-side logic of the reference that lives here (file:line under /root/reference/src/ode/):
//   solve dispatch Ode -> tableau -> loop        problem.rs:133-190
//   option defaults / minstep / maxstep           options.rs:283-299, problem.rs:219-225
//   initstep sign check -> InvalidInitstep        problem.rs:232-240
//   empty tspan handling                          problem.rs:211-214 (adaptive: empty Ok), :377 (fixed: panics)
// Device memory is owned by the handles; all work goes to the calling thread's stream (deq_set_stream).
#include <cuda_runtime.h>
#include <cfloat>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/diffeq_b200.h"
#include "launch.h"
#include "nvrtc_rhs.h"

namespace {

thread_local std::string g_err;
thread_local cudaStream_t g_stream = nullptr;

deq_status fail(deq_status s, const std::string& msg) { g_err = msg; return s; }

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            cudaGetLastError(); /* clear the per-thread last-error so the next call starts clean */    \
            return fail(DEQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
        }                                                                                              \
    } while (0)

const deql::TbEntry g_tb[deql::TB_COUNT] = {
#define EF(ID) {deql::launch_adapt_##ID, nullptr, nullptr, deql::launch_fixed_##ID, deql::info_##ID}  // fixed-step tableaus
#define E(ID) {deql::launch_adapt_##ID, deql::launch_adapt_fast_##ID, deql::launch_adapt_f32_##ID, deql::launch_fixed_##ID, deql::info_##ID}
    EF(0), EF(1), EF(2), EF(3), E(4), E(5), E(6), E(7), E(8), E(9), E(10), E(11), E(12)
#undef E
#undef EF
};

// (method, set) -> tableau id; -1 unsupported, -2 invalid
int tb_of(int method, int set) {
    const bool t = set == DEQ_TABLEAU_TEXTBOOK;
    switch (method) {
        case DEQ_FEULER: return deql::TB_FEULER;
        case DEQ_MIDPOINT: return deql::TB_MIDPOINT;
        case DEQ_HEUN: return deql::TB_HEUN;
        case DEQ_ODE4: return deql::TB_RK4;
        case DEQ_ODE21: return t ? deql::TB_RK21_TXT : deql::TB_RK21_REF;
        case DEQ_ODE23: return t ? deql::TB_RK23_TXT : deql::TB_RK23_REF;
        case DEQ_ODE45FE: return deql::TB_RK45;
        case DEQ_ODE45: return t ? deql::TB_DOPRI5_TXT : deql::TB_DOPRI5_REF;
        case DEQ_ODE78: return t ? deql::TB_FEH78_TXT : deql::TB_FEH78_REF;
        case DEQ_ODE23S: case DEQ_ODE4SKR: case DEQ_ODE4SS: return -1;
        default: return -2;
    }
}

double signum(double x) { return std::isnan(x) ? x : std::copysign(1.0, x); }  // f64::signum

// Stiff members of the `Ode` enum (mod.rs:14-26; problem.rs:139,143-144): 1 ode23s, 2 oderosenbrock kr4, 3 oderosenbrock s4.
int stiff_kind(int method) {
    switch (method) { case DEQ_ODE23S: return 1; case DEQ_ODE4SKR: return 2; case DEQ_ODE4SS: return 3; default: return 0; }
}

}  // namespace

namespace deql {
const TbEntry& tb_entry(int tb) { return g_tb[tb]; }
}

struct deq_rhs {
    int kind;  // 0 builtin, 1 nvrtc
    int sys;
    int n, np;
    deqn::Program* prog;  // nvrtc only (shared, ref-counted)
};

struct deq_ensemble {
    deq_rhs rhs;
    size_t ntraj = 0, nt = 0;
    int per_traj = 0;
    double* d_y0 = nullptr;      // SoA [n][ntraj]
    bool own_y0 = true;
    double* d_params = nullptr;  // SoA [np][ntraj]
    bool own_params = true;
    deqk::SharedParams shared{};
    std::vector<double> tspan;
    double* d_tspan = nullptr;
    bool has_tspan = false;
    cudaStream_t stream = nullptr;   // the stream its device arrays were allocated (and are freed) on
};

struct deq_result {
    size_t ntraj = 0, nt = 0;
    int n = 0;
    bool fixed = false, dense = false, traj_major = false;
    bool rosen = false;  // oderosenbrock: a fixed-grid solve that still carries per-trajectory counters and status
    cudaStream_t stream = nullptr;
    double* d_yfinal = nullptr;
    uint32_t *d_acc = nullptr, *d_rej = nullptr, *d_nfe = nullptr, *d_nvalid = nullptr;
    uint8_t* d_status = nullptr;
    double* d_treached = nullptr;
    double* d_dense = nullptr;
    double *d_f0 = nullptr, *d_h0 = nullptr;
    // DEQ_OUT_ALL: the reference's ragged (tout, yout) per trajectory
    bool all = false;
    uint32_t* d_rec_counts = nullptr;
    unsigned long long* d_rec_offsets = nullptr;
    double* d_rec_rows = nullptr;  // rows [t, y...] interleaved, trajectory-major
    std::vector<unsigned long long> rec_offsets;  // [ntraj + 1], host
    unsigned long long* d_scalars = nullptr;  // [0] work counter, [1..3] totals
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;  // both events were recorded
    bool totals_pending = false;  // adapt_kernel solves: d_scalars[1..3] are summed from the counter arrays on first request
    uint64_t fixed_nfe_per_traj = 0;
};

namespace {

// Device memory comes from a LIBRARY-OWNED stream-ordered pool per device (not the device's default pool, which a
// co-resident allocator such as torch's may be using): freed blocks stay cached up to kPoolKeepBytes, so repeated solves do
// not pay cudaMalloc-class latencies for their ~150 MB of result arrays (with the default threshold of 0 every synchronize
// hands the memory back to the driver: 91 ms instead of 42 ms per end-to-end step), while a freed multi-GB dense buffer is
// returned to the driver at the next synchronisation point.
constexpr unsigned long long kPoolKeepBytes = 2ull << 30;
cudaMemPool_t pool_for_current_device() {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) { cudaGetLastError(); pools[dev] = nullptr; return nullptr; }
        unsigned long long thr = kPoolKeepBytes;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return pools[dev];
}

template <class T>
cudaError_t dalloc(T** p, size_t count, cudaStream_t st) {
    *p = nullptr;
    if (count == 0) return cudaSuccess;
    cudaMemPool_t pool = pool_for_current_device();
    if (!pool) return cudaMallocAsync((void**)p, count * sizeof(T), st);
    return cudaMallocFromPoolAsync((void**)p, count * sizeof(T), pool, st);
}
void dfree(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }

void ensure_pool() {}   // (kept for the call sites: the pool is created on first allocation)

// Temporary device buffer that is released on every exit path (the CK macro returns early on errors).
template <class T>
struct TempBuf {
    T* p = nullptr;
    cudaStream_t st;
    explicit TempBuf(cudaStream_t s) : st(s) {}
    ~TempBuf() { dfree(p, st); }
    TempBuf(const TempBuf&) = delete;
    TempBuf& operator=(const TempBuf&) = delete;
};

// Host threads of the multi-device calls.  A fresh std::thread pays the CUDA runtime's per-thread set-up on its first call
// (milliseconds, every call); the workers here are created once, stay bound to their device between calls and keep a
// non-blocking stream of their own.  One multi-device call uses the pool at a time (a second concurrent caller spawns plain
// threads); the pool is leaked on purpose at process exit (its threads sleep on a condition variable).
class WorkerPool {
public:
    // runs fn(k, stream_of_worker_k) for k = 0..n-1 concurrently, worker k on device devs[k]; returns false if the pool is busy
    bool run(const std::vector<int>& devs, const std::function<void(size_t, cudaStream_t)>& fn) {
        std::unique_lock<std::mutex> busy(busy_, std::try_to_lock);
        if (!busy.owns_lock()) return false;
        while (workers_.size() < devs.size()) workers_.push_back(new Worker());
        for (size_t k = 0; k < devs.size(); ++k) workers_[k]->post(devs[k], k, &fn);
        for (size_t k = 0; k < devs.size(); ++k) workers_[k]->wait();
        return true;
    }
private:
    struct Worker {
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        const std::function<void(size_t, cudaStream_t)>* job = nullptr;
        size_t index = 0;
        int want_dev = -1, cur_dev = -1;
        bool pending = false, finished = false;
        cudaStream_t st = nullptr;
        Worker() { th = std::thread([this] { loop(); }); th.detach(); }
        void post(int dev, size_t k, const std::function<void(size_t, cudaStream_t)>* fn) {
            std::lock_guard<std::mutex> lk(m);
            want_dev = dev; index = k; job = fn; pending = true; finished = false;
            cv.notify_all();
        }
        void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return finished; }); }
        void loop() {
            for (;;) {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return pending; });
                pending = false;
                const auto* fn = job; const size_t k = index; const int dev = want_dev;
                lk.unlock();
                if (dev != cur_dev) {   // (re)bind: device and a stream on it
                    if (st) { cudaSetDevice(cur_dev); cudaStreamDestroy(st); st = nullptr; }
                    cur_dev = -1;
                    if (cudaSetDevice(dev) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess) cur_dev = dev;
                    else { cudaGetLastError(); st = nullptr; }
                }
                (*fn)(k, cur_dev == dev ? st : nullptr);
                lk.lock();
                finished = true;
                cv.notify_all();
            }
        }
    };
    std::mutex busy_;
    std::vector<Worker*> workers_;
};
WorkerPool& worker_pool() { static WorkerPool* p = new WorkerPool(); return *p; }

// Block-cyclic shards of deq_solve_host: the rows of this shard are runs (start, length) of the caller's arrays.  While the
// list is set, upload_soa treats its host pointer as the base of the FULL array and copies run by run, so no host-side
// packing is needed (and results go back the same way, download_runs).
typedef std::vector<std::pair<size_t, size_t>> Runs;
thread_local const Runs* g_runs = nullptr;

deq_status upload_soa(const double* h_aos, size_t ntraj, int n, double** d_soa, cudaStream_t st) {
    TempBuf<double> aos(st);
    CK(dalloc(&aos.p, ntraj * n, st));
    CK(dalloc(d_soa, ntraj * n, st));
    if (g_runs) {
        size_t w = 0;
        for (const auto& run : *g_runs) {
            CK(cudaMemcpyAsync(aos.p + w * n, h_aos + run.first * n, run.second * n * sizeof(double), cudaMemcpyHostToDevice, st));
            w += run.second;
        }
    } else {
        CK(cudaMemcpyAsync(aos.p, h_aos, ntraj * n * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CK(deql::launch_transpose_to_soa(aos.p, *d_soa, ntraj, n, st));
    return DEQ_OK;
}

// device array of `width`-byte rows, packed in shard order -> the runs of the caller's array `h_base`
deq_status download_runs(const void* d_packed, size_t width, void* h_base, const Runs& runs, cudaStream_t st) {
    if (!h_base) return DEQ_OK;
    size_t w = 0;
    for (const auto& run : runs) {
        CK(cudaMemcpyAsync((char*)h_base + run.first * width, (const char*)d_packed + w * width, run.second * width, cudaMemcpyDeviceToHost, st));
        w += run.second;
    }
    return DEQ_OK;
}

deq_status ensemble_common(const deq_rhs* rhs, size_t ntraj, const double* params, int per_traj, const double* tspan,
                           size_t ntspan, bool params_on_device, deq_ensemble* e) {
    e->rhs = *rhs;
    if (rhs->prog) deqn::retain(rhs->prog);
    e->ntraj = ntraj;
    e->per_traj = per_traj ? 1 : 0;
    cudaStream_t st = g_stream;
    e->stream = st;
    if (rhs->np > 0) {
        if (!params) return fail(DEQ_ERR_UNINITIALIZED, "Required problem parameters must be initialized");
        if (per_traj) {
            if (params_on_device) { e->d_params = const_cast<double*>(params); e->own_params = false; }
            else { deq_status s = upload_soa(params, ntraj, rhs->np, &e->d_params, st); if (s) return s; }
        } else {
            for (int j = 0; j < rhs->np; ++j) e->shared.v[j] = params[j];
        }
    }
    if (tspan) {
        e->has_tspan = true;
        e->nt = ntspan;
        e->tspan.assign(tspan, tspan + ntspan);
        CK(dalloc(&e->d_tspan, ntspan, st));
        if (ntspan) CK(cudaMemcpyAsync(e->d_tspan, e->tspan.data(), ntspan * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
    return DEQ_OK;
}

}  // namespace

extern "C" {

const char* deq_last_error_message(void) { return g_err.c_str(); }
const char* deq_version(void) { return "diffeq_b200 0.1 (sm_100a)"; }
size_t deq_abi_options_size(void) { return sizeof(deq_options); }

deq_status deq_rhs_builtin(int system, deq_rhs** out) {
    if (!out) return fail(DEQ_ERR_INVALID_ARG, "out is NULL");
    int n, np;
    switch (system) {
        case DEQ_SYS_LORENZ: n = 3; np = 3; break;
        case DEQ_SYS_VANDERPOL: n = 2; np = 1; break;
        case DEQ_SYS_PENDULUM: n = 2; np = 3; break;
        default: return fail(DEQ_ERR_INVALID_ARG, "unknown built-in system");
    }
    *out = new deq_rhs{0, system, n, np, nullptr};
    return DEQ_OK;
}

deq_status deq_rhs_from_source(const char* cuda_src, int n, int nparams, deq_rhs** out) {
    if (!out || !cuda_src) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (n < 1 || n > deqk::MAX_N || nparams < 0 || nparams > deqk::MAX_NP)
        return fail(DEQ_ERR_INVALID_ARG, "need 1 <= n <= 64 and 0 <= nparams <= 8");
    std::string err;
    deqn::Program* p = deqn::create(cuda_src, n, nparams, &err);
    if (!p) return fail(DEQ_ERR_NVRTC, err);
    *out = new deq_rhs{1, -1, n, nparams, p};
    return DEQ_OK;
}

int deq_rhs_dim(const deq_rhs* rhs) { return rhs ? rhs->n : -1; }
int deq_rhs_nparams(const deq_rhs* rhs) { return rhs ? rhs->np : -1; }
void deq_rhs_destroy(deq_rhs* rhs) {
    if (!rhs) return;
    if (rhs->prog) deqn::release(rhs->prog);
    delete rhs;
}

deq_status deq_tspan_linspace(double from, double to, size_t n, double* out) {
    if (n && !out) return fail(DEQ_ERR_INVALID_ARG, "out is NULL");
    const double step = n > 1 ? (to - from) / (double)(n - 1) : 0.0;
    for (size_t i = 0; i < n; ++i) out[i] = from + step * (double)i;
    return DEQ_OK;
}

deq_status deq_set_stream(void* s) { g_stream = (cudaStream_t)s; return DEQ_OK; }
deq_status deq_set_device(int device) { CK(cudaSetDevice(device)); return DEQ_OK; }
deq_status deq_get_device(int* device) { if (!device) return fail(DEQ_ERR_INVALID_ARG, "NULL argument"); CK(cudaGetDevice(device)); return DEQ_OK; }

deq_status deq_ensemble_create(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params,
                               int params_per_traj, const double* tspan, size_t ntspan, deq_ensemble** out) {
    if (!out) return fail(DEQ_ERR_INVALID_ARG, "out is NULL");
    if (!rhs) return fail(DEQ_ERR_UNINITIALIZED, "Required problem must be initialized");
    if (!y0 && ntraj) return fail(DEQ_ERR_UNINITIALIZED, "Initial starting point must be initialized");
    if (!tspan) return fail(DEQ_ERR_UNINITIALIZED, "Time span must be initialized");
    ensure_pool();
    deq_ensemble* e = new deq_ensemble();
    e->stream = g_stream;
    deq_status s = upload_soa(y0, ntraj, rhs->n, &e->d_y0, g_stream);
    if (!s) s = ensemble_common(rhs, ntraj, params, params_per_traj, tspan, ntspan, false, e);
    if (s) { deq_ensemble_destroy(e); return s; }
    *out = e;
    return DEQ_OK;
}

deq_status deq_ensemble_create_device(const deq_rhs* rhs, size_t ntraj, const double* d_y0_soa, const double* params,
                                      int params_per_traj, const double* tspan, size_t ntspan, deq_ensemble** out) {
    if (!out) return fail(DEQ_ERR_INVALID_ARG, "out is NULL");
    if (!rhs) return fail(DEQ_ERR_UNINITIALIZED, "Required problem must be initialized");
    if (!d_y0_soa && ntraj) return fail(DEQ_ERR_UNINITIALIZED, "Initial starting point must be initialized");
    if (!tspan) return fail(DEQ_ERR_UNINITIALIZED, "Time span must be initialized");
    deq_ensemble* e = new deq_ensemble();
    e->d_y0 = const_cast<double*>(d_y0_soa);
    e->own_y0 = false;
    deq_status s = ensemble_common(rhs, ntraj, params, params_per_traj, tspan, ntspan, true, e);
    if (s) { deq_ensemble_destroy(e); return s; }
    *out = e;
    return DEQ_OK;
}

void deq_ensemble_destroy(deq_ensemble* e) {
    if (!e) return;
    if (e->own_y0) dfree(e->d_y0, e->stream);
    if (e->own_params) dfree(e->d_params, e->stream);
    dfree(e->d_tspan, e->stream);
    if (e->rhs.prog) deqn::release(e->rhs.prog);
    delete e;
}

size_t deq_ensemble_ntraj(const deq_ensemble* e) { return e ? e->ntraj : 0; }

deq_options deq_options_default(void) {
    deq_options o;
    std::memset(&o, 0, sizeof(o));
    o.reltol = 1e-5;   // options.rs:283-287
    o.abstol = 1e-8;   // options.rs:289-293
    o.points = DEQ_POINTS_ALL;
    o.output = DEQ_OUT_FINAL;
    o.refill = 2;
    o.max_steps = 10000000ull;  // guard (extension): the reference can loop forever
    return o;
}

// (method, set) -> tableau id + description.  The stiff methods have no Butcher tableau: ode23s is described as an adaptive
// 3-stage FSAL scheme whose hinit order is 3 (problem.rs:435), oderosenbrock as a 4-stage fixed-grid scheme.
static deq_status resolve_method(int method, int tableau_set, int* tb, int* sk, deql::TableauInfo* info) {
    *tb = tb_of(method, tableau_set);
    *sk = stiff_kind(method);
    if (*tb < 0 && !*sk) return fail(DEQ_ERR_INVALID_ARG, "unknown method");
    if (*sk) {
        std::memset(info, 0, sizeof(*info));
        info->S = *sk == 1 ? 3 : 4; info->adaptive = *sk == 1; info->order_min = 3; info->fsal = *sk == 1;
    } else {
        g_tb[*tb].info(info);
    }
    return DEQ_OK;
}

deq_status deq_solve(deq_ensemble* e, int method, int tableau_set, const deq_options* opts_in, deq_result** out) {
    if (!e || !out) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (tableau_set != DEQ_TABLEAU_REFERENCE && tableau_set != DEQ_TABLEAU_TEXTBOOK)
        return fail(DEQ_ERR_INVALID_ARG, "unknown tableau set");
    int tb = -2, sk = 0;
    deql::TableauInfo info;
    if (deq_status ms = resolve_method(method, tableau_set, &tb, &sk, &info)) return ms;
    const deq_options o = opts_in ? *opts_in : deq_options_default();
    if (o.mode != DEQ_MODE_PARITY && o.mode != DEQ_MODE_FAST) return fail(DEQ_ERR_INVALID_ARG, "unknown arithmetic mode");
    if (o.precision != DEQ_PRECISION_F64 && o.precision != DEQ_PRECISION_F32) return fail(DEQ_ERR_INVALID_ARG, "unknown precision");
    if (o.libm != DEQ_LIBM_AS_WRITTEN && o.libm != DEQ_LIBM_LLVM_FOLDED) return fail(DEQ_ERR_INVALID_ARG, "unknown libm flavour");
    if (o.dense_layout != DEQ_LAYOUT_SOA && o.dense_layout != DEQ_LAYOUT_TRAJ_MAJOR) return fail(DEQ_ERR_INVALID_ARG, "unknown dense layout");
    if (o.dense_layout == DEQ_LAYOUT_TRAJ_MAJOR && (o.precision == DEQ_PRECISION_F32))
        return fail(DEQ_ERR_UNSUPPORTED, "the trajectory-major dense layout is implemented for the f64 kernels");
    if (o.precision == DEQ_PRECISION_F32 && (e->rhs.kind != 0 || o.output == DEQ_OUT_ALL))
        return fail(DEQ_ERR_UNSUPPORTED, "the f32 mode covers the built-in systems with FINAL / TSPAN output");
    const int n = e->rhs.n;
    const size_t ntraj = e->ntraj, nt = e->nt;
    if (sk) {
        if (n > 32) return fail(DEQ_ERR_UNSUPPORTED, "the stiff solvers carry systems of dimension <= 32 (Jacobian, W and its inverse are per-thread; see stiff.cuh)");
        if (o.precision != DEQ_PRECISION_F64) return fail(DEQ_ERR_UNSUPPORTED, "the stiff solvers are f64 only");
        if (o.output == DEQ_OUT_TSPAN && o.dense_layout != DEQ_LAYOUT_SOA) return fail(DEQ_ERR_UNSUPPORTED, "the stiff solvers write tspan rows in the SoA layout");
        if (o.points != DEQ_POINTS_ALL && o.points != DEQ_POINTS_SPECIFIED) return fail(DEQ_ERR_INVALID_ARG, "unknown points kind");
        // ode23s rescans the whole tspan on every accepted step (problem.rs:519); the kernel walks a cursor instead, which
        // emits the same rows only if the grid is monotone
        bool up = true, down = true;
        for (size_t i = 0; i + 1 < nt; ++i) { up = up && e->tspan[i] <= e->tspan[i + 1]; down = down && e->tspan[i] >= e->tspan[i + 1]; }
        if (sk == 1 && o.output != DEQ_OUT_FINAL && !(up || down)) return fail(DEQ_ERR_INVALID_ARG, "ode23s row output needs a monotone tspan");
    }
    ensure_pool();
    cudaStream_t st = g_stream;
    const bool want_dense = o.output == DEQ_OUT_TSPAN;
    const bool want_all = o.output == DEQ_OUT_ALL;
    if (o.output != DEQ_OUT_FINAL && !want_dense && !want_all) return fail(DEQ_ERR_INVALID_ARG, "unknown output kind");

    deq_result* r = new deq_result();
    r->ntraj = ntraj; r->nt = nt; r->n = n; r->stream = st; r->fixed = !info.adaptive; r->dense = want_dense;
    auto bail = [&](deq_status s) { deq_result_destroy(r); return s; };
#define CKR(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                        \
            return bail(fail(DEQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)));       \
        }                                                                                              \
    } while (0)

    CKR(cudaEventCreate(&r->ev0));
    CKR(cudaEventCreate(&r->ev1));
    CKR(dalloc(&r->d_scalars, 4, st));
    CKR(cudaMemsetAsync(r->d_scalars, 0, 4 * sizeof(unsigned long long), st));

    if (!info.adaptive && want_all)
        return bail(fail(DEQ_ERR_INVALID_ARG, "fixed-step methods output exactly the tspan rows: use DEQ_OUT_TSPAN"));
    if (!info.adaptive) {
        // ---- oderk_fixed / oderosenbrock ----
        if (nt == 0 && sk) { *out = r; r->ntraj = 0; return DEQ_OK; }  // problem.rs:567-570: nothing to solve
        if (nt == 0) return bail(fail(DEQ_ERR_INVALID_ARG, "fixed-step methods need a non-empty tspan (the reference panics, problem.rs:377)"));
        CKR(dalloc(&r->d_yfinal, ntraj * n, st));
        if (want_dense) CKR(dalloc(&r->d_dense, ntraj * n * nt, st));
        deqk::FixedParams P{};
        P.y0 = e->d_y0; P.params = e->d_params; P.shared = e->shared; P.tspan = e->d_tspan; P.nt = nt; P.ntraj = ntraj;
        P.yfinal = r->d_yfinal; P.yall = r->d_dense;
        r->fixed_nfe_per_traj = (uint64_t)info.S * (nt - 1);
        if (sk) {  // a Rosenbrock step can stop a trajectory (InvalidMatrix): per-trajectory bookkeeping like the adaptive path
            r->rosen = true;
            CKR(dalloc(&r->d_acc, ntraj, st)); CKR(dalloc(&r->d_rej, ntraj, st)); CKR(dalloc(&r->d_nfe, ntraj, st));
            CKR(dalloc(&r->d_status, ntraj, st)); CKR(dalloc(&r->d_treached, ntraj, st)); CKR(dalloc(&r->d_nvalid, ntraj, st));
            P.accepted = r->d_acc; P.rejected = r->d_rej; P.nfe = r->d_nfe; P.status = r->d_status; P.t_reached = r->d_treached;
            P.nvalid = r->d_nvalid; P.totals = r->d_scalars + 1;
        }
        CKR(cudaEventRecord(r->ev0, st));
        if (ntraj) {
            std::string err;
            if (sk) {
                if (e->rhs.kind == 0) CKR(deql::launch_rosenbrock(sk - 2, e->rhs.sys, e->per_traj, want_dense ? 1 : 0, P, st));
                else if (!deqn::launch_rosenbrock(e->rhs.prog, sk - 2, e->per_traj, want_dense ? 1 : 0, P, st, &err)) return bail(fail(DEQ_ERR_NVRTC, err));
            } else if (e->rhs.kind == 0) CKR(g_tb[tb].fixed(e->rhs.sys, e->per_traj, want_dense ? 1 : 0, P, st));
            else if (!deqn::launch_fixed(e->rhs.prog, tb, e->per_traj, want_dense ? 1 : 0, P, st, &err)) return bail(fail(DEQ_ERR_NVRTC, err));
        }
        CKR(cudaEventRecord(r->ev1, st));
    } else {
        // ---- oderk_adapt ----
        // Points::Specified on the explicit path (problem.rs:284-300) is reproduced with the reference's defect (y re-read from
        // the output vector, :262): parity arithmetic, final state or the full stream (whose rows are exactly the tspan points)
        const bool spec = !sk && o.points == DEQ_POINTS_SPECIFIED;
        if (!sk && o.points != DEQ_POINTS_ALL && !spec) return bail(fail(DEQ_ERR_INVALID_ARG, "unknown points kind"));
        if (spec && (o.mode != DEQ_MODE_PARITY || o.precision != DEQ_PRECISION_F64))
            return bail(fail(DEQ_ERR_UNSUPPORTED, "Points::Specified is carried by the f64 parity kernels only"));
        if (spec && want_dense)
            return bail(fail(DEQ_ERR_UNSUPPORTED, "Points::Specified: use output = DEQ_OUT_ALL (every row of that stream is a tspan point) or DEQ_OUT_FINAL"));
        if (nt == 0) { *out = r; r->ntraj = 0; return DEQ_OK; }  // problem.rs:211-214: nothing to solve
        if (want_dense && nt < 2) return bail(fail(DEQ_ERR_INVALID_ARG, "DEQ_OUT_TSPAN needs at least 2 tspan points"));
        r->all = want_all;
        const double t0 = e->tspan[0], tend = e->tspan[nt - 1];
        const double tdir = signum(tend - t0);
        if (tdir == 0.) return bail(fail(DEQ_ERR_ZERO_TIMESPAN, "Zero time span is not allowed"));
        if (!sk && o.initstep != 0. && !(std::fabs(signum(o.initstep) - tdir) < DBL_EPSILON))   // ode23s takes |initstep| (problem.rs:443)
            return bail(fail(DEQ_ERR_INVALID_INITSTEP, "Initial step has wrong sign"));
        CKR(dalloc(&r->d_yfinal, ntraj * n, st));
        CKR(dalloc(&r->d_acc, ntraj, st)); CKR(dalloc(&r->d_rej, ntraj, st)); CKR(dalloc(&r->d_nfe, ntraj, st));
        CKR(dalloc(&r->d_status, ntraj, st)); CKR(dalloc(&r->d_treached, ntraj, st));
        CKR(dalloc(&r->d_f0, ntraj * n, st)); CKR(dalloc(&r->d_h0, ntraj, st));
        if (want_dense) { CKR(dalloc(&r->d_dense, ntraj * n * nt, st)); CKR(dalloc(&r->d_nvalid, ntraj, st)); }
        if (want_all || spec) { CKR(dalloc(&r->d_nvalid, ntraj, st)); CKR(dalloc(&r->d_rec_counts, ntraj, st)); CKR(dalloc(&r->d_rec_offsets, ntraj + 1, st)); }
        deqk::HinitParams H{};
        H.y0 = e->d_y0; H.params = e->d_params; H.shared = e->shared; H.ntraj = ntraj; H.t0 = t0; H.tend = tend;
        H.reltol = o.reltol; H.abstol = o.abstol; H.order = info.order_min; H.f0 = r->d_f0; H.h0 = r->d_h0;
        H.libm_folded = o.libm == DEQ_LIBM_LLVM_FOLDED;
        deqk::AdaptParams P{};
        P.y0 = e->d_y0; P.f0 = r->d_f0; P.h0 = r->d_h0; P.params = e->d_params; P.shared = e->shared;
        P.tspan = e->d_tspan; P.nt = nt; P.ntraj = ntraj; P.t0 = t0; P.tend = tend; P.tdir = tdir; P.tdir_tend = tdir * tend;
        P.minstep = o.has_minstep ? o.minstep : std::fabs(tend - t0) / 1e18;   // problem.rs:219-221
        P.maxstep = o.has_maxstep ? o.maxstep : std::fabs(tend - t0) / 2.5;    // problem.rs:223-225
        P.reltol = o.reltol; P.abstol = o.abstol; P.initstep = o.initstep; P.use_initstep = o.initstep != 0.;
        P.refill = o.refill; P.max_steps = o.max_steps; P.libm_folded = o.libm == DEQ_LIBM_LLVM_FOLDED;
        P.yfinal = r->d_yfinal; P.accepted = r->d_acc; P.rejected = r->d_rej; P.nfe = r->d_nfe; P.status = r->d_status;
        P.t_reached = r->d_treached; P.dense = r->d_dense; P.nvalid = r->d_nvalid;
        P.work_counter = r->d_scalars; P.totals = r->d_scalars + 1;
        P.true_inverse = sk && tableau_set == DEQ_TABLEAU_TEXTBOOK; P.points_specified = o.points == DEQ_POINTS_SPECIFIED;
        int mode = spec ? (int)deqk::MODE_REC_SPEC   // a final-state request runs its counting pass only
                         : want_all ? (int)deqk::MODE_REC
                         : want_dense ? (o.dense_layout == DEQ_LAYOUT_TRAJ_MAJOR ? (int)deqk::MODE_DENSE_TM
                                         // SoA rows: the row-synchronous kernel (coalesced for any trajectory order) unless the caller
                                         // asks for the lane-refill kernel (refill == 1), the stiff solver or the f32 flavour
                                         : (!sk && o.precision == DEQ_PRECISION_F64 && o.refill != 1) ? (int)deqk::MODE_DENSE_ROWS : (int)deqk::MODE_DENSE)
                         : (int)deqk::MODE_FINAL;
        P.rec_mode = deqk::REC_COUNT;
        r->traj_major = want_dense && o.dense_layout == DEQ_LAYOUT_TRAJ_MAJOR;
        P.rec_counts = r->d_rec_counts;
        auto launch_adapt = [&](std::string* err) -> int {  // 0 ok, 1 cuda error (g_err set), 2 nvrtc error
            if (sk) {  // ode23s
                const bool fast = o.mode == DEQ_MODE_FAST;
                if (e->rhs.kind != 0) return deqn::launch_ode23s(e->rhs.prog, e->per_traj, mode, fast, P, st, err) ? 0 : 2;
                cudaError_t ce = fast ? deql::launch_ode23s_fast(e->rhs.sys, e->per_traj, mode, P, st) : deql::launch_ode23s(e->rhs.sys, e->per_traj, mode, P, st);
                if (ce != cudaSuccess) { fail(DEQ_ERR_CUDA, std::string("ode23s kernel launch: ") + cudaGetErrorString(ce)); return 1; }
                return 0;
            }
            if (e->rhs.kind == 0) {
                deql::AdaptLaunchFn fn = o.precision == DEQ_PRECISION_F32 ? g_tb[tb].adapt_f32 : o.mode == DEQ_MODE_FAST ? g_tb[tb].adapt_fast : g_tb[tb].adapt;
                cudaError_t ce = fn(e->rhs.sys, e->per_traj, mode, P, 0, st);
                if (ce != cudaSuccess) { fail(DEQ_ERR_CUDA, std::string("adaptive kernel launch: ") + cudaGetErrorString(ce)); return 1; }
                return 0;
            }
            return deqn::launch_adapt(e->rhs.prog, tb, e->per_traj, mode, o.mode == DEQ_MODE_FAST, P, st, err) ? 0 : 2;
        };
        if (ntraj && want_all) {
            // two passes: count the rows each trajectory emits, prefix-sum, then re-run writing them.  The
            // solver is deterministic, so both passes take identical steps.
            std::string err;
            if (e->rhs.kind == 0) CKR(deql::launch_hinit(e->rhs.sys, e->per_traj, H, st));
            else if (!deqn::launch_hinit(e->rhs.prog, e->per_traj, H, st, &err)) return bail(fail(DEQ_ERR_NVRTC, err));
            int rc = launch_adapt(&err);
            if (rc == 1) return bail(DEQ_ERR_CUDA);
            if (rc == 2) return bail(fail(DEQ_ERR_NVRTC, err));
            std::vector<uint32_t> counts(ntraj);
            CKR(cudaMemcpyAsync(counts.data(), r->d_rec_counts, ntraj * 4, cudaMemcpyDeviceToHost, st));
            CKR(cudaStreamSynchronize(st));
            r->rec_offsets.assign(ntraj + 1, 0ull);
            for (size_t i = 0; i < ntraj; ++i) r->rec_offsets[i + 1] = r->rec_offsets[i] + counts[i];
            const unsigned long long total = r->rec_offsets[ntraj];
            CKR(cudaMemcpyAsync(r->d_rec_offsets, r->rec_offsets.data(), (ntraj + 1) * 8, cudaMemcpyHostToDevice, st));
            CKR(dalloc(&r->d_rec_rows, total * (n + 1), st));
            CKR(cudaMemsetAsync(r->d_scalars, 0, 4 * sizeof(unsigned long long), st));
            P.rec_mode = deqk::REC_WRITE; P.rec_offsets = r->d_rec_offsets; P.rec_rows = r->d_rec_rows;
            CKR(cudaEventRecord(r->ev0, st));
            rc = launch_adapt(&err);
            if (rc == 1) return bail(DEQ_ERR_CUDA);
            if (rc == 2) return bail(fail(DEQ_ERR_NVRTC, err));
        } else if (ntraj) {
            std::string err;
            if (e->rhs.kind == 0) CKR(deql::launch_hinit(e->rhs.sys, e->per_traj, H, st));
            else if (!deqn::launch_hinit(e->rhs.prog, e->per_traj, H, st, &err)) return bail(fail(DEQ_ERR_NVRTC, err));
            bool launched = false, ev0_recorded = false;
            // (the trajectory-major layout profits from the same ordering — its lanes then drain their staging rings together and
            // no column permutation is needed, every stream is written where it belongs: 30 -> 21 ms on the permuted sweep)
            const bool probe_tm = mode == (int)deqk::MODE_DENSE_TM && o.refill == 2 && o.initstep == 0.;
            if ((mode == (int)deqk::MODE_DENSE_ROWS && o.refill == 2 && o.initstep == 0.) || probe_tm) {
                // SoA rows, choice left open: neighbours that start alike (a parameter-sorted sweep) advance together, and the
                // lane-refill kernel's own stores coalesce (and it keeps its lanes busier); otherwise the row-synchronous kernel
                unsigned long long hc[2] = {0, 0};
                CKR(cudaMemsetAsync(r->d_scalars + 1, 0, 2 * sizeof(unsigned long long), st));
                CKR(deql::launch_coherence(r->d_h0, ntraj, r->d_scalars + 1, st));
                CKR(cudaMemcpyAsync(hc, r->d_scalars + 1, sizeof(hc), cudaMemcpyDeviceToHost, st));
                CKR(cudaStreamSynchronize(st));
                CKR(cudaMemsetAsync(r->d_scalars + 1, 0, 2 * sizeof(unsigned long long), st));
                if (hc[1] > 0 && (double)hc[0] <= 0.02 * (double)hc[1]) { if (!probe_tm) mode = (int)deqk::MODE_DENSE; }
                else if (ntraj >= 4096 && ntraj < 0x7fffffffull && !getenv("DEQ_NO_SORTED_DENSE")) {   // (a smaller ensemble is one partial wave: nothing to gain, a sort and a second host round trip to lose)
                    // Not coherent as given.  Sorted by first step size it may be (a parameter sweep handed over in arbitrary order
                    // is): then the lane-refill kernel runs the work items in that order, its row stores coalesce again — columns in
                    // the order of the work items, and an in-place pass through an L2-sized scratch puts the columns back where the
                    // caller expects them (HBM-bound, each byte about once in each direction; no second copy of the rows — config 3
                    // is 67 GB).  The timed region starts here: sort, kernel and column permutation.
                    TempBuf<uint32_t> order(st), inv(st);
                    TempBuf<double> hs(st), rows(st);
                    TempBuf<char> scratch(st);
                    const size_t sb = deql::sort_by_h0_scratch_bytes(ntraj);
                    CKR(dalloc(&order.p, ntraj, st)); CKR(dalloc(&inv.p, ntraj, st)); CKR(dalloc(&hs.p, ntraj, st)); CKR(dalloc(&scratch.p, sb, st));
                    TempBuf<char> plan(st), pscratch(st);
                    const bool staged = !probe_tm && ntraj >= 65536 && !getenv("DEQ_SIMPLE_PERMUTE");   // two coalesced stages (plan: 8 bytes per trajectory, built once)
                    if (!probe_tm) CKR(dalloc(&rows.p, 2 * deql::permute_rows_batch(ntraj) * ntraj + 3, st));   // + the barrier word and item counters of the persistent permutation kernel
                    if (staged) { CKR(dalloc(&plan.p, deql::permute_plan_bytes(ntraj), st)); CKR(dalloc(&pscratch.p, deql::permute_plan_scratch_bytes(ntraj), st)); }
                    CKR(cudaEventRecord(r->ev0, st));   // (memory is obtained outside the timed region, like the result arrays)
                    ev0_recorded = true;
                    CKR(deql::launch_sort_by_h0(r->d_h0, ntraj, order.p, inv.p, hs.p, scratch.p, sb, st));
                    CKR(deql::launch_coherence(hs.p, ntraj, r->d_scalars + 1, st));
                    CKR(cudaMemcpyAsync(hc, r->d_scalars + 1, sizeof(hc), cudaMemcpyDeviceToHost, st));
                    CKR(cudaStreamSynchronize(st));
                    CKR(cudaMemsetAsync(r->d_scalars + 1, 0, 2 * sizeof(unsigned long long), st));
                    if (hc[1] > 0 && (double)hc[0] <= 0.02 * (double)hc[1]) {
                        P.order = order.p;
                        if (!probe_tm) mode = (int)deqk::MODE_DENSE;
                        const int rc = launch_adapt(&err);
                        if (rc == 1) return bail(DEQ_ERR_CUDA);
                        if (rc == 2) return bail(fail(DEQ_ERR_NVRTC, err));
                        if (!probe_tm) {
                            if (staged) {
                                CKR(deql::launch_permute_plan(order.p, ntraj, plan.p, pscratch.p, st));
                                CKR(deql::launch_permute_rows_staged(r->d_dense, plan.p, (unsigned long long)n * nt, ntraj, rows.p, st));
                            } else {
                                CKR(deql::launch_permute_rows_inplace(r->d_dense, inv.p, (unsigned long long)n * nt, ntraj, rows.p, st));
                            }
                        }
                        P.order = nullptr;
                        launched = true;
                    }
                }
            }
            if (!launched) {
                if (!ev0_recorded) CKR(cudaEventRecord(r->ev0, st));
                const int rc = launch_adapt(&err);
                if (rc == 1) return bail(DEQ_ERR_CUDA);
                if (rc == 2) return bail(fail(DEQ_ERR_NVRTC, err));
            }
        } else {
            CKR(cudaEventRecord(r->ev0, st));
        }
        CKR(cudaEventRecord(r->ev1, st));
        r->totals_pending = !sk && o.precision == DEQ_PRECISION_F64;
    }
    r->timed = true;
    if (!o.async) CKR(cudaStreamSynchronize(st));
#undef CKR
    *out = r;
    return DEQ_OK;
}

deq_status deq_result_final(deq_result* r, double* y) {
    if (!r || !y) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (r->ntraj == 0) return DEQ_OK;
    TempBuf<double> aos(r->stream);
    CK(dalloc(&aos.p, r->ntraj * r->n, r->stream));
    CK(deql::launch_transpose_to_aos(r->d_yfinal, aos.p, r->ntraj, r->n, r->stream));
    CK(cudaMemcpyAsync(y, aos.p, r->ntraj * r->n * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return DEQ_OK;
}

deq_status deq_result_counts(deq_result* r, uint32_t* accepted, uint32_t* rejected, uint32_t* nfe, uint8_t* status,
                             double* t_reached) {
    if (!r) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (r->fixed && !r->rosen) return fail(DEQ_ERR_INVALID_ARG, "fixed-step solves carry no per-trajectory counters");
    if (r->ntraj == 0) return DEQ_OK;
    const size_t n = r->ntraj;
    if (accepted) CK(cudaMemcpyAsync(accepted, r->d_acc, n * 4, cudaMemcpyDeviceToHost, r->stream));
    if (rejected) CK(cudaMemcpyAsync(rejected, r->d_rej, n * 4, cudaMemcpyDeviceToHost, r->stream));
    if (nfe) CK(cudaMemcpyAsync(nfe, r->d_nfe, n * 4, cudaMemcpyDeviceToHost, r->stream));
    if (status) CK(cudaMemcpyAsync(status, r->d_status, n, cudaMemcpyDeviceToHost, r->stream));
    if (t_reached) CK(cudaMemcpyAsync(t_reached, r->d_treached, n * 8, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return DEQ_OK;
}

deq_status deq_result_totals(deq_result* r, uint64_t* accepted, uint64_t* rejected, uint64_t* nfe) {
    if (!r) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    unsigned long long h[4] = {0, 0, 0, 0};
    if (r->totals_pending && r->ntraj) {
        CK(deql::launch_sum_counts(r->d_acc, r->d_rej, r->d_nfe, r->ntraj, r->d_scalars + 1, r->stream));
        r->totals_pending = false;
    }
    if (r->d_scalars) {
        CK(cudaMemcpyAsync(h, r->d_scalars, sizeof(h), cudaMemcpyDeviceToHost, r->stream));
        CK(cudaStreamSynchronize(r->stream));
    }
    if (r->fixed && !r->rosen) {  // every trajectory takes nt-1 steps of S evaluations
        h[1] = (unsigned long long)r->ntraj * (r->nt ? r->nt - 1 : 0); h[2] = 0; h[3] = r->fixed_nfe_per_traj * r->ntraj;
    }
    if (accepted) *accepted = h[1];
    if (rejected) *rejected = h[2];
    if (nfe) *nfe = h[3];
    return DEQ_OK;
}

deq_status deq_result_dense(deq_result* r, size_t b, size_t e, double* out, uint32_t* nvalid) {
    if (!r || !out) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (!r->dense || !r->d_dense) return fail(DEQ_ERR_INVALID_ARG, "solve was not run with output = DEQ_OUT_TSPAN");
    if (b > e || e > r->ntraj) return fail(DEQ_ERR_INVALID_ARG, "trajectory range out of bounds");
    const size_t w = e - b;
    if (w == 0) return DEQ_OK;
    if (r->traj_major) {  // [ntraj][nt][n]: the slice is contiguous
        const size_t per = (size_t)r->n * r->nt;
        CK(cudaMemcpyAsync(out, r->d_dense + b * per, w * per * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
        if (nvalid) CK(cudaMemcpyAsync(nvalid, r->d_nvalid + b, w * 4, cudaMemcpyDeviceToHost, r->stream));
        CK(cudaStreamSynchronize(r->stream));
        return DEQ_OK;
    }
    // rows of [n*nt][ntraj] restricted to columns [b,e)
    CK(cudaMemcpy2DAsync(out, w * sizeof(double), r->d_dense + b, r->ntraj * sizeof(double), w * sizeof(double),
                         (size_t)r->n * r->nt, cudaMemcpyDeviceToHost, r->stream));
    if (nvalid) {
        if (r->fixed && !r->rosen) { CK(cudaStreamSynchronize(r->stream)); for (size_t i = 0; i < w; ++i) nvalid[i] = (uint32_t)r->nt; }
        else CK(cudaMemcpyAsync(nvalid, r->d_nvalid + b, w * 4, cudaMemcpyDeviceToHost, r->stream));
    }
    CK(cudaStreamSynchronize(r->stream));
    return DEQ_OK;
}

deq_status deq_result_all_offsets(deq_result* r, uint64_t* offsets) {
    if (!r || !offsets) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (!r->all) return fail(DEQ_ERR_INVALID_ARG, "solve was not run with output = DEQ_OUT_ALL");
    for (size_t i = 0; i < r->rec_offsets.size(); ++i) offsets[i] = r->rec_offsets[i];
    return DEQ_OK;
}

deq_status deq_result_all_copy(deq_result* r, size_t b, size_t e, double* tout, double* yout) {
    if (!r || !tout || !yout) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (!r->all) return fail(DEQ_ERR_INVALID_ARG, "solve was not run with output = DEQ_OUT_ALL");
    if (b > e || e > r->ntraj) return fail(DEQ_ERR_INVALID_ARG, "trajectory range out of bounds");
    const unsigned long long r0 = r->rec_offsets[b], r1 = r->rec_offsets[e];
    if (r1 > r0) {  // device rows are [t, y...] interleaved (one staged stream per trajectory): split them for the caller
        const size_t w = (size_t)r->n + 1, rows = (size_t)(r1 - r0);
        std::vector<double> tmp(rows * w);
        CK(cudaMemcpyAsync(tmp.data(), r->d_rec_rows + r0 * w, rows * w * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
        CK(cudaStreamSynchronize(r->stream));
        for (size_t i = 0; i < rows; ++i) {
            tout[i] = tmp[i * w];
            for (int d = 0; d < r->n; ++d) yout[i * r->n + d] = tmp[i * w + 1 + d];
        }
    }
    return DEQ_OK;
}

deq_status deq_result_device_ptrs(deq_result* r, const double** yf, const double** dn) {
    if (!r) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (yf) *yf = r->d_yfinal;
    if (dn) *dn = r->d_dense;
    return DEQ_OK;
}

deq_status deq_result_trajectory(deq_result* r, size_t traj, double* yout) {
    if (!r || !yout) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    if (!r->dense || !r->d_dense) return fail(DEQ_ERR_INVALID_ARG, "solve was not run with output = DEQ_OUT_TSPAN");
    if (traj >= r->ntraj) return fail(DEQ_ERR_INVALID_ARG, "trajectory index out of bounds");
    if (r->traj_major) {
        const size_t per = (size_t)r->n * r->nt;
        CK(cudaMemcpyAsync(yout, r->d_dense + traj * per, per * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
        CK(cudaStreamSynchronize(r->stream));
        return DEQ_OK;
    }
    std::vector<double> tmp((size_t)r->n * r->nt);
    CK(cudaMemcpy2DAsync(tmp.data(), sizeof(double), r->d_dense + traj, r->ntraj * sizeof(double), sizeof(double),
                         (size_t)r->n * r->nt, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    for (size_t i = 0; i < r->nt; ++i)
        for (int d = 0; d < r->n; ++d) yout[i * r->n + d] = tmp[(size_t)d * r->nt + i];
    return DEQ_OK;
}

deq_status deq_result_kernel_ms(deq_result* r, float* ms) {
    if (!r || !ms) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    *ms = 0.f;
    if (!r->timed || !r->ev0 || !r->ev1) return DEQ_OK;
    CK(cudaEventSynchronize(r->ev1));
    CK(cudaEventElapsedTime(ms, r->ev0, r->ev1));
    return DEQ_OK;
}

void deq_result_destroy(deq_result* r) {
    if (!r) return;
    cudaStream_t st = r->stream;
    dfree(r->d_yfinal, st); dfree(r->d_acc, st); dfree(r->d_rej, st); dfree(r->d_nfe, st); dfree(r->d_nvalid, st);
    dfree(r->d_status, st); dfree(r->d_treached, st); dfree(r->d_dense, st); dfree(r->d_f0, st); dfree(r->d_h0, st);
    dfree(r->d_scalars, st); dfree(r->d_rec_counts, st); dfree(r->d_rec_offsets, st); dfree(r->d_rec_rows, st);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    delete r;
}

// One-shot multi-GPU solve on host buffers: contiguous shards of the trajectory range, one host thread and one stream per
// device, no inter-GPU traffic; every shard writes its slice of the caller's output arrays.
static deq_status solve_host_dense_impl(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                                       const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options& o_in,
                                       const std::vector<int>& devs, int cur, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                                       uint32_t* nfe, uint8_t* status, double* t_reached, double* dense, uint32_t* nvalid);

static deq_status solve_host_impl(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                          const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options* opts,
                          const int* devices, int ndevices, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                          uint32_t* nfe, uint8_t* status, double* t_reached, double* dense, uint32_t* nvalid) {
    if (!rhs) return fail(DEQ_ERR_UNINITIALIZED, "Required problem must be initialized");
    if (!y0 && ntraj) return fail(DEQ_ERR_UNINITIALIZED, "Initial starting point must be initialized");
    if (!tspan) return fail(DEQ_ERR_UNINITIALIZED, "Time span must be initialized");
    if (!yfinal) return fail(DEQ_ERR_INVALID_ARG, "yfinal is NULL");
    deq_options o = opts ? *opts : deq_options_default();
    if (dense) {
        if (o.output != DEQ_OUT_TSPAN) return fail(DEQ_ERR_INVALID_ARG, "deq_solve_host_dense needs output = DEQ_OUT_TSPAN");
        int cur0 = 0;
        CK(cudaGetDevice(&cur0));
        std::vector<int> dv;
        if (devices && ndevices > 0) dv.assign(devices, devices + ndevices); else dv.push_back(cur0);
        return solve_host_dense_impl(rhs, ntraj, y0, params, params_per_traj, tspan, ntspan, method, tableau_set, o, dv, cur0, yfinal,
                                     accepted, rejected, nfe, status, t_reached, dense, nvalid);
    }
    if (o.output == DEQ_OUT_TSPAN) return fail(DEQ_ERR_INVALID_ARG, "deq_solve_host returns final states and counters; the tspan rows come from deq_solve_host_dense");
    if (o.output != DEQ_OUT_FINAL)
        return fail(DEQ_ERR_UNSUPPORTED, "DEQ_OUT_ALL has a data-dependent size (rows per trajectory): use deq_solve + deq_result_all_offsets / deq_result_all_copy");
    o.async = 0;
    int cur = 0;
    CK(cudaGetDevice(&cur));
    std::vector<int> devs;
    if (devices && ndevices > 0) devs.assign(devices, devices + ndevices); else devs.push_back(cur);
    const size_t nd = devs.size(), n = (size_t)rhs->n, np = (size_t)rhs->np;
    deql::TableauInfo info;
    int sk = 0;
    { int tb = -2; if (deq_status ms = resolve_method(method, tableau_set, &tb, &sk, &info)) return ms; }
    std::vector<deq_status> rc(nd, DEQ_OK);
    std::vector<std::string> msg(nd);
    const size_t chunk = (nd > 1 && o.shard_chunk > 0) ? (size_t)o.shard_chunk : 0;
    auto work = [&](size_t k, cudaStream_t pool_stream) {
        auto bad = [&](deq_status s) { rc[k] = s; msg[k] = g_err; };
        if (cudaSetDevice(devs[k]) != cudaSuccess) { rc[k] = DEQ_ERR_CUDA; msg[k] = "cudaSetDevice failed"; cudaGetLastError(); return; }
        // a worker thread runs its shard on a non-blocking stream of its own — the pool worker's persistent one, or a temporary one
        // (the calling thread keeps the stream it set with deq_set_stream)
        struct OwnStream {
            cudaStream_t st = nullptr; bool own = false, bound = false;
            ~OwnStream() { if (own) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } if (bound) g_stream = nullptr; }
        } ws;
        if (!(nd == 1 && devs[0] == cur)) {
            if (pool_stream) ws.st = pool_stream;
            else {
                if (cudaStreamCreateWithFlags(&ws.st, cudaStreamNonBlocking) != cudaSuccess) { rc[k] = DEQ_ERR_CUDA; msg[k] = "cudaStreamCreate failed"; cudaGetLastError(); return; }
                ws.own = true;
            }
            ws.bound = true;
            g_stream = ws.st;
        }
        const bool per = params && params_per_traj;
        // this shard's trajectories: one contiguous block, or (block-cyclic) the chunks k, k + nd, k + 2 nd, ... — copied run by
        // run straight between the caller's arrays and the device, in both directions
        size_t b = 0, cnt = 0;
        Runs runs;
        if (!chunk) {
            const size_t base = ntraj / nd, extra = ntraj % nd;
            b = k * base + (k < extra ? k : extra);
            cnt = base + (k < extra ? 1 : 0);
        } else {
            for (size_t c0 = k * chunk; c0 < ntraj; c0 += nd * chunk) { runs.emplace_back(c0, std::min(chunk, ntraj - c0)); cnt += runs.back().second; }
        }
        if (cnt == 0) return;   // more devices than chunks: nothing to do here
        deq_ensemble* ens = nullptr;
        deq_result* res = nullptr;
        g_runs = chunk ? &runs : nullptr;
        deq_status s = deq_ensemble_create(rhs, cnt, chunk ? y0 : y0 + b * n, (!per || chunk) ? params : params + b * np, params_per_traj, tspan, ntspan, &ens);
        g_runs = nullptr;
        if (s) { bad(s); return; }
        s = deq_solve(ens, method, tableau_set, &o, &res);
        const bool have = ntspan != 0, counts = info.adaptive || sk;
        if (!s && have && !chunk) {
            s = deq_result_final(res, yfinal + b * n);
            if (!s && counts)
                s = deq_result_counts(res, accepted ? accepted + b : nullptr, rejected ? rejected + b : nullptr, nfe ? nfe + b : nullptr,
                                      status ? status + b : nullptr, t_reached ? t_reached + b : nullptr);
        } else if (!s && have) {
            cudaStream_t st = res->stream;
            double* d_aos = nullptr;
            auto ck = [&](cudaError_t e_) { if (e_ != cudaSuccess && !s) { cudaGetLastError(); s = fail(DEQ_ERR_CUDA, cudaGetErrorString(e_)); } };
            ck(dalloc(&d_aos, cnt * n, st));
            if (!s) ck(deql::launch_transpose_to_aos(res->d_yfinal, d_aos, cnt, (int)n, st));
            if (!s) s = download_runs(d_aos, n * sizeof(double), yfinal, runs, st);
            if (!s && counts) s = download_runs(res->d_acc, 4, accepted, runs, st);
            if (!s && counts) s = download_runs(res->d_rej, 4, rejected, runs, st);
            if (!s && counts) s = download_runs(res->d_nfe, 4, nfe, runs, st);
            if (!s && counts) s = download_runs(res->d_status, 1, status, runs, st);
            if (!s && counts) s = download_runs(res->d_treached, 8, t_reached, runs, st);
            dfree(d_aos, st);
            ck(cudaStreamSynchronize(st));
        }
        if (s) bad(s);
        if (res) deq_result_destroy(res);
        deq_ensemble_destroy(ens);
    };
    if (nd == 1 && devs[0] == cur) {
        work(0, nullptr);
    } else {
        if (!worker_pool().run(devs, work)) {   // pool in use by a concurrent call: plain threads
            std::vector<std::thread> th;
            for (size_t k = 0; k < nd; ++k) th.emplace_back(work, k, (cudaStream_t) nullptr);
            for (auto& t : th) t.join();
        }
        cudaSetDevice(cur);
    }
    for (size_t k = 0; k < nd; ++k)
        if (rc[k]) return fail(rc[k], "device " + std::to_string(devs[k]) + ": " + msg[k]);
    return DEQ_OK;
}

deq_status deq_solve_host(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                          const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options* opts,
                          const int* devices, int ndevices, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                          uint32_t* nfe, uint8_t* status, double* t_reached) {
    return solve_host_impl(rhs, ntraj, y0, params, params_per_traj, tspan, ntspan, method, tableau_set, opts, devices, ndevices, yfinal,
                           accepted, rejected, nfe, status, t_reached, nullptr, nullptr);
}

deq_status deq_solve_host_dense(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                                const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options* opts,
                                const int* devices, int ndevices, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                                uint32_t* nfe, uint8_t* status, double* t_reached, double* dense, uint32_t* nvalid) {
    if (!dense) return fail(DEQ_ERR_INVALID_ARG, "dense is NULL");
    return solve_host_impl(rhs, ntraj, y0, params, params_per_traj, tspan, ntspan, method, tableau_set, opts, devices, ndevices, yfinal,
                           accepted, rejected, nfe, status, t_reached, dense, nvalid);
}

// The dense-output flavour of deq_solve_host.  The trajectory range is cut into chunks (~128 MiB of rows each); chunk j goes
// to device j % ndevices (a block-cyclic deal, which also balances parameter-sorted sweeps); every device thread keeps three
// chunks in flight: while chunk i travels to the caller's buffer on the copy stream, chunk i + 1 is integrated on the
// compute stream.  For ensembles with dense rows the PCIe link, not the kernel, is the bound (config 3: 16 KB of rows per
// trajectory against ~12 ns of kernel time), so the measure of this path is how close it stays to a plain pinned copy.
static deq_status solve_host_dense_impl(const deq_rhs* rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                                       const double* tspan, size_t ntspan, int method, int tableau_set, const deq_options& o_in,
                                       const std::vector<int>& devs, int cur, double* yfinal, uint32_t* accepted, uint32_t* rejected,
                                       uint32_t* nfe, uint8_t* status, double* t_reached, double* dense, uint32_t* nvalid) {
    deql::TableauInfo info;
    int sk = 0, tb = -2;
    if (deq_status ms = resolve_method(method, tableau_set, &tb, &sk, &info)) return ms;
    if (ntspan == 0) return DEQ_OK;   // nothing to solve (problem.rs:211-214)
    deq_options o = o_in;
    o.async = 1;
    // no per-chunk coherence probe / sorted dispatch (a host round trip on the compute stream): the kernel time hides behind the
    // copies anyway.  SoA rows: the row-synchronous kernel; trajectory-major: plain lane refill (that layout has no tail compaction,
    // so refill = 1 is what refill = 2 runs once the probe is out of the way)
    if (o.refill == 2) o.refill = o.dense_layout == DEQ_LAYOUT_SOA ? 0 : 1;
    const size_t nd = devs.size(), n = (size_t)rhs->n, np = (size_t)rhs->np;
    const bool tm = o.dense_layout == DEQ_LAYOUT_TRAJ_MAJOR && info.adaptive && !sk;
    const bool counts = info.adaptive || sk;
    const bool per = params && params_per_traj;
    const size_t row_bytes = n * ntspan * sizeof(double);   // dense bytes of one trajectory
    size_t chunk = (size_t)(128u << 20) / (row_bytes ? row_bytes : 1);
    if (const char* e = getenv("DEQ_HOST_CHUNK")) chunk = (size_t)strtoull(e, nullptr, 10);
    chunk = std::max<size_t>(chunk / 1024 * 1024, 1024);
    const size_t nchunks = (ntraj + chunk - 1) / chunk;
    std::vector<deq_status> rc(nd, DEQ_OK);
    std::vector<std::string> msg(nd);
    auto work = [&](size_t k) {
        if (cudaSetDevice(devs[k]) != cudaSuccess) { rc[k] = DEQ_ERR_CUDA; msg[k] = "cudaSetDevice failed"; cudaGetLastError(); return; }
        cudaStream_t sc = nullptr, sx = nullptr;
        const cudaStream_t caller_stream = g_stream;
        struct Slot { deq_ensemble* ens = nullptr; deq_result* res = nullptr; double* d_aos = nullptr; cudaEvent_t computed = nullptr, copied = nullptr; bool busy = false; } slot[3];   // three chunks in flight: computed / travelling / being set up
        deq_status s = DEQ_OK;
        auto ck = [&](cudaError_t e_) { if (e_ != cudaSuccess && !s) { cudaGetLastError(); s = fail(DEQ_ERR_CUDA, cudaGetErrorString(e_)); } };
        ck(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
        ck(cudaStreamCreateWithFlags(&sx, cudaStreamNonBlocking));
        for (auto& sl : slot) { ck(cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming)); ck(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming)); }
        g_stream = sc;
        auto retire = [&](Slot& sl) {
            if (!sl.busy) return;
            ck(cudaEventSynchronize(sl.copied));
            dfree(sl.d_aos, sx); sl.d_aos = nullptr;
            if (sl.res) deq_result_destroy(sl.res);
            if (sl.ens) deq_ensemble_destroy(sl.ens);
            sl.res = nullptr; sl.ens = nullptr; sl.busy = false;
        };
        size_t it = 0;
        for (size_t j = k; j < nchunks && !s; j += nd, ++it) {
            Slot& sl = slot[it % 3];
            retire(sl);
            const size_t b = j * chunk, cnt = std::min(chunk, ntraj - b);
            s = deq_ensemble_create(rhs, cnt, y0 + b * n, per ? params + b * np : params, params_per_traj, tspan, ntspan, &sl.ens);
            if (s) break;
            s = deq_solve(sl.ens, method, tableau_set, &o, &sl.res);
            if (s) break;
            sl.busy = true;
            deq_result* r = sl.res;
            ck(cudaEventRecord(sl.computed, sc));
            ck(cudaStreamWaitEvent(sx, sl.computed, 0));
            // rows
            if (tm) ck(cudaMemcpyAsync(dense + b * n * ntspan, r->d_dense, cnt * row_bytes, cudaMemcpyDeviceToHost, sx));
            else ck(cudaMemcpy2DAsync(dense + b, ntraj * sizeof(double), r->d_dense, cnt * sizeof(double), cnt * sizeof(double), n * ntspan, cudaMemcpyDeviceToHost, sx));
            if (nvalid) {
                if (r->d_nvalid) ck(cudaMemcpyAsync(nvalid + b, r->d_nvalid, cnt * 4, cudaMemcpyDeviceToHost, sx));
                else for (size_t i = 0; i < cnt; ++i) nvalid[b + i] = (uint32_t)ntspan;   // fixed-step methods write every row
            }
            // final states ([n][cnt] on the device -> [cnt][n]) and counters
            ck(dalloc(&sl.d_aos, cnt * n, sx));
            if (!s) ck(deql::launch_transpose_to_aos(r->d_yfinal, sl.d_aos, cnt, (int)n, sx));
            ck(cudaMemcpyAsync(yfinal + b * n, sl.d_aos, cnt * n * sizeof(double), cudaMemcpyDeviceToHost, sx));
            if (counts && r->d_acc) {
                if (accepted) ck(cudaMemcpyAsync(accepted + b, r->d_acc, cnt * 4, cudaMemcpyDeviceToHost, sx));
                if (rejected) ck(cudaMemcpyAsync(rejected + b, r->d_rej, cnt * 4, cudaMemcpyDeviceToHost, sx));
                if (nfe) ck(cudaMemcpyAsync(nfe + b, r->d_nfe, cnt * 4, cudaMemcpyDeviceToHost, sx));
                if (status) ck(cudaMemcpyAsync(status + b, r->d_status, cnt, cudaMemcpyDeviceToHost, sx));
                if (t_reached) ck(cudaMemcpyAsync(t_reached + b, r->d_treached, cnt * 8, cudaMemcpyDeviceToHost, sx));
            }
            ck(cudaEventRecord(sl.copied, sx));
        }
        for (auto& sl : slot) retire(sl);
        if (sc) { cudaStreamSynchronize(sc); }
        g_stream = caller_stream;
        for (auto& sl : slot) { if (sl.computed) cudaEventDestroy(sl.computed); if (sl.copied) cudaEventDestroy(sl.copied); }
        if (sx) cudaStreamDestroy(sx);
        if (sc) cudaStreamDestroy(sc);
        if (s) { rc[k] = s; msg[k] = g_err; }
    };
    if (nd == 1 && devs[0] == cur) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t k = 0; k < nd; ++k) th.emplace_back(work, k);
        for (auto& t : th) t.join();
        cudaSetDevice(cur);
    }
    for (size_t k = 0; k < nd; ++k)
        if (rc[k]) return fail(rc[k], "device " + std::to_string(devs[k]) + ": " + msg[k]);
    return DEQ_OK;
}

deq_status deq_tableau_get(int method, int set, int* stages, double* a, double* b, double* c, int* adaptive,
                           int* order_min, int* fsal) {
    const int tb = tb_of(method, set);
    if (tb == -1) return fail(DEQ_ERR_UNSUPPORTED, "the stiff methods have no Butcher tableau (see deq_rosenbrock_get)");
    if (tb < 0) return fail(DEQ_ERR_INVALID_ARG, "unknown method");
    deql::TableauInfo info;
    g_tb[tb].info(&info);
    if (stages) *stages = info.S;
    if (adaptive) *adaptive = info.adaptive;
    if (order_min) *order_min = info.order_min;
    if (fsal) *fsal = info.fsal;
    if (a) std::memcpy(a, info.a, sizeof(double) * info.S * info.S);
    if (b) std::memcpy(b, info.b, sizeof(double) * info.S * 2);
    if (c) std::memcpy(c, info.c, sizeof(double) * info.S);
    return DEQ_OK;
}

deq_status deq_rosenbrock_get(int method, double* gamma, double* a, double* b, double* c) {
    const int sk = stiff_kind(method);
    if (sk < 2) return fail(DEQ_ERR_INVALID_ARG, "deq_rosenbrock_get takes DEQ_ODE4SKR or DEQ_ODE4SS");
    double g = 0., aa[16], bb[4], cc[16];
    deql::rosenbrock_info(sk - 2, &g, aa, bb, cc);
    if (gamma) *gamma = g;
    if (a) std::memcpy(a, aa, sizeof(aa));
    if (b) std::memcpy(b, bb, sizeof(bb));
    if (c) std::memcpy(c, cc, sizeof(cc));
    return DEQ_OK;
}

deq_status deq_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(DEQ_ERR_INVALID_ARG, "out is NULL");
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return DEQ_OK;
}
void deq_host_free(void* p) { if (p) cudaFreeHost(p); }

deq_status deq_test_pow(const double* x, const double* y, double* out, size_t n) {
    if (n == 0) return DEQ_OK;
    cudaStream_t st = g_stream;
    TempBuf<double> dx(st), dy(st), dout(st);
    CK(dalloc(&dx.p, n, st)); CK(dalloc(&dy.p, n, st)); CK(dalloc(&dout.p, n, st));
    CK(cudaMemcpyAsync(dx.p, x, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dy.p, y, n * 8, cudaMemcpyHostToDevice, st));
    CK(deql::launch_test_pow(dx.p, dy.p, dout.p, n, st));
    CK(cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DEQ_OK;
}

deq_status deq_test_div(uint64_t seed, uint64_t count, uint64_t* mismatches, uint64_t* checked) {
    if (!mismatches || !checked) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    cudaStream_t st = g_stream;
    TempBuf<unsigned long long> d(st);
    CK(dalloc(&d.p, 2, st));
    CK(cudaMemsetAsync(d.p, 0, 2 * sizeof(unsigned long long), st));
    CK(deql::launch_test_div(seed, count, d.p, st));
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *mismatches = h[0]; *checked = h[1];
    return DEQ_OK;
}

deq_status deq_test_log10(const double* x, double* out, size_t n) {
    if (n == 0) return DEQ_OK;
    cudaStream_t st = g_stream;
    TempBuf<double> dx(st), dout(st);
    CK(dalloc(&dx.p, n, st)); CK(dalloc(&dout.p, n, st));
    CK(cudaMemcpyAsync(dx.p, x, n * 8, cudaMemcpyHostToDevice, st));
    CK(deql::launch_test_log10(dx.p, dout.p, n, st));
    CK(cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DEQ_OK;
}

deq_status deq_measure_fp64_peak(double* tflops, double* sm_clock_mhz_hint) {
    if (!tflops) return fail(DEQ_ERR_INVALID_ARG, "NULL argument");
    cudaStream_t st = g_stream;
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const int blocks = sms * 8, iters = 4096;
    TempBuf<double> sink(st);
    CK(dalloc(&sink.p, (size_t)blocks * 256, st));
    struct Events { cudaEvent_t a = nullptr, b = nullptr; ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } ev;
    CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ev.a, st));
        CK(deql::launch_dfma_peak(sink.p, iters, blocks, st));
        CK(cudaEventRecord(ev.b, st));
        CK(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
        const double fl = (double)blocks * 256.0 * iters * 64.0 * 2.0;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    CK(cudaStreamSynchronize(st));
    *tflops = best;
    if (sm_clock_mhz_hint) *sm_clock_mhz_hint = khz / 1000.0;
    return DEQ_OK;
}

}  // extern "C"
