// diffeq_b200.hpp — header-only C++ mirror of the reference's problem/solver surface over the C ABI
// (include/diffeq_b200.h).  Names follow /root/reference/src/ode/: OdeProblem::builder().fun().init().tspan().build()
// (problem.rs:60-131), solve(Ode, OdeOptionMap) (problem.rs:133-147), OdeSolution (solution.rs:22-29), OdeError
// (error.rs:4-23).  The reference is compiled Rust; with no Rust toolchain in the build image this is the compiled-
// language host side.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "diffeq_b200.h"

namespace deq {

struct OdeError : std::runtime_error {
    deq_status status;
    OdeError(deq_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(deq_status s) { if (s != DEQ_OK) throw OdeError(s, deq_last_error_message()); }

enum class Ode { Feuler = 0, Heun, Midpoint, Ode23, Ode23s, Ode4, Ode45, Ode45fe, Ode4skr, Ode4ss, Ode78, Ode21 };

class Rhs {
public:
    static Rhs builtin(deq_system s) { deq_rhs* h = nullptr; check(deq_rhs_builtin(s, &h)); return Rhs(h); }
    static Rhs from_source(const std::string& src, int n, int np) { deq_rhs* h = nullptr; check(deq_rhs_from_source(src.c_str(), n, np, &h)); return Rhs(h); }
    Rhs(Rhs&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Rhs(const Rhs&) = delete;
    ~Rhs() { if (h_) deq_rhs_destroy(h_); }
    int dim() const { return deq_rhs_dim(h_); }
    const deq_rhs* get() const { return h_; }
private:
    explicit Rhs(deq_rhs* h) : h_(h) {}
    deq_rhs* h_;
};

struct OdeOptionMap {  // options.rs:8-25, defaults options.rs:283-299
    deq_options o = deq_options_default();
    OdeOptionMap& reltol(double v) { o.reltol = v; return *this; }
    OdeOptionMap& abstol(double v) { o.abstol = v; return *this; }
    OdeOptionMap& minstep(double v) { o.has_minstep = 1; o.minstep = v; return *this; }
    OdeOptionMap& maxstep(double v) { o.has_maxstep = 1; o.maxstep = v; return *this; }
    OdeOptionMap& initstep(double v) { o.initstep = v; return *this; }
    OdeOptionMap& mode(deq_mode m) { o.mode = m; return *this; }
    OdeOptionMap& points(deq_points p) { o.points = p; return *this; }          // options.rs:120-137
    OdeOptionMap& max_steps(uint64_t n) { o.max_steps = n; return *this; }      // extension (guard)
    OdeOptionMap& libm(deq_libm l) { o.libm = l; return *this; }
    OdeOptionMap& shard_chunk(uint64_t c) { o.shard_chunk = c; return *this; }
};

struct OdeSolution {  // solution.rs:22-29 (+ the diagnostics the reference drops)
    std::vector<double> tout;
    std::vector<std::vector<double>> yout;  // one final state per trajectory
    std::vector<uint32_t> accepted, rejected;
};

class OdeProblem;
class OdeBuilder {  // problem.rs:49-119
public:
    OdeBuilder& fun(const Rhs& f) { f_ = &f; return *this; }
    OdeBuilder& params(std::vector<double> p) { params_ = std::move(p); return *this; }
    OdeBuilder& init(std::vector<std::vector<double>> y0) { y0_ = std::move(y0); has_y0_ = true; return *this; }
    OdeBuilder& tspan(std::vector<double> t) { tspan_ = std::move(t); has_t_ = true; return *this; }
    OdeBuilder& tspan_linspace(double a, double b, size_t n) { tspan_.assign(n, 0.0); check(deq_tspan_linspace(a, b, n, tspan_.data())); has_t_ = true; return *this; }
    OdeProblem build();
private:
    const Rhs* f_ = nullptr; std::vector<double> params_; std::vector<std::vector<double>> y0_; std::vector<double> tspan_;
    bool has_y0_ = false, has_t_ = false;
};

class OdeProblem {  // problem.rs:32-47
public:
    static OdeBuilder builder() { return OdeBuilder(); }
    OdeProblem(OdeProblem&& o) noexcept : h_(o.h_), n_(o.n_), ntraj_(o.ntraj_), tspan_(std::move(o.tspan_)) { o.h_ = nullptr; }
    ~OdeProblem() { if (h_) deq_ensemble_destroy(h_); }
    OdeSolution solve(Ode ode, const OdeOptionMap& opts = OdeOptionMap(), deq_tableau_set set = DEQ_TABLEAU_REFERENCE) {  // problem.rs:133-147
        deq_result* r = nullptr;
        check(deq_solve(h_, (int)ode, set, &opts.o, &r));
        OdeSolution s;
        std::vector<double> flat(ntraj_ * n_);
        deq_status a = deq_result_final(r, flat.data());
        int adaptive = 0;
        deq_tableau_get((int)ode, set, nullptr, nullptr, nullptr, nullptr, &adaptive, nullptr, nullptr);
        deq_status b = DEQ_OK;
        if (adaptive || ode == Ode::Ode23s || ode == Ode::Ode4skr || ode == Ode::Ode4ss) { s.accepted.resize(ntraj_); s.rejected.resize(ntraj_); b = deq_result_counts(r, s.accepted.data(), s.rejected.data(), nullptr, nullptr, nullptr); }
        deq_result_destroy(r);
        check(a); check(b);
        for (size_t i = 0; i < ntraj_; ++i) s.yout.emplace_back(flat.begin() + i * n_, flat.begin() + (i + 1) * n_);
        if (!tspan_.empty()) s.tout.push_back(tspan_.back());
        return s;
    }
    // The reference-shaped result of every trajectory: the complete (tout, yout) that `OdeProblem::solve` returns for that
    // problem (problem.rs:354-357 adaptive Points::All / Specified stream, :401-405 fixed grid, :556 ode23s, :639 oderosenbrock
    // with its `tout = diff(tspan)`).  One OdeSolution per trajectory; accepted / rejected hold that trajectory's counters.
    std::vector<OdeSolution> solve_trajectories(Ode ode, OdeOptionMap opts = OdeOptionMap(), deq_tableau_set set = DEQ_TABLEAU_REFERENCE) {
        const bool stiff = ode == Ode::Ode23s || ode == Ode::Ode4skr || ode == Ode::Ode4ss;
        int adaptive = ode == Ode::Ode23s;
        if (!stiff) check(deq_tableau_get((int)ode, set, nullptr, nullptr, nullptr, nullptr, &adaptive, nullptr, nullptr));
        opts.o.output = adaptive ? DEQ_OUT_ALL : DEQ_OUT_TSPAN;
        deq_result* r = nullptr;
        check(deq_solve(h_, (int)ode, set, &opts.o, &r));
        std::vector<OdeSolution> out(ntraj_);
        deq_status st = DEQ_OK;
        const size_t nt = tspan_.size();
        if (nt == 0) { deq_result_destroy(r); return out; }   // nothing to solve: OdeSolution::default()
        if (adaptive) {
            std::vector<uint64_t> off(ntraj_ + 1);
            st = deq_result_all_offsets(r, off.data());
            std::vector<double> t(off[ntraj_]), y(off[ntraj_] * n_);
            std::vector<uint32_t> acc(ntraj_), rej(ntraj_);
            if (!st) st = deq_result_all_copy(r, 0, ntraj_, t.data(), y.data());
            if (!st) st = deq_result_counts(r, acc.data(), rej.data(), nullptr, nullptr, nullptr);
            for (size_t i = 0; !st && i < ntraj_; ++i) {
                out[i].tout.assign(t.begin() + off[i], t.begin() + off[i + 1]);
                for (uint64_t k = off[i]; k < off[i + 1]; ++k) out[i].yout.emplace_back(y.begin() + k * n_, y.begin() + (k + 1) * n_);
                out[i].accepted.assign(1, acc[i]); out[i].rejected.assign(1, rej[i]);
            }
        } else {
            std::vector<double> rows(nt * n_);
            std::vector<uint32_t> nv(ntraj_, (uint32_t)nt), nvtmp(1);
            for (size_t i = 0; !st && i < ntraj_; ++i) {
                st = deq_result_trajectory(r, i, rows.data());
                if (!st && stiff) { std::vector<double> tmp((size_t)n_ * nt); st = deq_result_dense(r, i, i + 1, tmp.data(), nvtmp.data()); nv[i] = nvtmp[0]; }
                for (size_t k = 0; !st && k < nv[i]; ++k) out[i].yout.emplace_back(rows.begin() + k * n_, rows.begin() + (k + 1) * n_);
                if (stiff) for (size_t k = 0; k + 1 < nt; ++k) out[i].tout.push_back(tspan_[k + 1] - tspan_[k]);   // diff(tspan), sic
                else out[i].tout = tspan_;
            }
        }
        deq_result_destroy(r);
        check(st);
        return out;
    }
    OdeSolution ode45(const OdeOptionMap& o = OdeOptionMap()) { return solve(Ode::Ode45, o); }
    OdeSolution ode45_fe(const OdeOptionMap& o = OdeOptionMap()) { return solve(Ode::Ode45fe, o); }
    OdeSolution ode78(const OdeOptionMap& o = OdeOptionMap()) { return solve(Ode::Ode78, o); }
    OdeSolution ode4() { return solve(Ode::Ode4); }
    OdeSolution ode23s(const OdeOptionMap& o = OdeOptionMap()) { return solve(Ode::Ode23s, o); }   // problem.rs:409
    OdeSolution ode4s_kr() { return solve(Ode::Ode4skr); }                                        // problem.rs:643
    OdeSolution ode4s_s() { return solve(Ode::Ode4ss); }                                          // problem.rs:648
private:
    friend class OdeBuilder;
    OdeProblem(deq_ensemble* h, size_t n, size_t ntraj, std::vector<double> t) : h_(h), n_(n), ntraj_(ntraj), tspan_(std::move(t)) {}
    deq_ensemble* h_; size_t n_, ntraj_; std::vector<double> tspan_;
};

inline OdeProblem OdeBuilder::build() {  // problem.rs:92-104
    if (!f_) throw OdeError(DEQ_ERR_UNINITIALIZED, "Required problem must be initialized");
    if (!has_y0_) throw OdeError(DEQ_ERR_UNINITIALIZED, "Initial starting point must be initialized");
    if (!has_t_) throw OdeError(DEQ_ERR_UNINITIALIZED, "Time span must be initialized");
    // the C ABI reads ntraj * n state and nparams (or ntraj * nparams) parameter doubles through raw pointers: check the lengths here
    const size_t n = (size_t)f_->dim(), np = (size_t)deq_rhs_nparams(f_->get()), ntraj = y0_.size();
    std::vector<double> flat;
    flat.reserve(ntraj * n);
    for (size_t i = 0; i < ntraj; ++i) {
        if (y0_[i].size() != n) throw OdeError(DEQ_ERR_INVALID_ARG, "initial state " + std::to_string(i) + " has " + std::to_string(y0_[i].size()) + " components, the right-hand side expects " + std::to_string(n));
        flat.insert(flat.end(), y0_[i].begin(), y0_[i].end());
    }
    int per_traj = 0;
    if (np == 0) { if (!params_.empty()) throw OdeError(DEQ_ERR_INVALID_ARG, "the right-hand side takes no parameters"); }
    else if (params_.empty()) throw OdeError(DEQ_ERR_UNINITIALIZED, "Required problem parameters must be initialized");
    else if (params_.size() == np) per_traj = 0;
    else if (params_.size() == np * ntraj) per_traj = 1;   // one row per trajectory, [ntraj][nparams]
    else throw OdeError(DEQ_ERR_INVALID_ARG, std::to_string(params_.size()) + " parameters given, expected " + std::to_string(np) + " or " + std::to_string(np * ntraj));
    deq_ensemble* h = nullptr;
    check(deq_ensemble_create(f_->get(), ntraj, flat.empty() ? nullptr : flat.data(), params_.empty() ? nullptr : params_.data(), per_traj,
                              tspan_.data(), tspan_.size(), &h));
    return OdeProblem(h, (size_t)f_->dim(), y0_.size(), tspan_);
}

}  // namespace deq
