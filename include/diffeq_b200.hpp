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
    std::vector<double> flat;
    for (auto& y : y0_) flat.insert(flat.end(), y.begin(), y.end());
    deq_ensemble* h = nullptr;
    check(deq_ensemble_create(f_->get(), y0_.size(), flat.data(), params_.empty() ? nullptr : params_.data(), 0, tspan_.data(), tspan_.size(), &h));
    return OdeProblem(h, (size_t)f_->dim(), y0_.size(), tspan_);
}

}  // namespace deq
