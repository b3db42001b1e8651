// diffeq_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain C++ restatement of the Runge–Kutta hot path of mattsse/diffeq, written to follow the
// reference operation-for-operation so that its results are what the Rust crate would produce on
// the same inputs (IEEE f64, no FMA contraction: build with -ffp-contract=off, see Makefile).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (libdiffeq_b200.so) never links, includes or calls it.
//
// PARITY UNPINNED BY THE REFERENCE: the reference's own tests assert no numeric solver result
//  and no Rust toolchain exists in this image, so this restatement cannot be
// checked against reference outputs.  It is pinned instead by (i) an independent pure-Python
// restatement (tests/pyref.py) that must agree bit-for-bit, and (ii) golden vectors generated from
// it and committed under tests/golden/.
//
// Reference anchors (all under /root/reference/src/ode/):
//   oderk_fixed        problem.rs:361-406      -> solve_fixed
//   oderk_adapt        problem.rs:193-358      -> solve_adapt
//   calc_coefficients  problem.rs:702-737      -> stages
//   embedded_step      problem.rs:740-780      -> embedded
//   hermite_interp     problem.rs:783-798      -> hermite
//   stepsize_hw92      problem.rs:801-845      -> stepsize_hw92  (+ pnorm types.rs:91-116)
//   hinit              problem.rs:850-899      -> hinit
//   tableaus           runge_kutta.rs:231-817  -> make_tableau (literal expressions, nalgebra
//                                                 row-major constructor semantics, defects kept)
//   option defaults    options.rs:283-299, problem.rs:219-225
// Third-party arithmetic the reference reaches: libm pow/log10 (f64::powf/log10 on linux-gnu call
// the system glibc; 2.39 here) — called directly below, as Rust would.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int MAXS = 13;
constexpr int MAXN = 64;

// ------------------------------------------------------------------------------------------
// Tableaus (runge_kutta.rs:231-817).  `set` 0 = REFERENCE (bug-for-bug), 1 = TEXTBOOK.
// ------------------------------------------------------------------------------------------
struct Tableau {
    int S = 0;
    double a[MAXS][MAXS] = {};
    double b[MAXS][2] = {};  // b[s][0] stepping weights, b[s][1] error-comparison weights
    double c[MAXS] = {};
    bool adaptive = false;
    int order_min = 0;  // RKOrder::min(), runge_kutta.rs:20-55
    bool fsal = false;  // is_first_same_as_last, runge_kutta.rs:185-193
};

// nalgebra `MatrixN::new(..)` / `from_row_slice` take elements row-major.
void fill_a(Tableau& t, const std::vector<double>& rowmajor) {
    for (int r = 0; r < t.S; ++r)
        for (int c = 0; c < t.S; ++c) t.a[r][c] = rowmajor[r * t.S + c];
}
void fill_b2(Tableau& t, const std::vector<double>& rowmajor) {  // S x 2, row-major
    for (int s = 0; s < t.S; ++s) {
        t.b[s][0] = rowmajor[2 * s];
        t.b[s][1] = rowmajor[2 * s + 1];
    }
    t.adaptive = true;
}
void fill_b1(Tableau& t, const std::vector<double>& v) {
    for (int s = 0; s < t.S; ++s) t.b[s][0] = v[s];
    t.adaptive = false;
}
void fill_c(Tableau& t, const std::vector<double>& v) {
    for (int s = 0; s < t.S; ++s) t.c[s] = v[s];
}
// runge_kutta.rs:185-193: compares the last row of `a` with the first S entries of b.as_slice()
// (column-major storage => column 0 for adaptive weights) and requires c[S-1] == 1.
bool compute_fsal(const Tableau& t) {
    int row = t.S - 1;
    for (int col = 0; col < t.S; ++col)
        if (t.a[row][col] != t.b[col][0]) return false;
    return t.c[row] == 1.0;
}

// Method ids follow `Ode` (mod.rs:14-26) + ODE21 extension for `.ode21()` (problem.rs:164-166).
enum Method { FEULER = 0, HEUN, MIDPOINT, ODE23, ODE23S, ODE4, ODE45, ODE45FE, ODE4SKR, ODE4SS, ODE78, ODE21 };

bool make_tableau(int method, int set, Tableau& t) {
    const bool tb = (set == 1);
    t = Tableau();
    switch (method) {
        case FEULER:  // :238-249
            t.S = 1; fill_a(t, {0.}); fill_b1(t, {1.}); fill_c(t, {0.}); t.order_min = 1; break;
        case MIDPOINT:  // :261-272
            t.S = 2; fill_a(t, {0., 0., 0.5, 0.0}); fill_b1(t, {0., 1.0}); fill_c(t, {0., 0.5}); t.order_min = 2; break;
        case HEUN:  // :281-292
            t.S = 2; fill_a(t, {0., 0., 1., 0.}); fill_b1(t, {0.5, 0.5}); fill_c(t, {0., 1.}); t.order_min = 2; break;
        case ODE21:  // rk21 :294-305 — reference weights are transposed
            t.S = 2; fill_a(t, {0., 0., 1., 0.});
            if (!tb) fill_b2(t, {0.5, 0.5, 1., 0.});
            else     fill_b2(t, {0.5, 1., 0.5, 0.});
            fill_c(t, {0., 1.}); t.order_min = 1; break;
        case ODE23:  // rk23 :319-357 — reference col1[0] is 1/9 (textbook 2/9)
            t.S = 4;
            fill_a(t, {0., 0., 0., 0., 0.5, 0., 0., 0., 0., 0.75, 0., 0., 2. / 9., 1. / 3., 4. / 9., 0.});
            fill_b2(t, {7. / 24., tb ? 2. / 9. : 1. / 9., 0.25, 1. / 3., 1. / 3., 4. / 9., 0.125, 0.});
            fill_c(t, {0., 0.5, 0.75, 1.}); t.order_min = 2; break;
        case ODE4:  // rk4 :370-383
            t.S = 4;
            fill_a(t, {0., 0., 0., 0., 0.5, 0., 0., 0., 0., 0.5, 0., 0., 0., 0., 1., 0.});
            fill_b1(t, {1. / 6., 1. / 3., 1. / 3., 1. / 6.}); fill_c(t, {0., 0.5, 0.5, 1.}); t.order_min = 4; break;
        case ODE45FE:  // rk45 (Fehlberg) :401-463
            t.S = 6;
            fill_a(t, {0., 0., 0., 0., 0., 0.,
                       0.25, 0., 0., 0., 0., 0.,
                       0.09375, 0.28125, 0., 0., 0., 0.,
                       1932. / 2197., -7200. / 2197., 7296. / 2197., 0., 0., 0.,
                       439. / 216., -8., 3680. / 513., -845. / 4104., 0., 0.,
                       -8. / 27., 2., -3544. / 2565., 1859. / 4104., -0.275, 0.});
            fill_b2(t, {25. / 216., 16. / 135., 0., 0., 1408. / 2565., 6656. / 12825., 2197. / 4104.,
                        28561. / 56430., -0.2, -0.18, 0., 2. / 55.});
            fill_c(t, {0., 0.25, 0.375, 12. / 13., 1., 0.5}); t.order_min = 4; break;
        case ODE45:  // dopri5 :478-562 — reference c[3] = 0.75 (textbook 4/5)
            t.S = 7;
            fill_a(t, {0., 0., 0., 0., 0., 0., 0.,
                       0.2, 0., 0., 0., 0., 0., 0.,
                       0.075, 0.225, 0., 0., 0., 0., 0.,
                       44. / 45., -56. / 15., 32. / 9., 0., 0., 0., 0.,
                       19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729., 0., 0., 0.,
                       9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656., 0., 0.,
                       35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 0.});
            fill_b2(t, {35. / 384., 5179. / 57600., 0., 0., 500. / 1113., 7571. / 16695., 125. / 192.,
                        393. / 640., -2187. / 6784., -92097. / 339200., 11. / 84., 187. / 2100., 0., 0.025});
            fill_c(t, {0., 0.2, 0.3, tb ? 4. / 5. : 0.75, 8. / 9., 1., 1.}); t.order_min = 4; break;
        case ODE78: {  // feh78 :582-816 — reference a-row 5 and b are mis-transcribed
            t.S = 13;
            std::vector<double> a = {
                0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.,
                2. / 27., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.,
                1. / 36., 1. / 12., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.,
                1. / 24., 0., 0.125, 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.,
                5. / 12., 0., -25. / 16., 25. / 16., 0., 0., 0., 0., 0., 0., 0., 0., 0.,
                0.5, 0., 0., 0.75, 0.2, 0., 0., 0., 0., 0., 0., 0., 0.,
                -25. / 108., 0., 0., 125. / 108., -65. / 27., 125. / 54., 0., 0., 0., 0., 0., 0., 0.,
                31. / 300., 0., 0., 0., 61. / 225., -2. / 9., 13. / 900., 0., 0., 0., 0., 0., 0.,
                2., 0., 0., -53. / 6., 704. / 45., -107. / 9., 67. / 90., 3., 0., 0., 0., 0., 0.,
                -91. / 108., 0., 0., 23. / 108., -976. / 135., 311. / 54., -19. / 60., 17. / 6., -1. / 12., 0., 0., 0., 0.,
                2383. / 4100., 0., 0., -341. / 164., 4496. / 1025., -301. / 82., 2133. / 4100., 45. / 82., 45. / 164., 18. / 41., 0., 0., 0.,
                3. / 205., 0., 0., 0., 0., -6. / 41., -3. / 205., -3. / 41., 3. / 41., 6. / 41., 0., 0., 0.,
                -1777. / 4100., 0., 0., -341. / 164., 4496. / 1025., -289. / 82., 2193. / 4100., 51. / 82., 33. / 164., 12. / 41., 0., 1., 0.};
            if (tb) { a[5 * 13 + 0] = 1. / 20.; a[5 * 13 + 3] = 1. / 4.; a[5 * 13 + 4] = 1. / 5.; }
            fill_a(t, a);
            if (!tb)
                fill_b2(t, {41. / 840., 9. / 35., 0., 9. / 35., 0., 9. / 35., 0., 9. / 280., 0., 9. / 280., 0.,
                            9. / 280., 0., 9. / 280., 0., 41. / 840., 0., 0., 0., 0., 34. / 105., 41. / 840.,
                            34. / 105., 0., 9. / 35., 41. / 840.});
            else {
                const double b7[13] = {41. / 840., 0., 0., 0., 0., 34. / 105., 9. / 35., 9. / 35., 9. / 280., 9. / 280., 41. / 840., 0., 0.};
                const double b8[13] = {0., 0., 0., 0., 0., 34. / 105., 9. / 35., 9. / 35., 9. / 280., 9. / 280., 0., 41. / 840., 41. / 840.};
                std::vector<double> bb;
                for (int s = 0; s < 13; ++s) { bb.push_back(b7[s]); bb.push_back(b8[s]); }
                fill_b2(t, bb);
            }
            fill_c(t, {0., 2. / 27., 1. / 9., 1. / 6., 5. / 12., 0.5, 5. / 6., 1. / 6., 2. / 3., 1. / 3., 1., 0., 1.});
            t.order_min = 7; break;
        }
        default: return false;
    }
    t.fsal = compute_fsal(t);
    return true;
}

// ------------------------------------------------------------------------------------------
// Built-in right-hand sides (the reference takes a closure F: Fn(f64,&Y)->Y, problem.rs:32-36).
// ------------------------------------------------------------------------------------------
enum Rhs { LORENZ = 0, VANDERPOL = 1, PENDULUM = 2, RING8 = 3, RING16 = 4, RING32 = 5, RING64 = 6, CVDP2 = 7, CVDP3 = 8, CVDP6 = 9, CVDP12 = 10 };

struct LorenzF {  // problem.rs:989-1000: params = {sigma, rho, beta}
    static constexpr int N = 3;
    static inline void eval(double, const double* v, double* d, const double* p) {
        double x = v[0], y = v[1], z = v[2];
        d[0] = p[0] * (y - x);
        d[1] = x * (p[1] - z) - y;
        d[2] = x * y - p[2] * z;
    }
};
// Test systems of larger dimension (not in the reference; they pin the generic-dimension NVRTC path of the product, whose user
// source states the same expression): a ring of N logistic cells with diffusive coupling, params = {kappa, r},
//   d_i = kappa * (y_{i+1} - y_i) + r * (y_i * (1 - y_{i-1}))   (indices mod N)
template <int NN>
struct RingF {
    static constexpr int N = NN;
    static inline void eval(double, const double* v, double* d, const double* p) {
        for (int i = 0; i < NN; ++i) {
            const int ip = (i + 1) % NN, im = (i + NN - 1) % NN;
            d[i] = p[0] * (v[ip] - v[i]) + p[1] * (v[i] * (1.0 - v[im]));
        }
    }
};
// Test systems for the stiff path beyond n = 3 (not in the reference; the product's user source states the same expression):
// M Van der Pol oscillators in a ring, state (x_0, y_0, x_1, y_1, ...), params = {mu, kappa},
//   x_i' = y_i,   y_i' = mu * (1 - x_i^2) * y_i - x_i + kappa * (x_{i+1} - x_i)
template <int M>
struct CoupledVdpF {
    static constexpr int N = 2 * M;
    static inline void eval(double, const double* v, double* d, const double* p) {
        for (int i = 0; i < M; ++i) {
            const double x = v[2 * i], y = v[2 * i + 1], xn = v[2 * ((i + 1) % M)];
            d[2 * i] = y;
            d[2 * i + 1] = p[0] * (1.0 - x * x) * y - x + p[1] * (xn - x);
        }
    }
};
struct VdpF {  // x' = y, y' = mu (1 - x^2) y - x ; params = {mu}
    static constexpr int N = 2;
    static inline void eval(double, const double* v, double* d, const double* p) {
        double x = v[0], y = v[1];
        d[0] = y;
        d[1] = p[0] * (1.0 - x * x) * y - x;
    }
};
struct PendulumF {  // th' = w, w' = -gamma w - sin th + A cos(Omega t) ; params = {gamma, A, Omega}
    static constexpr int N = 2;
    static inline void eval(double t, const double* v, double* d, const double* p) {
        double th = v[0], w = v[1];
        d[0] = w;
        d[1] = -p[0] * w - std::sin(th) + p[1] * std::cos(p[2] * t);
    }
};

struct Opts {
    double reltol, abstol;
    int has_minstep; double minstep;
    int has_maxstep; double maxstep;
    double initstep;
    int points;  // 0 = All, 1 = Specified (reference semantics, including its lagging-y defect)
    uint64_t max_steps;  // oracle guard (the reference has none); 0 = unlimited
    // 0: libm calls as written in the source — pow(s, 0.5) (types.rs:115) and pow(10, p) (problem.rs:893).
    // 1: what an optimised rustc/LLVM build evaluates instead: LLVM's SimplifyLibCalls rewrites pow(x, 0.5) with a
    //    constant exponent to sqrt(x) (the exponent 1/p folds to 0.5 once pnorm(PNorm::default()) is inlined) and
    //    pow(10.0, x) to exp10(x).  Which of the two a given build of the crate runs cannot be established without a
    //    Rust toolchain; both are provided and pinned.
    int libm_folded;
};

enum Status { ST_DONE = 0, ST_MINSTEP_BREAK = 1, ST_MAX_STEPS = 2 };
enum Err { OK = 0, ERR_UNINITIALIZED = 1, ERR_WEIGHT_TYPE = 2, ERR_NAN = 3, ERR_ZERO_TIMESPAN = 4, ERR_INVALID_INITSTEP = 5, ERR_INVALID_MATRIX = 6, ERR_UNSUPPORTED = 9, ERR_INVALID_ARG = 10 };

struct Counters { uint32_t accepted = 0, rejected = 0, nfe = 0; uint8_t status = ST_DONE; double t_reached = 0; };

inline double signum(double x) { return std::isnan(x) ? x : std::copysign(1.0, x); }  // f64::signum
inline double fmax_rs(double a, double b) { return std::fmax(a, b); }  // f64::max (NaN-ignoring)
inline double fmin_rs(double a, double b) { return std::fmin(a, b); }

// calc_coefficients, problem.rs:702-737.  k[0] must already hold k0; fills k[1..S-1].
template <int N, class F>
inline void stages(const Tableau& tb, double t, const double* y, double dt, double k[][MAXN], const double* p) {
    for (int row = 1; row < tb.S; ++row) {
        double yi[N];
        for (int d = 0; d < N; ++d) yi[d] = y[d];
        for (int col = 0; col < row; ++col) {
            const double w = tb.a[row][col] * dt;  // (btab.a[(row,col)] * dt), zeros included
            for (int d = 0; d < N; ++d) yi[d] += k[col][d] * w;
        }
        const double tn = t + tb.c[row] * dt;
        F::eval(tn, yi, k[row], p);
    }
}

// oderk_fixed, problem.rs:361-406.  yout is [nt][N] (AoS, like Vec<Y>); tout == tspan.
template <int N, class F>
void solve_fixed(const Tableau& tb, const double* y0, const double* tspan, size_t nt, const double* p,
                 double* yout, double* yfinal, uint32_t* nfe) {
    double y[N];
    for (int d = 0; d < N; ++d) y[d] = y0[d];
    if (yout) for (int d = 0; d < N; ++d) yout[d] = y[d];
    uint32_t ev = 0;
    for (size_t i = 0; i + 1 < nt; ++i) {
        const double dt = tspan[i + 1] - tspan[i];
        double k[MAXS][MAXN];
        F::eval(tspan[i], y, k[0], p);
        stages<N, F>(tb, tspan[i], y, dt, k, p);
        ev += tb.S;
        for (int s = 0; s < tb.S; ++s)
            for (int d = 0; d < N; ++d) y[d] += k[s][d] * tb.b[s][0] * dt;  // (k*b[s])*dt
        if (yout) for (int d = 0; d < N; ++d) yout[(i + 1) * N + d] = y[d];
    }
    if (yfinal) for (int d = 0; d < N; ++d) yfinal[d] = y[d];
    if (nfe) *nfe = ev;
}

// hinit, problem.rs:850-899
template <int N, class F>
inline void hinit(const double* x0, double t0, double tend, int order, double reltol, double abstol,
                  const double* p, double& h, double& tdir, double* f0, int libm_folded = 0) {
    tdir = signum(tend - t0);
    double norm = 0.0;
    for (int d = 0; d < N; ++d) { double a = std::fabs(x0[d]); if (a > norm) norm = a; }
    const double tau = fmax_rs(norm * reltol, 1.0 * abstol);
    const double d0 = norm / tau;
    F::eval(t0, x0, f0, p);
    double nf0 = 0.0;
    for (int d = 0; d < N; ++d) { double a = std::fabs(f0[d]); if (a > nf0) nf0 = a; }
    const double d1 = nf0 / tau;
    const double h0 = (d0 < 1.0 * 1e-5 || d1 < 1.0 * 1e-5) ? 1.0e-6 : 0.01 * (d0 / d1);
    double x1[N];
    for (int d = 0; d < N; ++d) x1[d] = x0[d] + f0[d] * h0 * tdir;
    double f1[N];
    F::eval(t0 + tdir * h0, x1, f1, p);
    for (int d = 0; d < N; ++d) f1[d] -= f0[d];
    double nf1 = 0.0;
    for (int d = 0; d < N; ++d) { double a = std::fabs(f1[d]); if (a > nf1) nf1 = a; }
    const double d2 = nf1 / (tau * h0);
    double h1;
    const double m = fmax_rs(d1, d2);
    if (m < 1.0 * 1e-15) {
        h1 = fmax_rs(1.0e-6, 1.0e-3 * h0);
    } else {
        const double pw = -(2. + std::log10(m)) / (double)(order + 1);
        h1 = libm_folded ? exp10(pw) : std::pow(10.0, pw);
    }
    h = tdir * fmin_rs(fmin_rs(h1, 100. * h0), tdir * (tend - t0));
}

// stepsize_hw92, problem.rs:801-845 (+ pnorm P(2), types.rs:109-115)
template <int N>
inline void stepsize_hw92(double dt, double tdir, const double* x0, const double* xtrial, double* xerr, int order,
                          unsigned timeout, double abstol, double reltol, double maxstep, double& err, double& newdt,
                          unsigned& timeout_out, int libm_folded = 0) {
    const double fac = 0.8, facmin = 0.2;
    for (int d = 0; d < N; ++d) {
        if (std::isnan(xtrial[d])) { err = 10.; newdt = facmin * dt; timeout_out = 5; return; }
        xerr[d] /= fmax_rs(std::fabs(x0[d]), std::fabs(xtrial[d])) * reltol + abstol;
    }
    double s = 0.0;
    for (int d = 0; d < N; ++d) { const double a = std::fabs(xerr[d]); s = s + a * a; }
    err = libm_folded ? std::sqrt(s) : std::pow(s, 1.0 * (1. / 2.0));
    const double pw = 1. / (double)(order + 1);
    double nd = fmin_rs(maxstep, fmax_rs(facmin, std::pow(1.0 / err, pw) * fac) * tdir * dt);
    if (timeout > 0) { nd = fmin_rs(nd, dt); timeout -= 1; }
    newdt = tdir * nd;
    timeout_out = timeout;
}

// hermite_interp, problem.rs:783-798
template <int N>
inline void hermite(double tq, double t, double dt, const double* y0, const double* y1, const double* f0,
                    const double* f1, double* out) {
    const double theta = (tq - t) / dt;
    for (int i = 0; i < N; ++i) {
        out[i] = (y0[i] * (1. - theta) + y1[i] * theta) +
                 ((y1[i] - y0[i]) * (1. - 2. * theta) + f0[i] * (theta - 1.) * dt + f1[i] * theta * dt) * theta * (theta - 1.);
    }
}

// Output sink: receives every (t, y) the reference would push into (tout, yout).  `idx` is the tspan index
// for rows produced by hermite_interp and -1 for y0 / step end points.
struct Sink {
    double* tout = nullptr; double* yout = nullptr; size_t cap = 0; size_t n = 0; int N = 0; bool overflow = false;
    // dense (tspan-filtered, SoA [N][nt][ntraj]) target used by oracle_ensemble_dense
    double* dense = nullptr; size_t nt = 0, ntraj = 0, traj = 0; size_t ndense = 1;
    inline void push(double t, const double* y, long long idx) {
        if (tout) { if (n < cap) { tout[n] = t; for (int d = 0; d < N; ++d) yout[n * N + d] = y[d]; } else overflow = true; }
        if (dense && idx >= 0) { for (int d = 0; d < N; ++d) dense[((size_t)d * nt + (size_t)idx) * ntraj + traj] = y[d]; ndense = (size_t)idx + 1; }
        ++n;
    }
};

// oderk_adapt, problem.rs:193-358
template <int N, class F>
int solve_adapt(const Tableau& tb, const double* y0, const double* tspan, size_t nt, const double* p, const Opts& o,
                Sink& sink, Counters& cnt, double* yfinal) {
    if (!tb.adaptive) return ERR_WEIGHT_TYPE;       // :204-209
    if (nt == 0) return OK;                          // :211-214
    double t = tspan[0];
    const double tend = tspan[nt - 1];
    const double minstep = o.has_minstep ? o.minstep : std::fabs(tend - t) / 1e18;
    const double maxstep = o.has_maxstep ? o.maxstep : std::fabs(tend - t) / 2.5;
    const double reltol = o.reltol, abstol = o.abstol;
    const int order = tb.order_min;
    double h, tdir, ck[N], cy[N];
    hinit<N, F>(y0, t, tend, order, reltol, abstol, p, h, tdir, ck, o.libm_folded);
    cnt.nfe += 2;
    if (tdir == 0.) return ERR_ZERO_TIMESPAN;        // unreachable for f64 (signum(0)=1)
    double dt;
    if (o.initstep != 0.) {
        if (std::fabs(signum(o.initstep) - tdir) < DBL_EPSILON) dt = o.initstep; else return ERR_INVALID_INITSTEP;
    } else dt = h;
    unsigned timeout = 0;
    bool last_step = std::fabs(t + dt - tend) <= DBL_EPSILON;
    double ylast[N];  // == ys[ys.len()-1]
    for (int d = 0; d < N; ++d) { cy[d] = y0[d]; ylast[d] = y0[d]; }
    sink.push(t, y0, -1);
    size_t it = 1;
    uint64_t attempts = 0;
    cnt.status = ST_DONE;
    for (;;) {
        if (o.max_steps && attempts >= o.max_steps) { cnt.status = ST_MAX_STEPS; break; }
        ++attempts;
        double k[MAXS][MAXN];
        for (int d = 0; d < N; ++d) k[0][d] = ck[d];
        stages<N, F>(tb, t, cy, dt, k, p);           // :261 (stages start from coeff.y)
        cnt.nfe += tb.S - 1;
        double y[N];
        for (int d = 0; d < N; ++d) y[d] = ylast[d];  // :262 (y re-read from the OUTPUT vector)
        // embedded_step :761-771
        double ytrial[N], yerr[N];
        for (int d = 0; d < N; ++d) { ytrial[d] = 0.0; yerr[d] = 0.0; }
        for (int s = 0; s < tb.S; ++s)
            for (int d = 0; d < N; ++d) { ytrial[d] += k[s][d] * tb.b[s][0]; yerr[d] += k[s][d] * tb.b[s][1]; }
        for (int d = 0; d < N; ++d) { yerr[d] = (ytrial[d] - yerr[d]) * dt; ytrial[d] = y[d] + (ytrial[d] * dt); }
        double err, ndt; unsigned tmo;
        stepsize_hw92<N>(dt, tdir, y, ytrial, yerr, order, timeout, abstol, reltol, maxstep, err, ndt, tmo, o.libm_folded);
        timeout = tmo;
        if (err < 1.) {
            cnt.accepted += 1;
            double f1[N];
            if (tb.fsal) { for (int d = 0; d < N; ++d) f1[d] = k[tb.S - 1][d]; }
            else { F::eval(t + dt, ytrial, f1, p); cnt.nfe += 1; }
            if (o.points == 1) {  // Points::Specified :284-300
                while (it < nt && (tdir * tspan[it] < tdir * (t + dt) || last_step)) {
                    double yo[N];
                    hermite<N>(tspan[it], t, dt, y, ytrial, k[0], f1, yo);
                    sink.push(tspan[it], yo, (long long)it);
                    for (int d = 0; d < N; ++d) ylast[d] = yo[d];
                    ++it;
                }
            } else {              // Points::All :301-323
                while (it < nt && tdir * t < tdir * tspan[it] && tdir * tspan[it] < tdir * (t + dt)) {
                    double yo[N];
                    hermite<N>(tspan[it], t, dt, y, ytrial, k[0], f1, yo);
                    sink.push(tspan[it], yo, (long long)it);
                    ++it;
                }
                sink.push(t + dt, ytrial, -1);
                for (int d = 0; d < N; ++d) ylast[d] = ytrial[d];
            }
            for (int d = 0; d < N; ++d) { ck[d] = f1[d]; cy[d] = ytrial[d]; }
            if (last_step) { cnt.t_reached = t + dt; break; }
            t += dt;
            dt = ndt;
            if (tdir * (t + dt + dt / 100.) >= tdir * tend) { dt = tend - t; last_step = true; }
        } else if (std::fabs(ndt) < minstep) {
            cnt.status = ST_MINSTEP_BREAK; cnt.t_reached = t;
            break;
        } else {
            cnt.rejected += 1;
            last_step = false;
            dt = ndt;
            timeout = 5;
        }
    }
    if (cnt.status == ST_MAX_STEPS) cnt.t_reached = t;
    if (yfinal) for (int d = 0; d < N; ++d) yfinal[d] = ylast[d];  // last element of yout
    return OK;
}


// ==========================================================================================
// Stiff path: ode23s problem.rs:409-557, oderosenbrock problem.rs:560-640, fdjacobian :903-926,
// coefficient sets rosenbrock.rs:14-118.
//
// Third-party arithmetic on this path: nalgebra = "0.20" (Cargo.toml; not vendored, no Cargo.lock), used through
// DMatrix/DVector operators, `lu()`, `lu_internal()` and `try_inverse()` at problem.rs:466-504,584-625.  Restated from
// nalgebra's published algorithms (UNPINNED: no reference test exercises them and the crate cannot be built here):
//   * `&A * s`, `&A - &B`, `&a + &b`            elementwise, one rounding per element;
//   * `&M * &v` (gemm -> gemv -> axcpy)         y = M[:,0]*v0, then y = M[:,j]*v_j + y for j = 1.. (column order);
//   * `M.lu()` (linalg/lu.rs)                   partial pivoting on the first largest |.| of the column (icamax),
//                                               multipliers l = a_ji * (1/pivot), update a_jk = (-p_k)*l_j + a_jk,
//                                               a column whose pivot is exactly zero is skipped;
//     `lu_internal()`                           the packed L\U matrix, permutation dropped;
//   * `M.try_inverse()` (linalg/inverse.rs)     closed-form adjugate formulas for n = 1, 2, 3; `do_inverse4` (the 4x4
//                                               cofactor routine) for n = 4; `lu::try_invert_to` for n >= 5.
// ==========================================================================================
enum { ST_INVALID_MATRIX = 3 };

template <int N>
inline void na_lu_packed(double m[N][N]) {
    for (int i = 0; i < N; ++i) {
        int piv = i; double best = std::fabs(m[i][i]);
        for (int r = i + 1; r < N; ++r) { const double v = std::fabs(m[r][i]); if (v > best) { best = v; piv = r; } }
        const double diag = m[piv][i];
        if (diag == 0.0) continue;
        if (piv != i) for (int c = 0; c < N; ++c) { const double tmp = m[i][c]; m[i][c] = m[piv][c]; m[piv][c] = tmp; }
        const double inv = 1.0 / diag;
        for (int r = i + 1; r < N; ++r) m[r][i] = m[r][i] * inv;
        for (int k = i + 1; k < N; ++k) {
            const double a = -m[i][k];
            for (int r = i + 1; r < N; ++r) m[r][k] = a * m[r][i] + m[r][k];
        }
    }
}

template <int N>
inline bool na_try_inverse(const double m[N][N], double o[N][N]) {
    if constexpr (N == 4) {
        // `do_inverse4` (nalgebra 0.20 linalg/inverse.rs; MESA's gluInvertMatrix cofactor expansion) on the column-major slice
        // f[k] = M[k % 4][k / 4]: products and sums left to right, det from the first row of cofactors, then every cofactor is
        // MULTIPLIED by 1/det (the n <= 3 formulas divide)
        double f[16], c[16];
        for (int k = 0; k < 16; ++k) f[k] = m[k % 4][k / 4];
        c[0] = f[5] * f[10] * f[15] - f[5] * f[11] * f[14] - f[9] * f[6] * f[15] + f[9] * f[7] * f[14] + f[13] * f[6] * f[11] - f[13] * f[7] * f[10];
        c[1] = -f[1] * f[10] * f[15] + f[1] * f[11] * f[14] + f[9] * f[2] * f[15] - f[9] * f[3] * f[14] - f[13] * f[2] * f[11] + f[13] * f[3] * f[10];
        c[2] = f[1] * f[6] * f[15] - f[1] * f[7] * f[14] - f[5] * f[2] * f[15] + f[5] * f[3] * f[14] + f[13] * f[2] * f[7] - f[13] * f[3] * f[6];
        c[3] = -f[1] * f[6] * f[11] + f[1] * f[7] * f[10] + f[5] * f[2] * f[11] - f[5] * f[3] * f[10] - f[9] * f[2] * f[7] + f[9] * f[3] * f[6];
        c[4] = -f[4] * f[10] * f[15] + f[4] * f[11] * f[14] + f[8] * f[6] * f[15] - f[8] * f[7] * f[14] - f[12] * f[6] * f[11] + f[12] * f[7] * f[10];
        c[5] = f[0] * f[10] * f[15] - f[0] * f[11] * f[14] - f[8] * f[2] * f[15] + f[8] * f[3] * f[14] + f[12] * f[2] * f[11] - f[12] * f[3] * f[10];
        c[6] = -f[0] * f[6] * f[15] + f[0] * f[7] * f[14] + f[4] * f[2] * f[15] - f[4] * f[3] * f[14] - f[12] * f[2] * f[7] + f[12] * f[3] * f[6];
        c[7] = f[0] * f[6] * f[11] - f[0] * f[7] * f[10] - f[4] * f[2] * f[11] + f[4] * f[3] * f[10] + f[8] * f[2] * f[7] - f[8] * f[3] * f[6];
        c[8] = f[4] * f[9] * f[15] - f[4] * f[11] * f[13] - f[8] * f[5] * f[15] + f[8] * f[7] * f[13] + f[12] * f[5] * f[11] - f[12] * f[7] * f[9];
        c[9] = -f[0] * f[9] * f[15] + f[0] * f[11] * f[13] + f[8] * f[1] * f[15] - f[8] * f[3] * f[13] - f[12] * f[1] * f[11] + f[12] * f[3] * f[9];
        c[10] = f[0] * f[5] * f[15] - f[0] * f[7] * f[13] - f[4] * f[1] * f[15] + f[4] * f[3] * f[13] + f[12] * f[1] * f[7] - f[12] * f[3] * f[5];
        c[11] = -f[0] * f[5] * f[11] + f[0] * f[7] * f[9] + f[4] * f[1] * f[11] - f[4] * f[3] * f[9] - f[8] * f[1] * f[7] + f[8] * f[3] * f[5];
        c[12] = -f[4] * f[9] * f[14] + f[4] * f[10] * f[13] + f[8] * f[5] * f[14] - f[8] * f[6] * f[13] - f[12] * f[5] * f[10] + f[12] * f[6] * f[9];
        c[13] = f[0] * f[9] * f[14] - f[0] * f[10] * f[13] - f[8] * f[1] * f[14] + f[8] * f[2] * f[13] + f[12] * f[1] * f[10] - f[12] * f[2] * f[9];
        c[14] = -f[0] * f[5] * f[14] + f[0] * f[6] * f[13] + f[4] * f[1] * f[14] - f[4] * f[2] * f[13] - f[12] * f[1] * f[6] + f[12] * f[2] * f[5];
        c[15] = f[0] * f[5] * f[10] - f[0] * f[6] * f[9] - f[4] * f[1] * f[10] + f[4] * f[2] * f[9] + f[8] * f[1] * f[6] - f[8] * f[2] * f[5];
        const double det = f[0] * c[0] + f[1] * c[4] + f[2] * c[8] + f[3] * c[12];
        if (det == 0.0) return false;
        const double inv_det = 1.0 / det;
        for (int k = 0; k < 16; ++k) o[k % 4][k / 4] = c[k] * inv_det;
        return true;
    } else if constexpr (N >= 5) {
        // `lu::try_invert_to` (nalgebra 0.20 linalg/lu.rs), what try_inverse does for n >= 5: Gaussian elimination with partial
        // pivoting on a copy (first largest |.|, multipliers a_ji * (1/pivot), update a_jk = (-p_k) l_j + a_jk; a zero pivot fails),
        // the row swaps applied to an identity `o`; then o <- L^-1 o (unit diagonal, forward, column by column) and o <- U^-1 o
        // (backward: coeff = b_i / u_ii, b_j = (-coeff) u_ji + b_j for j < i)
        double a[N][N];
        for (int r = 0; r < N; ++r) {
            for (int cc = 0; cc < N; ++cc) { a[r][cc] = m[r][cc]; o[r][cc] = r == cc ? 1.0 : 0.0; }
        }
        bool ok = true;
        for (int i = 0; i < N; ++i) {
            int piv = i;
            double best = std::fabs(a[i][i]);
            for (int r = i + 1; r < N; ++r) { const double v = std::fabs(a[r][i]); if (v > best) { best = v; piv = r; } }
            for (int r = i + 1; r < N; ++r) {
                if (piv == r) {
                    for (int cc = 0; cc < N; ++cc) {
                        double tmp = a[i][cc]; a[i][cc] = a[r][cc]; a[r][cc] = tmp;
                        tmp = o[i][cc]; o[i][cc] = o[r][cc]; o[r][cc] = tmp;
                    }
                }
            }
            const double diag = a[i][i];
            if (diag == 0.0) ok = false;
            const double inv = 1.0 / diag;
            for (int r = i + 1; r < N; ++r) a[r][i] = a[r][i] * inv;
            for (int k = i + 1; k < N; ++k) {
                const double p = -a[i][k];
                for (int r = i + 1; r < N; ++r) a[r][k] = p * a[r][i] + a[r][k];
            }
        }
        if (!ok) return false;   // (the reference returns at the first zero pivot; nothing computed after it is used)
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i + 1 < N; ++i) {
                const double coeff = o[i][k] / 1.0;
                for (int r = i + 1; r < N; ++r) o[r][k] = (-coeff) * a[r][i] + o[r][k];
            }
        }
        for (int k = 0; k < N; ++k) {
            for (int i = N - 1; i >= 0; --i) {
                const double diag = a[i][i];
                if (diag == 0.0) ok = false;
                const double coeff = o[i][k] / diag;
                o[i][k] = coeff;
                for (int r = 0; r < i; ++r) o[r][k] = (-coeff) * a[r][i] + o[r][k];
            }
        }
        if (!ok) return false;
        return true;
    } else
    if constexpr (N == 1) {
        const double det = m[0][0];
        if (det == 0.0) return false;
        o[0][0] = 1.0 / det;
    } else if constexpr (N == 2) {
        const double m11 = m[0][0], m12 = m[0][1], m21 = m[1][0], m22 = m[1][1];
        const double det = m11 * m22 - m21 * m12;
        if (det == 0.0) return false;
        o[0][0] = m22 / det; o[0][1] = -m12 / det; o[1][0] = -m21 / det; o[1][1] = m11 / det;
    } else {
        const double m11 = m[0][0], m12 = m[0][1], m13 = m[0][2], m21 = m[1][0], m22 = m[1][1], m23 = m[1][2],
                     m31 = m[2][0], m32 = m[2][1], m33 = m[2][2];
        const double minor_m12_m23 = m22 * m33 - m32 * m23;
        const double minor_m11_m23 = m21 * m33 - m31 * m23;
        const double minor_m11_m22 = m21 * m32 - m31 * m22;
        const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        if (det == 0.0) return false;
        o[0][0] = minor_m12_m23 / det; o[0][1] = (m13 * m32 - m33 * m12) / det; o[0][2] = (m12 * m23 - m22 * m13) / det;
        o[1][0] = -minor_m11_m23 / det; o[1][1] = (m11 * m33 - m31 * m13) / det; o[1][2] = (m13 * m21 - m23 * m11) / det;
        o[2][0] = minor_m11_m22 / det; o[2][1] = (m12 * m31 - m32 * m11) / det; o[2][2] = (m11 * m22 - m21 * m12) / det;
    }
    return true;
}

template <int N>
inline void na_gemv(const double m[N][N], const double* v, double* y) {
    for (int r = 0; r < N; ++r) y[r] = m[r][0] * v[0];
    for (int j = 1; j < N; ++j)
        for (int r = 0; r < N; ++r) y[r] = m[r][j] * v[j] + y[r];
}

// fdjacobian, problem.rs:903-926
template <int N, class F>
inline void fdjacobian(double t, const double* x, const double* p, double J[N][N], uint32_t& nfe) {
    double ftx[N];
    F::eval(t, x, ftx, p);
    for (int c = 0; c < N; ++c) {
        double xj = x[c];
        if (xj == 0.0) xj += 1.0;
        const double dxj = xj * 0.01;
        double tmp[N], yj[N];
        for (int d = 0; d < N; ++d) tmp[d] = x[d];
        tmp[c] += dxj;
        F::eval(t, tmp, yj, p);
        for (int r = 0; r < N; ++r) J[r][c] = (yj[r] - ftx[r]) / dxj;
    }
    nfe += N + 1;
}

template <int N>
inline double pnorm2(const double* v, int libm_folded) {   // types.rs:109-115 with PNorm::default() = P(2)
    double s = 0.0;
    for (int d = 0; d < N; ++d) { const double a = std::fabs(v[d]); s = s + a * a; }
    return libm_folded ? std::sqrt(s) : std::pow(s, 1.0 * (1. / 2.0));
}

// ode23s, problem.rs:409-557.  `true_inverse` (TEXTBOOK set) inverts W itself instead of its packed LU factors
// (the reference's `w.lu().lu_internal().clone()` followed by `try_inverse`, :469,481).
template <int N, class F>
int solve_ode23s(const double* y0, const double* tspan, size_t nt, const double* p, const Opts& o, bool true_inverse,
                 Sink& sink, Counters& cnt, double* yfinal) {
    if (nt == 0) return OK;
    double t = tspan[0];
    const double tfinal = tspan[nt - 1];
    const double reltol = o.reltol, abstol = o.abstol;
    const double minstep = o.has_minstep ? o.minstep : std::fabs(tfinal - t) / 1e18;
    const double maxstep = o.has_maxstep ? o.maxstep : std::fabs(tfinal - t) / 2.5;
    const double two_sqrt = std::sqrt(2.0);
    const double d = 1. / (2. + two_sqrt);
    const double e32 = 6. + two_sqrt;
    double hh, tdir, f0[N];
    if (o.initstep == 0.) {
        hinit<N, F>(y0, t, tfinal, 3, reltol, abstol, p, hh, tdir, f0, o.libm_folded);
        cnt.nfe += 2;
    } else {
        hh = o.initstep; tdir = signum(tfinal - t);
        F::eval(t, y0, f0, p);
        cnt.nfe += 1;
    }
    double h = tdir * fmin_rs(std::fabs(hh), maxstep);
    sink.push(t, y0, 0);
    double tlast = t;   // tout[tout.len() - 1]
    double jac[N][N];
    fdjacobian<N, F>(t, y0, p, jac, cnt.nfe);
    double y[N];
    for (int i = 0; i < N; ++i) y[i] = y0[i];
    uint64_t attempts = 0;
    cnt.status = ST_DONE;
    bool guard = false;
    while (std::fabs(t - tfinal) > 0. && minstep < std::fabs(h)) {
        if (o.max_steps && attempts >= o.max_steps) { guard = true; break; }
        ++attempts;
        if (std::fabs(t - tfinal) < std::fabs(h)) h = tfinal - t;
        const double hd = 1.0 * (h * d);
        double w[N][N];
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) w[r][c] = (r == c ? 1.0 : 0.0) - jac[r][c] * hd;
        if (N * N != 1 && !true_inverse) na_lu_packed<N>(w);
        double fdt[N];
        F::eval(t + h / 100., y, fdt, p);
        for (int i = 0; i < N; ++i) { const double fdti = fdt[i] - f0[i]; fdt[i] = fdti * ((h * d) / (h / 100.)); }
        double winv[N][N];
        if (!na_try_inverse<N>(w, winv)) { cnt.nfe += 1; cnt.status = ST_INVALID_MATRIX; cnt.t_reached = t; if (yfinal) for (int i = 0; i < N; ++i) yfinal[i] = y[i]; return OK; }
        double rhs[N], k1[N], k2[N], k3[N], f1[N], f2[N], f1y[N], ynew[N];
        for (int i = 0; i < N; ++i) rhs[i] = f0[i] + fdt[i];
        na_gemv<N>(winv, rhs, k1);
        for (int i = 0; i < N; ++i) f1y[i] = y[i] + k1[i] * 0.5 * h;
        F::eval(t + 0.5 * h, f1y, f1, p);
        for (int i = 0; i < N; ++i) rhs[i] = f1[i] - k1[i];
        na_gemv<N>(winv, rhs, k2);
        for (int i = 0; i < N; ++i) k2[i] = k2[i] + k1[i];
        for (int i = 0; i < N; ++i) ynew[i] = y[i] + k2[i] * h;
        F::eval(t + h, ynew, f2, p);
        cnt.nfe += 3;
        for (int i = 0; i < N; ++i) rhs[i] = ((f2[i] - (k2[i] - f1[i]) * (1.0 * e32)) - (k1[i] - f0[i]) * (1.0 * 2.)) + fdt[i];
        na_gemv<N>(winv, rhs, k3);
        double kerr[N];
        for (int i = 0; i < N; ++i) kerr[i] = (k1[i] - k2[i] * (1.0 * 2.)) + k3[i];
        const double err = pnorm2<N>(kerr, o.libm_folded) * (std::fabs(h) / 6.);
        const double delta = fmax_rs(fmax_rs(pnorm2<N>(y, o.libm_folded), pnorm2<N>(ynew, o.libm_folded)) * reltol, 1.0 * abstol);
        if (err <= delta) {
            cnt.accepted += 1;
            for (size_t q = 0; q < nt; ++q) {
                const double toi = tspan[q];
                if (toi > t && toi <= t + h) {
                    const double s = (toi - t) / h;
                    const double c1 = 1.0 * (s * (1. - s) / (1. - 2. * d));
                    const double c2 = 1.0 * (s * (s - 2. * d) / (1. - 2. * d));
                    double ytmp[N];
                    for (int i = 0; i < N; ++i) { const double kt = k1[i] * c1 + k2[i] * c2; ytmp[i] = y[i] + kt * h; }
                    sink.push(toi, ytmp, (long long)q);
                    tlast = toi;
                }
            }
            if (o.points == 0 && std::fabs(tlast - (t + h)) > DBL_EPSILON) { sink.push(t + h, ynew, -1); tlast = t + h; }
            t += h;
            for (int i = 0; i < N; ++i) { y[i] = ynew[i]; f0[i] = f2[i]; }
            fdjacobian<N, F>(t, y, p, jac, cnt.nfe);
        } else {
            cnt.rejected += 1;
        }
        const double r = delta / err;
        h = fmin_rs(maxstep, std::pow(r, 1. / 3.) * std::fabs(h) * 0.8) * tdir;
    }
    if (guard) cnt.status = ST_MAX_STEPS;
    else if (std::fabs(t - tfinal) > 0.) cnt.status = ST_MINSTEP_BREAK;   // the loop ended on `minstep < |h|` (silently, like the reference)
    cnt.t_reached = t;
    if (yfinal) for (int i = 0; i < N; ++i) yfinal[i] = y[i];
    return OK;
}

struct RosenbrockCoeffs { double gamma, a[4][4], b[4], c[4][4]; };
// rosenbrock.rs:16-68 and :71-117 (Matrix4::new takes its arguments row-major: a[i][j] = a[(i, j)])
inline RosenbrockCoeffs rosenbrock_coeffs(int which) {
    if (which == 0)
        return RosenbrockCoeffs{0.231,
            {{0., 0., 0., 0.}, {2., 0., 0., 0.}, {4.452470820736, 4.16352878860, 0., 0.}, {4.452470820736, 4.16352878860, 0., 0.}},
            {3.95750374663, 4.62489238836, 0.617477263873, 1.2826129456},
            {{0., 0., 0., 0.}, {-5.07167533877, 0., 0., 0.}, {6.02015272865, 0.1597500684673, 0., 0.}, {-1.856343618677, -8.50538085819, -2.08407513602, 0.0}}};
    return RosenbrockCoeffs{0.5,
        {{0., 0., 0., 0.}, {2., 0., 0., 0.}, {48. / 25., 6. / 250., 0., 0.}, {48. / 25., 6. / 250., 0., 0.}},
        {19. / 9., 1. / 2., 25. / 108., 125. / 108.},
        {{0., 0., 0., 0.}, {-8., 0., 0., 0.}, {372. / 25., 12. / 5., 0., 0.}, {-112. / 125., -54. / 125., -2. / 5., 0.}}};
}

// oderosenbrock, problem.rs:560-640.  yout [nt][N] (rows after an InvalidMatrix stop are left untouched); the
// reference returns tout = diff(tspan) (sic, :572,639) — callers of the oracle rebuild that from tspan.
template <int N, class F>
int solve_rosenbrock(const RosenbrockCoeffs& co, const double* y0, const double* tspan, size_t nt, const double* p,
                     double* yout, double* yfinal, uint8_t* status, uint32_t* nfe_out, size_t* rows) {
    uint32_t nfe = 0;
    double x[N];
    for (int i = 0; i < N; ++i) x[i] = y0[i];
    if (yout) for (int i = 0; i < N; ++i) yout[i] = x[i];
    size_t nrows = nt ? 1 : 0;
    uint8_t st = ST_DONE;
    for (size_t solstep = 0; solstep + 1 < nt; ++solstep) {
        const double hs = tspan[solstep + 1] - tspan[solstep];
        const double ts = tspan[solstep];
        double xs[N];
        for (int i = 0; i < N; ++i) xs[i] = x[i];
        double dfdx[N][N], jac[N][N], jinv[N][N];
        fdjacobian<N, F>(ts, xs, p, dfdx, nfe);
        const double dg = 1.0 * (1. / (co.gamma * hs));
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) jac[r][c] = (r == c ? dg : 0.0) - dfdx[r][c];
        if (!na_try_inverse<N>(jac, jinv)) { st = ST_INVALID_MATRIX; break; }
        double g[4][N], fx[N];
        F::eval(ts + co.b[0] * hs, xs, fx, p); ++nfe;
        na_gemv<N>(jinv, fx, g[0]);
        double nx[N];
        for (int i = 0; i < N; ++i) nx[i] = x[i];
        for (int i = 0; i < N; ++i) nx[i] += g[0][i] * co.b[0];
        for (int i = 1; i < 4; ++i) {
            double dx[N], df[N];
            for (int d = 0; d < N; ++d) { dx[d] = 0.0; df[d] = 0.0; }
            for (int j = 0; j < i - 1; ++j)                 // `.take(i - 1)` (:614): the newest stage is left out
                for (int d = 0; d < N; ++d) { dx[d] += g[j][d] * co.a[i][j]; df[d] += g[j][d] * co.c[i][j]; }
            double xa[N];
            for (int d = 0; d < N; ++d) xa[d] = xs[d] + dx[d];
            F::eval(ts + co.b[i] * hs, xa, fx, p); ++nfe;
            na_gemv<N>(jinv, fx, g[i]);
            for (int d = 0; d < N; ++d) g[i][d] = g[i][d] + df[d] * (1. / hs);
            for (int d = 0; d < N; ++d) nx[d] += g[i][d] * co.b[i];
        }
        for (int i = 0; i < N; ++i) x[i] = nx[i];
        if (yout) for (int i = 0; i < N; ++i) yout[(solstep + 1) * N + i] = x[i];
        ++nrows;
    }
    if (yfinal) for (int i = 0; i < N; ++i) yfinal[i] = x[i];
    if (status) *status = st;
    if (nfe_out) *nfe_out = nfe;
    if (rows) *rows = nrows;
    return OK;
}


// ------------------------------------------------------------------------------------------
// Allocation-faithful variant of oderk_adapt (problem.rs:193-358 with `Y = Vec<f64>`): the same arithmetic as solve_adapt,
// but with the reference's memory behaviour — every state is a heap vector, `coeff.clone()` (:261), one `yi` clone and one
// freshly returned `Vec` per stage (:722,733), `y` cloned from the output vector (:262), two clones in embedded_step
// (:755,759), `ys.push(ytrial.clone())` into a growing Vec<Vec<f64>> (:321) — about 20 allocations per dopri5 step attempt.
// Used only to report how the CPU baseline moves with the reference's allocation pattern (bench.py cpu_baseline);
// results are bit-identical to solve_adapt (tests/test_oracle_cpu.py).  Points::All.
// ------------------------------------------------------------------------------------------
typedef std::vector<double> VecY;
struct CoefficientPointV { VecY k, y; };

template <int N, class F>
inline VecY rhs_vec(double t, const VecY& y, const double* p) {
    VecY out(N);
    F::eval(t, y.data(), out.data(), p);
    return out;
}

template <int N, class F>
int solve_adapt_vec(const Tableau& tb, const double* y0p, const double* tspan_in, size_t nt, const double* p, const Opts& o,
                    Counters& cnt, double* yfinal) {
    if (!tb.adaptive) return ERR_WEIGHT_TYPE;
    if (nt == 0) return OK;
    const VecY y0(y0p, y0p + N);
    double t = tspan_in[0];
    const double tend = tspan_in[nt - 1];
    const double minstep = o.has_minstep ? o.minstep : std::fabs(tend - t) / 1e18;
    const double maxstep = o.has_maxstep ? o.maxstep : std::fabs(tend - t) / 2.5;
    const double reltol = o.reltol, abstol = o.abstol;
    const int order = tb.order_min;
    double h, tdir, f0a[N];
    hinit<N, F>(y0.data(), t, tend, order, reltol, abstol, p, h, tdir, f0a, o.libm_folded);
    cnt.nfe += 2;
    double dt;
    if (o.initstep != 0.) {
        if (std::fabs(signum(o.initstep) - tdir) < DBL_EPSILON) dt = o.initstep; else return ERR_INVALID_INITSTEP;
    } else dt = h;
    unsigned timeout = 0;
    bool last_step = std::fabs(t + dt - tend) <= DBL_EPSILON;
    std::vector<double> tout; tout.reserve(nt); tout.push_back(t);
    std::vector<VecY> ys; ys.reserve(nt); ys.push_back(y0);
    CoefficientPointV coeff{VecY(f0a, f0a + N), y0};
    size_t it = 1;
    uint64_t attempts = 0;
    cnt.status = ST_DONE;
    for (;;) {
        if (o.max_steps && attempts >= o.max_steps) { cnt.status = ST_MAX_STEPS; break; }
        ++attempts;
        // calc_coefficients(btab, t, coeff.clone(), dt)
        std::vector<CoefficientPointV> coeffs; coeffs.reserve(tb.S);
        coeffs.push_back(coeff);
        for (int row = 1; row < tb.S; ++row) {
            VecY yi = coeffs[0].y;
            for (int col = 0; col < row; ++col) {
                const VecY& k = coeffs[col].k;
                for (int d = 0; d < N; ++d) yi[d] += k[d] * (tb.a[row][col] * dt);
            }
            const double tn = t + tb.c[row] * dt;
            VecY kn = rhs_vec<N, F>(tn, yi, p);
            coeffs.push_back(CoefficientPointV{std::move(kn), std::move(yi)});
        }
        cnt.nfe += tb.S - 1;
        const VecY y = ys[ys.size() - 1];
        // embedded_step
        VecY ytrial = y; for (int d = 0; d < N; ++d) ytrial[d] = 0.0;
        VecY yerr = ytrial;
        for (int s = 0; s < tb.S; ++s) {
            const VecY& k = coeffs[s].k;
            for (int d = 0; d < N; ++d) { ytrial[d] += k[d] * tb.b[s][0]; yerr[d] += k[d] * tb.b[s][1]; }
        }
        for (int d = 0; d < N; ++d) { yerr[d] = (ytrial[d] - yerr[d]) * dt; ytrial[d] = y[d] + (ytrial[d] * dt); }
        double err, ndt; unsigned tmo;
        stepsize_hw92<N>(dt, tdir, y.data(), ytrial.data(), yerr.data(), order, timeout, abstol, reltol, maxstep, err, ndt, tmo, o.libm_folded);
        timeout = tmo;
        if (err < 1.) {
            cnt.accepted += 1;
            const VecY& f0 = coeffs[0].k;
            VecY f1;
            if (tb.fsal) f1 = coeffs[tb.S - 1].k; else { f1 = rhs_vec<N, F>(t + dt, ytrial, p); cnt.nfe += 1; }
            while (it < nt && tdir * t < tdir * tspan_in[it] && tdir * tspan_in[it] < tdir * (t + dt)) {
                VecY yo(N);
                hermite<N>(tspan_in[it], t, dt, y.data(), ytrial.data(), f0.data(), f1.data(), yo.data());
                ys.push_back(std::move(yo));
                tout.push_back(tspan_in[it]);
                ++it;
            }
            ys.push_back(ytrial);
            tout.push_back(t + dt);
            coeff = CoefficientPointV{std::move(f1), std::move(ytrial)};
            if (last_step) { cnt.t_reached = t + dt; break; }
            t += dt;
            dt = ndt;
            if (tdir * (t + dt + dt / 100.) >= tdir * tend) { dt = tend - t; last_step = true; }
        } else if (std::fabs(ndt) < minstep) {
            cnt.status = ST_MINSTEP_BREAK; cnt.t_reached = t;
            break;
        } else {
            cnt.rejected += 1;
            last_step = false;
            dt = ndt;
            timeout = 5;
        }
    }
    if (cnt.status == ST_MAX_STEPS) cnt.t_reached = t;
    if (yfinal) for (int d = 0; d < N; ++d) yfinal[d] = ys.back()[d];
    return OK;
}

template <class Fn>
int dispatch_rhs(int rhs, Fn&& fn) {
    switch (rhs) {
        case LORENZ: return fn(LorenzF());
        case VANDERPOL: return fn(VdpF());
        case PENDULUM: return fn(PendulumF());
        case RING8: return fn(RingF<8>());
        case RING16: return fn(RingF<16>());
        case RING32: return fn(RingF<32>());
        case RING64: return fn(RingF<64>());
        case CVDP2: return fn(CoupledVdpF<2>());
        case CVDP3: return fn(CoupledVdpF<3>());
        case CVDP6: return fn(CoupledVdpF<6>());
        case CVDP12: return fn(CoupledVdpF<12>());
        default: return ERR_UNSUPPORTED;
    }
}
inline bool is_fixed(int m) { return m == FEULER || m == HEUN || m == MIDPOINT || m == ODE4; }
inline bool is_adapt(int m) { return m == ODE23 || m == ODE45 || m == ODE45FE || m == ODE78 || m == ODE21; }

}  // namespace

extern "C" {

struct oracle_opts {
    double reltol, abstol;
    int has_minstep; double minstep;
    int has_maxstep; double maxstep;
    double initstep;
    int points;
    uint64_t max_steps;
    int libm_folded;
};

int oracle_rhs_dim(int rhs) { return rhs == LORENZ ? 3 : (rhs == VANDERPOL || rhs == PENDULUM) ? 2 : rhs == RING8 ? 8 : rhs == RING16 ? 16 : rhs == RING32 ? 32 : rhs == RING64 ? 64 : rhs == CVDP2 ? 4 : rhs == CVDP3 ? 6 : rhs == CVDP6 ? 12 : rhs == CVDP12 ? 24 : -1; }
int oracle_rhs_nparams(int rhs) { return rhs == LORENZ ? 3 : rhs == VANDERPOL ? 1 : rhs == PENDULUM ? 3 : (rhs >= RING8 && rhs <= CVDP12) ? 2 : -1; }

// itertools_num::linspace as recalled (unpinned third-party): a + ((b-a)/(n-1))*i
void oracle_linspace(double a, double b, size_t n, double* out) {
    if (n == 0) return;
    const double step = (n > 1) ? (b - a) / (double)(n - 1) : 0.0;
    for (size_t i = 0; i < n; ++i) out[i] = a + step * (double)i;
}

// Tableau export for tests: a[S*S] row-major, b[S*2], c[S].
int oracle_tableau(int method, int set, int* S, double* a, double* b, double* c, int* adaptive, int* order_min, int* fsal) {
    Tableau t;
    if (!make_tableau(method, set, t)) return ERR_UNSUPPORTED;
    *S = t.S; *adaptive = t.adaptive; *order_min = t.order_min; *fsal = t.fsal;
    for (int r = 0; r < t.S; ++r) { for (int cc = 0; cc < t.S; ++cc) a[r * t.S + cc] = t.a[r][cc]; b[2 * r] = t.b[r][0]; b[2 * r + 1] = t.b[r][1]; c[r] = t.c[r]; }
    return OK;
}

// Single trajectory, fixed step.  yout [nt][n] (may be NULL), yfinal [n] (may be NULL).
int oracle_solve_fixed(int method, int rhs, const double* params, const double* y0, const double* tspan, size_t nt,
                       double* yout, double* yfinal, uint32_t* nfe) {
    if (!is_fixed(method)) return is_adapt(method) ? ERR_WEIGHT_TYPE : ERR_UNSUPPORTED;
    if (nt == 0) return ERR_INVALID_ARG;  // the reference panics (usize underflow), problem.rs:377
    Tableau tb; make_tableau(method, 0, tb);
    return dispatch_rhs(rhs, [&](auto f) { using F = decltype(f); solve_fixed<F::N, F>(tb, y0, tspan, nt, params, yout, yfinal, nfe); return (int)OK; });
}

// Single trajectory, adaptive.  tout/yout capacity `cap` rows (may be NULL to only count); *nout = rows the
// reference would have produced.  counts = {accepted, rejected, nfe}.
int oracle_solve_adaptive(int method, int set, int rhs, const double* params, const double* y0, const double* tspan,
                          size_t nt, const oracle_opts* o, double* tout, double* yout, size_t cap, size_t* nout,
                          double* yfinal, uint32_t* counts, uint8_t* status, double* t_reached) {
    if (!is_adapt(method)) return is_fixed(method) ? ERR_WEIGHT_TYPE : ERR_UNSUPPORTED;
    Tableau tb; make_tableau(method, set, tb);
    Opts op{o->reltol, o->abstol, o->has_minstep, o->minstep, o->has_maxstep, o->maxstep, o->initstep, o->points, o->max_steps, o->libm_folded};
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        Sink sink; sink.tout = tout; sink.yout = yout; sink.cap = cap; sink.N = F::N;
        Counters c;
        int rc = solve_adapt<F::N, F>(tb, y0, tspan, nt, params, op, sink, c, yfinal);
        if (nout) *nout = sink.n;
        if (counts) { counts[0] = c.accepted; counts[1] = c.rejected; counts[2] = c.nfe; }
        if (status) *status = c.status;
        if (t_reached) *t_reached = c.t_reached;
        return rc;
    });
}

// Ensemble, adaptive, final state + counters (what BASELINE configs 2/4/5 need).  y0 [ntraj][n] AoS;
// params [ntraj][np] if params_per_traj else [np].  Parallelised over trajectories with OpenMP
// (stand-in for rayon par_iter).  nthreads <= 0 -> all cores.
int oracle_ensemble_adaptive(int method, int set, int rhs, size_t ntraj, const double* y0, const double* params,
                             int params_per_traj, const double* tspan, size_t nt, const oracle_opts* o, int nthreads,
                             double* yfinal, uint32_t* accepted, uint32_t* rejected, uint32_t* nfe, uint8_t* status,
                             double* t_reached) {
    if (!is_adapt(method)) return is_fixed(method) ? ERR_WEIGHT_TYPE : ERR_UNSUPPORTED;
    Tableau tb; make_tableau(method, set, tb);
    Opts op{o->reltol, o->abstol, o->has_minstep, o->minstep, o->has_maxstep, o->maxstep, o->initstep, o->points, o->max_steps, o->libm_folded};
    const int np = oracle_rhs_nparams(rhs);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
        int rc_all = OK;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 256)
        for (long long i = 0; i < (long long)ntraj; ++i) {
            Sink sink; sink.N = N;
            Counters c;
            double yf[N];
            const double* p = params_per_traj ? params + (size_t)i * np : params;
            int rc = solve_adapt<N, F>(tb, y0 + (size_t)i * N, tspan, nt, p, op, sink, c, yf);
            if (rc != OK) {
#pragma omp critical
                rc_all = rc;
            }
            if (yfinal) for (int d = 0; d < N; ++d) yfinal[(size_t)i * N + d] = yf[d];
            if (accepted) accepted[i] = c.accepted;
            if (rejected) rejected[i] = c.rejected;
            if (nfe) nfe[i] = c.nfe;
            if (status) status[i] = c.status;
            if (t_reached) t_reached[i] = c.t_reached;
        }
        return rc_all;
    });
}

// Ensemble, adaptive, allocation-faithful variant (solve_adapt_vec): final states + counters.
int oracle_ensemble_adaptive_vec(int method, int set, int rhs, size_t ntraj, const double* y0, const double* params,
                                 int params_per_traj, const double* tspan, size_t nt, const oracle_opts* o, int nthreads,
                                 double* yfinal, uint32_t* accepted, uint32_t* rejected, uint32_t* nfe, uint8_t* status) {
    if (!is_adapt(method)) return is_fixed(method) ? ERR_WEIGHT_TYPE : ERR_UNSUPPORTED;
    Tableau tb; make_tableau(method, set, tb);
    Opts op{o->reltol, o->abstol, o->has_minstep, o->minstep, o->has_maxstep, o->maxstep, o->initstep, 0, o->max_steps, o->libm_folded};
    const int np = oracle_rhs_nparams(rhs);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 256)
        for (long long i = 0; i < (long long)ntraj; ++i) {
            Counters c; double yf[N] = {};
            const double* p = params_per_traj ? params + (size_t)i * np : params;
            solve_adapt_vec<N, F>(tb, y0 + (size_t)i * N, tspan, nt, p, op, c, yf);
            if (yfinal) for (int d = 0; d < N; ++d) yfinal[(size_t)i * N + d] = yf[d];
            if (accepted) accepted[i] = c.accepted;
            if (rejected) rejected[i] = c.rejected;
            if (nfe) nfe[i] = c.nfe;
            if (status) status[i] = c.status;
        }
        return (int)OK;
    });
}

// Ensemble, adaptive, dense output filtered to the tspan times (BASELINE config 3): runs the reference's
// Points::All and keeps the rows whose time is a tspan entry.  out is SoA [n][nt][ntraj]; row 0 = y0, rows
// 1..nt-2 = Hermite values the reference emitted, row nt-1 = last pushed state; nvalid[i] = number of leading
// tspan rows the reference emitted (rows >= nvalid[i], except the last, are left untouched).
int oracle_ensemble_dense(int method, int set, int rhs, size_t ntraj, const double* y0, const double* params,
                          int params_per_traj, const double* tspan, size_t nt, const oracle_opts* o, int nthreads,
                          double* out, uint32_t* nvalid, uint32_t* accepted, uint32_t* rejected, uint8_t* status) {
    if (!is_adapt(method)) return is_fixed(method) ? ERR_WEIGHT_TYPE : ERR_UNSUPPORTED;
    if (nt < 2) return ERR_INVALID_ARG;
    Tableau tb; make_tableau(method, set, tb);
    Opts op{o->reltol, o->abstol, o->has_minstep, o->minstep, o->has_maxstep, o->maxstep, o->initstep, 0, o->max_steps, o->libm_folded};
    const int np = oracle_rhs_nparams(rhs);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
        for (long long i = 0; i < (long long)ntraj; ++i) {
            const double* p = params_per_traj ? params + (size_t)i * np : params;
            Sink sink; sink.N = N; sink.dense = out; sink.nt = nt; sink.ntraj = ntraj; sink.traj = (size_t)i;
            Counters c; double yf[N];
            for (int d = 0; d < N; ++d) out[((size_t)d * nt + 0) * ntraj + i] = y0[(size_t)i * N + d];
            solve_adapt<N, F>(tb, y0 + (size_t)i * N, tspan, nt, p, op, sink, c, yf);
            for (int d = 0; d < N; ++d) out[((size_t)d * nt + (nt - 1)) * ntraj + i] = yf[d];
            if (nvalid) nvalid[i] = (uint32_t)sink.ndense;
            if (accepted) accepted[i] = c.accepted;
            if (rejected) rejected[i] = c.rejected;
            if (status) status[i] = c.status;
        }
        return (int)OK;
    });
}

// Ensemble, fixed step, final state only.
int oracle_ensemble_fixed(int method, int rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                          const double* tspan, size_t nt, int nthreads, double* yfinal) {
    if (!is_fixed(method)) return is_adapt(method) ? ERR_WEIGHT_TYPE : ERR_UNSUPPORTED;
    if (nt == 0) return ERR_INVALID_ARG;
    Tableau tb; make_tableau(method, 0, tb);
    const int np = oracle_rhs_nparams(rhs);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
        for (long long i = 0; i < (long long)ntraj; ++i) {
            const double* p = params_per_traj ? params + (size_t)i * np : params;
            solve_fixed<N, F>(tb, y0 + (size_t)i * N, tspan, nt, p, nullptr, yfinal + (size_t)i * N, nullptr);
        }
        return (int)OK;
    });
}


// ---- stiff path exports ----
// Single trajectory ode23s.  Same calling convention as oracle_solve_adaptive; tableau set 1 (TEXTBOOK) = true W^-1.
int oracle_solve_ode23s(int set, int rhs, const double* params, const double* y0, const double* tspan, size_t nt,
                        const oracle_opts* o, double* tout, double* yout, size_t cap, size_t* nout, double* yfinal,
                        uint32_t* counts, uint8_t* status, double* t_reached) {
    Opts op{o->reltol, o->abstol, o->has_minstep, o->minstep, o->has_maxstep, o->maxstep, o->initstep, o->points, o->max_steps, o->libm_folded};
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        if constexpr (F::N > 32) return (int)ERR_UNSUPPORTED; else {
        Sink sink; sink.tout = tout; sink.yout = yout; sink.cap = cap; sink.N = F::N;
        Counters c;
        int rc = solve_ode23s<F::N, F>(y0, tspan, nt, params, op, set == 1, sink, c, yfinal);
        if (nout) *nout = sink.n;
        if (counts) { counts[0] = c.accepted; counts[1] = c.rejected; counts[2] = c.nfe; }
        if (status) *status = c.status;
        if (t_reached) *t_reached = c.t_reached;
        return rc; }
    });
}

// Ensemble ode23s: final loop state + counters, optionally the rows at the tspan times (dense SoA [n][nt][ntraj],
// row q = value the reference pushed for tspan[q]; nvalid = 1 + index of the last tspan row pushed).
int oracle_ensemble_ode23s(int set, int rhs, size_t ntraj, const double* y0, const double* params, int params_per_traj,
                           const double* tspan, size_t nt, const oracle_opts* o, int nthreads, double* yfinal,
                           uint32_t* accepted, uint32_t* rejected, uint32_t* nfe, uint8_t* status, double* t_reached,
                           double* dense, uint32_t* nvalid, uint32_t* nrows) {
    Opts op{o->reltol, o->abstol, o->has_minstep, o->minstep, o->has_maxstep, o->maxstep, o->initstep, o->points, o->max_steps, o->libm_folded};
    const int np = oracle_rhs_nparams(rhs);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
        if constexpr (N > 32) return (int)ERR_UNSUPPORTED; else {
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
        for (long long i = 0; i < (long long)ntraj; ++i) {
            const double* p = params_per_traj ? params + (size_t)i * np : params;
            Sink sink; sink.N = N;
            if (dense) { sink.dense = dense; sink.nt = nt; sink.ntraj = ntraj; sink.traj = (size_t)i; }
            Counters c; double yf[N];
            solve_ode23s<N, F>(y0 + (size_t)i * N, tspan, nt, p, op, set == 1, sink, c, yf);
            if (yfinal) for (int d = 0; d < N; ++d) yfinal[(size_t)i * N + d] = yf[d];
            if (accepted) accepted[i] = c.accepted;
            if (rejected) rejected[i] = c.rejected;
            if (nfe) nfe[i] = c.nfe;
            if (status) status[i] = c.status;
            if (t_reached) t_reached[i] = c.t_reached;
            if (nvalid) nvalid[i] = (uint32_t)sink.ndense;
            if (nrows) nrows[i] = (uint32_t)sink.n;
        }
        return (int)OK; }
    });
}

// Single trajectory Rosenbrock (which: 0 = kr4 / Ode4skr, 1 = s4 / Ode4ss).  yout [nt][n].
int oracle_solve_rosenbrock(int which, int rhs, const double* params, const double* y0, const double* tspan, size_t nt,
                            double* yout, double* yfinal, uint8_t* status, uint32_t* nfe, size_t* rows) {
    if (which != 0 && which != 1) return ERR_INVALID_ARG;
    const RosenbrockCoeffs co = rosenbrock_coeffs(which);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        if constexpr (F::N > 32) return (int)ERR_UNSUPPORTED; else
        return solve_rosenbrock<F::N, F>(co, y0, tspan, nt, params, yout, yfinal, status, nfe, rows);
    });
}

int oracle_ensemble_rosenbrock(int which, int rhs, size_t ntraj, const double* y0, const double* params,
                               int params_per_traj, const double* tspan, size_t nt, int nthreads, double* yfinal,
                               uint8_t* status, double* yall /* [ntraj][nt][n] or NULL */) {
    if (which != 0 && which != 1) return ERR_INVALID_ARG;
    const RosenbrockCoeffs co = rosenbrock_coeffs(which);
    const int np = oracle_rhs_nparams(rhs);
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
        if constexpr (N > 32) return (int)ERR_UNSUPPORTED; else {
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
        for (long long i = 0; i < (long long)ntraj; ++i) {
            const double* p = params_per_traj ? params + (size_t)i * np : params;
            solve_rosenbrock<N, F>(co, y0 + (size_t)i * N, tspan, nt, p, yall ? yall + (size_t)i * nt * N : nullptr,
                                   yfinal ? yfinal + (size_t)i * N : nullptr, status ? status + i : nullptr, nullptr, nullptr);
        }
        return (int)OK; }
    });
}

// fdjacobian (problem.rs:903-926) on its own: J is [n][n] row-major (dfdx[(m, n)] = J[m * dim + n]).
int oracle_fdjacobian(int rhs, const double* params, double t, const double* x, double* J) {
    return dispatch_rhs(rhs, [&](auto f) {
        using F = decltype(f);
        constexpr int N = F::N;
        double jac[N][N];
        uint32_t nfe = 0;
        fdjacobian<N, F>(t, x, params, jac, nfe);
        for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) J[r * N + c] = jac[r][c];
        return (int)OK;
    });
}

int oracle_rosenbrock_coeffs(int which, double* gamma, double* a, double* b, double* c) {
    if (which != 0 && which != 1) return ERR_INVALID_ARG;
    const RosenbrockCoeffs co = rosenbrock_coeffs(which);
    *gamma = co.gamma;
    for (int i = 0; i < 4; ++i) { b[i] = co.b[i]; for (int j = 0; j < 4; ++j) { a[4 * i + j] = co.a[i][j]; c[4 * i + j] = co.c[i][j]; } }
    return OK;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// libm hooks so the tests can compare the product's device pow/log10 against what the oracle (and Rust) call.
double oracle_pow(double x, double y) { return std::pow(x, y); }
double oracle_log10(double x) { return std::log10(x); }
double oracle_exp10(double x) { return exp10(x); }
double oracle_sin(double x) { return std::sin(x); }
double oracle_cos(double x) { return std::cos(x); }

}  // extern "C"
