// deq_sincos.cuh — sin / cos with the SAME bits as glibc 2.39 (s_sin.c, IBM Accurate Mathematical Library, in the
// FMA-contracted form of the x86_64 `__sin_fma` / `__cos_fma` ifunc variants).
//
// Why: a user closure of the reference that calls f64::sin / f64::cos (e.g. the damped driven pendulum of BASELINE
// config 4) reaches glibc; CUDA's sin/cos differ by up to 2 ulp, which at rtol = 1e-12 changed the accepted-step count
// of 4 % of the trajectories.  With these replicas the built-in pendulum is bit-identical to the CPU oracle.
// Every fused multiply-add is explicit (placement from GCC 13's -mfma contraction of the published source, verified
// bit-for-bit against the running libm: tests/cpp/math_check.cpp); compile with -fmad=false / -ffp-contract=off.
// Domain: |x| < 105414350 (glibc's `reduce_sincos` range).  Larger arguments use glibc's Payne–Hanek `__branred`,
// which is not replicated: they fall back to the platform sin/cos (documented loss of bit parity, never hit by the
// built-in systems).
#pragma once
#include "deq_math.cuh"
#include "deq_sincos_table.h"

namespace deqm {

#if !defined(__CUDA_ARCH__)
namespace host_tables { static const double sincos_tab[440] = {DEQ_SINCOSTAB_INIT}; }
struct SinCosHost { DEQ_HD double at(int i) const { return host_tables::sincos_tab[i]; } };
#endif
#if defined(__CUDACC__)
static __device__ const double g_sincos_tab[440] = {DEQ_SINCOSTAB_INIT};
struct SinCosGlobal { __device__ __forceinline__ double at(int i) const { return __ldg(g_sincos_tab + i); } };
// Per-CTA copy in shared memory: the lookups are data dependent (one index per lane), which shared memory serves in a few
// wavefronts while the global path pays a tag lookup per distinct line.  Kernels call sincos_stage() once per CTA.
__shared__ double s_sincos_tab[440];
struct SinCosShared { __device__ __forceinline__ double at(int i) const { return s_sincos_tab[i]; } };
__device__ __forceinline__ void sincos_stage() {
    for (int i = threadIdx.x; i < 440; i += blockDim.x) s_sincos_tab[i] = g_sincos_tab[i];
}
#endif

namespace sc {
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ECEp-7, s3 = -0x1.A01A019DB08B8p-13,
                 s4 = 0x1.71DE27B9A7ED9p-19, s5 = -0x1.ADDFFC2FCDF59p-26;
constexpr double sn3 = -1.66666666666664880952546298448555E-01, sn5 = 8.33333214285722277379541354343671E-03;
constexpr double cs2 = 4.99999999999999999999950396842453E-01, cs4 = -4.16666666666664434524222570944589E-02,
                 cs6 = 1.38888874007937613028114285595617E-03;
constexpr double big = 0x1.8p45, hp0 = 0x1.921FB54442D18p0, hp1 = 0x1.1A62633145C07p-54;
constexpr double hpinv = 0x1.45F306DC9C883p-1, toint = 0x1.8p52;
constexpr double mp1 = 0x1.921FB58p0, mp2 = -0x1.DDE973Cp-27, pp3 = -0x1.CB3B398p-55, pp4 = -0x1.d747f23e32ed7p-83;
}  // namespace sc

DEQ_HD double copysign_(double a, double b) {
    return asf64((asu64(a) & 0x7fffffffffffffffULL) | (asu64(b) & 0x8000000000000000ULL));
}

// TAYLOR_SIN(xx, x, dx)
DEQ_HD double taylor_sin(double x, double dx) {
    const double xx = x * x;
    double p = fma_(xx, sc::s5, sc::s4);
    p = fma_(xx, p, sc::s3);
    p = fma_(xx, p, sc::s2);
    p = fma_(xx, p, sc::s1);
    const double t = fma_(xx, fma_(p, x, -(dx * 0.5)), dx);
    return t + x;
}

template <class ST>
DEQ_HD double do_sin(double x, double dx, const ST& tab) {
    const double xold = x;
    if (fabs(x) < 0.126) return taylor_sin(x, dx);
    if (x <= 0) dx = -dx;
    const double ax = fabs(x);
    const double u = ax + sc::big;
    x = ax - (u - sc::big);
    const double xx = x * x;
    const double s = x + fma_(x * xx, fma_(xx, sc::sn5, sc::sn3), dx);
    const double c = fma_(x, dx, xx * fma_(xx, fma_(xx, sc::cs6, sc::cs4), sc::cs2));
    const int k = (int)(uint32_t)asu64(u) * 4;
    const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
    const double cor = fma_(s, cs, fma_(-c, sn, fma_(s, ccs, ssn)));
    return copysign_(sn + cor, xold);
}

template <class ST>
DEQ_HD double do_cos(double x, double dx, const ST& tab) {
    if (x < 0) dx = -dx;
    const double ax = fabs(x);
    const double u = ax + sc::big;
    x = (ax - (u - sc::big)) + dx;
    const double xx = x * x;
    const double s = fma_(x * xx, fma_(xx, sc::sn5, sc::sn3), x);
    const double c = xx * fma_(xx, fma_(xx, sc::cs6, sc::cs4), sc::cs2);
    const int k = (int)(uint32_t)asu64(u) * 4;
    const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
    const double cor = fma_(-s, sn, fma_(-c, cs, fma_(-s, ssn, ccs)));
    return cs + cor;
}

// reduce_sincos: x = n*pi/2 + (a + da)
DEQ_HD int reduce_sincos(double x, double& a, double& da) {
    const double t = fma_(x, sc::hpinv, sc::toint);
    const double xn = t - sc::toint;
    const double y = fma_(-xn, sc::mp2, fma_(-xn, sc::mp1, x));
    const int n = (int)(uint32_t)asu64(t) & 3;
    const double t2 = fma_(-xn, sc::pp3, y);
    double db = fma_(-xn, sc::pp3, y - t2);
    const double b = fma_(-xn, sc::pp4, t2);
    db = db + fma_(-xn, sc::pp4, t2 - b);
    a = b; da = db;
    return n;
}

template <class ST>
DEQ_HD double do_sincos(double a, double da, int n, const ST& tab) {
    const double r = (n & 1) ? do_cos(a, da, tab) : do_sin(a, da, tab);
    return (n & 2) ? -r : r;
}

// do_sin and do_cos evaluated side by side and selected: every lane of a warp runs the same instruction stream whatever
// quadrant its argument fell into (on the GPU the two-way branch cost a warp both paths anyway, plus the divergence
// bookkeeping).  Each flavour's arithmetic is exactly glibc's do_sin / do_cos above; only the small-argument Taylor
// branch of do_sin stays a branch.
template <class ST>
DEQ_HD double sincos_dual(double a, double da, int n, const ST& tab) {
    const bool want_cos = (n & 1) != 0;
    const double ax = fabs(a);
    double r;
    if (!want_cos && ax < 0.126) {
        r = taylor_sin(a, da);
    } else {
        const double u = ax + sc::big;
        const double xr = ax - (u - sc::big);
        const int k = (int)(uint32_t)asu64(u) * 4;
        const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
        // do_sin flavour
        const double dxs = (a <= 0) ? -da : da;
        const double xxs = xr * xr;
        const double ss = xr + fma_(xr * xxs, fma_(xxs, sc::sn5, sc::sn3), dxs);
        const double csn = fma_(xr, dxs, xxs * fma_(xxs, fma_(xxs, sc::cs6, sc::cs4), sc::cs2));
        const double rs = copysign_(sn + fma_(ss, cs, fma_(-csn, sn, fma_(ss, ccs, ssn))), a);
        // do_cos flavour
        const double dxc = (a < 0) ? -da : da;
        const double xc = xr + dxc;
        const double xxc = xc * xc;
        const double sc_ = fma_(xc * xxc, fma_(xxc, sc::sn5, sc::sn3), xc);
        const double cc = xxc * fma_(xxc, fma_(xxc, sc::cs6, sc::cs4), sc::cs2);
        const double rc = cs + fma_(-sc_, sn, fma_(-cc, cs, fma_(-sc_, ssn, ccs)));
        r = want_cos ? rc : rs;
    }
    return r;
}

// Range dispatch of s_sin.c `__sin` / `__cos`: every range first produces (a, da, n) — cheap, a handful of operations —
// and then ONE sincos_dual evaluation follows.  The arithmetic of each range is exactly glibc's.
template <class ST>
DEQ_HD double sin_glibc(double x, const ST& tab) {
    const uint32_t k = (uint32_t)(asu64(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;
    if (k >= 0x419921FBu) return sin(x);  // outside reduce_sincos' range (see header)
    double a, da;
    int n;
    bool from_cos = false;
    if (k < 0x3feb6000u) { a = x; da = 0.0; n = 0; }
    else if (k < 0x400368fdu) { a = sc::hp0 - fabs(x); da = sc::hp1; n = 1; from_cos = true; }
    else n = reduce_sincos(x, a, da);
    const double r = sincos_dual(a, da, n, tab);
    if (from_cos) return copysign_(r, x);
    return (n & 2) ? -r : r;
}

template <class ST>
DEQ_HD double cos_glibc(double x, const ST& tab) {
    const uint32_t k = (uint32_t)(asu64(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e400000u) return 1.0;
    if (k >= 0x419921FBu) return cos(x);
    double a, da;
    int n;
    if (k < 0x3feb6000u) { a = x; da = 0.0; n = 1; }
    else if (k < 0x400368fdu) {
        const double y = sc::hp0 - fabs(x);
        a = y + sc::hp1;
        da = (y - a) + sc::hp1;
        n = 0;
    } else n = reduce_sincos(x, a, da) + 1;
    const double r = sincos_dual(a, da, n, tab);
    return (n & 2) ? -r : r;
}

}  // namespace deqm
