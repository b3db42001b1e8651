// deq_sincos.cuh — sin / cos with the SAME bits as glibc 2.39 (s_sin.c, IBM Accurate Mathematical Library, in the
// FMA-contracted form of the x86_64 `__sin_fma` / `__cos_fma` ifunc variants).
//
// Why: a user closure of the reference that calls f64::sin / f64::cos (e.g. the damped driven pendulum of BASELINE
// config 4) reaches glibc; CUDA's sin/cos differ by up to 2 ulp, which at rtol = 1e-12 changed the accepted-step count
// of 4 % of the trajectories.  With these replicas the built-in pendulum is bit-identical to the CPU oracle.
// Every fused multiply-add is explicit (placement from GCC 13's -mfma contraction of the published source, verified
// bit-for-bit against the running libm: tests/cpp/math_check.cpp); compile with -fmad=false / -ffp-contract=off.
// Domain: every double.  |x| < 105414350 is glibc's `reduce_sincos` range; larger finite arguments go through `branred`
// below, glibc's Payne–Hanek style reduction against 1800 bits of 2/pi (branred.c; the libm of this image runs it WITHOUT
// fused multiply-adds — the one copy in libm.so.6 is plain SSE2 — so it is written with separate roundings); inf / NaN
// give x / x.
// ---- third-party notice -------------------------------------------------------------------------------------------
// The algorithms and the table restated in this file are those of the GNU C Library 2.39 (sysdeps/ieee754/dbl-64/s_sin.c,
// sincostab.c, usncs.h, branred.c: the IBM Accurate Mathematical Library, written by International Business Machines Corp.).
// Copyright (C) 2001-2024 Free Software Foundation, Inc.  The GNU C Library is free software, distributed under the terms
// of the GNU Lesser General Public License, version 2.1 or (at your option) any later version
// (https://www.gnu.org/licenses/lgpl-2.1.html).  This file is a derived work for the purpose of bit-compatibility with glibc
// and is provided under the same terms, WITHOUT ANY WARRANTY.
// --------------------------------------------------------------------------------------------------------------------
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

// Scalar constants of s_sin.c.  In device code they live in a __constant__ array (one LDCU.64 per use) instead of being
// materialised as 64-bit immediates (two UMOV issue slots each - 17 per call before, measured on BASELINE config 4).
#define DEQ_SCK_LIST(X) \
    X(s1, -0x1.5555555555555p-3) X(s2, 0x1.1111111110ECEp-7) X(s3, -0x1.A01A019DB08B8p-13) X(s4, 0x1.71DE27B9A7ED9p-19) \
    X(s5, -0x1.ADDFFC2FCDF59p-26) X(sn3, -1.66666666666664880952546298448555E-01) X(sn5, 8.33333214285722277379541354343671E-03) \
    X(cs2, 4.99999999999999999999950396842453E-01) X(cs4, -4.16666666666664434524222570944589E-02) \
    X(cs6, 1.38888874007937613028114285595617E-03) X(big, 0x1.8p45) X(hp0, 0x1.921FB54442D18p0) X(hp1, 0x1.1A62633145C07p-54) \
    X(hpinv, 0x1.45F306DC9C883p-1) X(toint, 0x1.8p52) X(mp1, 0x1.921FB58p0) X(mp2, -0x1.DDE973Cp-27) X(pp3, -0x1.CB3B398p-55) \
    X(pp4, -0x1.d747f23e32ed7p-83)
namespace sck {
enum : int {
#define X(name, lit) name,
    DEQ_SCK_LIST(X)
#undef X
    COUNT
};
}  // namespace sck
#if defined(__CUDACC__)
static __constant__ double c_sck[sck::COUNT] = {
#define X(name, lit) lit,
    DEQ_SCK_LIST(X)
#undef X
};
#endif
namespace sc {
DEQ_HD double sck_get(int i, double lit) {
#if defined(__CUDA_ARCH__)
    (void)lit;
    return c_sck[i];
#else
    (void)i;
    return lit;
#endif
}
#define X(name, lit) DEQ_HD double name##_() { return sck_get(sck::name, lit); }
DEQ_SCK_LIST(X)
#undef X
}  // namespace sc

DEQ_HD double copysign_(double a, double b) {
    return asf64((asu64(a) & 0x7fffffffffffffffULL) | (asu64(b) & 0x8000000000000000ULL));
}

// ---- branred.c: x = n*pi/2 + (a + aa) for |x| >= 105414350 -------------------------------------------------------------
// toverp[i] = the i-th 24-bit digit of 2/pi.  The argument is scaled by 2^-600 and split into two 26-bit halves; each half is
// multiplied by the six digits of 2/pi that matter at its magnitude (scaled by `gor` = 2^(576 - 24k)), the integer parts are
// taken off with the 1.5 * 2^52 trick, and the fractional parts are summed in double-double.  Every operation is one IEEE
// rounding in glibc's order.  Cold path (a state of 1e8 rad and more): out of line on the device.
#define DEQ_TOVERP_INIT \
    10680707.0, 7228996.0, 1387004.0, 2578385.0, 16069853.0, 12639074.0, 9804092.0, 4427841.0, 16666979.0, 11263675.0, \
    12935607.0, 2387514.0, 4345298.0, 14681673.0, 3074569.0, 13734428.0, 16653803.0, 1880361.0, 10960616.0, 8533493.0, \
    3062596.0, 8710556.0, 7349940.0, 6258241.0, 3772886.0, 3769171.0, 3798172.0, 8675211.0, 12450088.0, 3874808.0, \
    9961438.0, 366607.0, 15675153.0, 9132554.0, 7151469.0, 3571407.0, 2607881.0, 12013382.0, 4155038.0, 6285869.0, \
    7677882.0, 13102053.0, 15825725.0, 473591.0, 9065106.0, 15363067.0, 6271263.0, 9264392.0, 5636912.0, 4652155.0, \
    7056368.0, 13614112.0, 10155062.0, 1944035.0, 9527646.0, 15080200.0, 6658437.0, 6231200.0, 6832269.0, 16767104.0, \
    5075751.0, 3212806.0, 1398474.0, 7579849.0, 6349435.0, 12618859.0, 4703257.0, 12806093.0, 14477321.0, 2786137.0, \
    12875403.0, 9837734.0, 14528324.0, 13719321.0, 343717.0
#if !defined(__CUDA_ARCH__)
namespace host_tables { static const double toverp_tab[75] = {DEQ_TOVERP_INIT}; }
#endif
#if defined(__CUDACC__)
static __device__ const double g_toverp_tab[75] = {DEQ_TOVERP_INIT};
#endif
DEQ_HD double toverp_at(int i) {
#if defined(__CUDA_ARCH__)
    return g_toverp_tab[i];
#else
    return host_tables::toverp_tab[i];
#endif
}

DEQ_HD void branred_half(double xh, double& b, double& bb, double& sum_out) {
    const double big = 0x1.8p52, big1 = 0x1.8p54, tm24 = 0x1p-24;
    int k = (int)((asu64(xh) >> 52) & 2047u);
    k = (k - 450) / 24;
    if (k < 0) k = 0;
    double gor = asf64(asu64(0x1p576) - ((uint64_t)(k * 24) << 52));
    double r[6];
    for (int i = 0; i < 6; ++i) { r[i] = xh * toverp_at(k + i) * gor; gor = gor * tm24; }
    double sum = 0.0, s, t = 0.0;
    for (int i = 0; i < 3; ++i) { s = (r[i] + big) - big; sum = sum + s; r[i] = r[i] - s; }
    for (int i = 0; i < 6; ++i) t = t + r[5 - i];
    bb = (((((r[0] - t) + r[1]) + r[2]) + r[3]) + r[4]) + r[5];
    s = (t + big) - big;
    sum = sum + s;
    t = t - s;
    b = t + bb;
    bb = (t - b) + bb;
    s = (sum + big1) - big1;
    sum_out = sum - s;
}

#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
int branred(double x, double& a, double& aa) {
    const double split = 134217729.0;   // 2^27 + 1
    x = x * 0x1p-600;
    double t = x * split;
    const double x1 = t - (t - x), x2 = x - x1;
    double b1, bb1, sum1, b2, bb2, sum2;
    branred_half(x1, b1, bb1, sum1);
    branred_half(x2, b2, bb2, sum2);
    double sum = sum1 + sum2;
    double b = b1 + b2;
    double bb = (fabs(b1) > fabs(b2)) ? (b1 - b) + b2 : (b2 - b) + b1;
    if (b > 0.5) { b = b - 1.0; sum = sum + 1.0; }
    else if (b < -0.5) { b = b + 1.0; sum = sum - 1.0; }
    double s = b + (bb + bb1 + bb2);
    t = ((b - s) + bb) + (bb1 + bb2);
    b = s * split;
    const double t1 = b - (b - s), t2 = s - t1;
    b = s * 0x1.921FB54442D18p0;   // (branred.h: mp1 + mp2 == hp0 exactly; its mp2 is not the mp2 of reduce_sincos)
    bb = (((t1 * 0x1.921FB58p0 - b) + t1 * -0x1.DDE974p-27) + t2 * 0x1.921FB58p0) +
         (t2 * -0x1.DDE974p-27 + s * 0x1.1A62633145C07p-54 + t * 0x1.921FB54442D18p0);
    s = b + bb;
    t = (b - s) + bb;
    a = s;
    aa = t;
    return ((int)sum) & 3;
}

// TAYLOR_SIN(xx, x, dx)
DEQ_HD double taylor_sin(double x, double dx) {
    const double xx = x * x;
    double p = fma_(xx, sc::s5_(), sc::s4_());
    p = fma_(xx, p, sc::s3_());
    p = fma_(xx, p, sc::s2_());
    p = fma_(xx, p, sc::s1_());
    const double t = fma_(xx, fma_(p, x, -(dx * 0.5)), dx);
    return t + x;
}

template <class ST>
DEQ_HD double do_sin(double x, double dx, const ST& tab) {
    const double xold = x;
    if (fabs(x) < 0.126) return taylor_sin(x, dx);
    if (x <= 0) dx = -dx;
    const double ax = fabs(x);
    const double u = ax + sc::big_();
    x = ax - (u - sc::big_());
    const double xx = x * x;
    const double s = x + fma_(x * xx, fma_(xx, sc::sn5_(), sc::sn3_()), dx);
    const double c = fma_(x, dx, xx * fma_(xx, fma_(xx, sc::cs6_(), sc::cs4_()), sc::cs2_()));
    const int k = (int)(uint32_t)asu64(u) * 4;
    const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
    const double cor = fma_(s, cs, fma_(-c, sn, fma_(s, ccs, ssn)));
    return copysign_(sn + cor, xold);
}

template <class ST>
DEQ_HD double do_cos(double x, double dx, const ST& tab) {
    if (x < 0) dx = -dx;
    const double ax = fabs(x);
    const double u = ax + sc::big_();
    x = (ax - (u - sc::big_())) + dx;
    const double xx = x * x;
    const double s = fma_(x * xx, fma_(xx, sc::sn5_(), sc::sn3_()), x);
    const double c = xx * fma_(xx, fma_(xx, sc::cs6_(), sc::cs4_()), sc::cs2_());
    const int k = (int)(uint32_t)asu64(u) * 4;
    const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
    const double cor = fma_(-s, sn, fma_(-c, cs, fma_(-s, ssn, ccs)));
    return cs + cor;
}

// reduce_sincos: x = n*pi/2 + (a + da)
DEQ_HD int reduce_sincos(double x, double& a, double& da) {
    const double t = fma_(x, sc::hpinv_(), sc::toint_());
    const double xn = t - sc::toint_();
    const double y = fma_(-xn, sc::mp2_(), fma_(-xn, sc::mp1_(), x));
    const int n = (int)(uint32_t)asu64(t) & 3;
    const double t2 = fma_(-xn, sc::pp3_(), y);
    double db = fma_(-xn, sc::pp3_(), y - t2);
    const double b = fma_(-xn, sc::pp4_(), t2);
    db = db + fma_(-xn, sc::pp4_(), t2 - b);
    a = b; da = db;
    return n;
}

template <class ST>
DEQ_HD double do_sincos(double a, double da, int n, const ST& tab) {
    const double r = (n & 1) ? do_cos(a, da, tab) : do_sin(a, da, tab);
    return (n & 2) ? -r : r;
}

// |x| >= 105414350 (s_sin.c: `n = __branred (x, &a, &da); retval = do_sincos (a, da, n)`, n + 1 for the cosine), inf / NaN.
template <class ST>
DEQ_HD double sincos_large(double x, int quadrant_offset, const ST& tab) {
    if (((uint32_t)(asu64(x) >> 32) & 0x7fffffffu) >= 0x7ff00000u) return x / x;
    double a, da;
    const int n = branred(x, a, da);
    return do_sincos(a, da, n + quadrant_offset, tab);
}

DEQ_HD double xor_sign_(double v, uint64_t signbit_source) {   // v with its sign flipped iff bit 63 of signbit_source is set
    return asf64(asu64(v) ^ (signbit_source & 0x8000000000000000ULL));
}

// do_sin OR do_cos through ONE instruction stream: the operands are selected, not the results (17 FP64 operations instead of
// 13 + 13 for the two flavours side by side; the small-argument Taylor branch of do_sin is still evaluated alongside).
// With (A, AA, B, BB) = (sn, ssn, cs, ccs) for the sine and (cs, ccs, sn, ssn) for the cosine both results are
//     A + fma(s', B, fma(-c, A, fma(s', BB, AA)))            s' = s (sine), -s (cosine)
// and with X = xr (sine), xr + dx (cosine), xx = X X, P = fma(xx, sn5, sn3), C = xx fma(xx, fma(xx, cs6, cs4), cs2):
//     sine:   s = xr + fma(X xx, P, dx),   c = fma(xr, dx, C)
//     cosine: s =      fma(X xx, P, X),    c = C
// The cosine's forms are obtained from the sine's instructions bit for bit: s = -0.0 + f (x + -0.0 == x for EVERY x, -0.0
// and +0.0 included) and c = fma(xr, 0.0, C) (xr is finite, so the product is +-0, and C >= +0 is not changed by adding it:
// C is xx >= +0 times a polynomial close to 1/2).  dx is da with the sign of a transferred on (see sincos_dual2 below for why
// `x <= 0` and `x < 0` of do_sin / do_cos need not be told apart).
template <class ST>
DEQ_HD double sincos_unified(double a, double da, bool want_cos, const ST& tab) {
    const double ax = fabs(a);
    const double u = ax + sc::big_();
    const double xr = ax - (u - sc::big_());
    const int k = (int)(uint32_t)asu64(u) * 4;
    const int ka = want_cos ? k + 2 : k, kb = want_cos ? k : k + 2;
    const double A = tab.at(ka), AA = tab.at(ka + 1), B = tab.at(kb), BB = tab.at(kb + 1);
    const double dx = xor_sign_(da, asu64(a));
    const double xc = xr + dx;
    const double X = want_cos ? xc : xr;
    const double xx = X * X;
    const double P = fma_(xx, sc::sn5_(), sc::sn3_());
    const double C = xx * fma_(xx, fma_(xx, sc::cs6_(), sc::cs4_()), sc::cs2_());
    const double f = fma_(X * xx, P, want_cos ? xc : dx);
    const double s = (want_cos ? -0.0 : xr) + f;
    const double c = fma_(xr, want_cos ? 0.0 : dx, C);
    const double sp = xor_sign_(s, want_cos ? 0x8000000000000000ULL : 0ULL);
    const double res = A + fma_(sp, B, fma_(-c, A, fma_(sp, BB, AA)));
    const double rt = taylor_sin(a, da);   // do_sin's small-argument branch (reads the original da)
    // sine: copysign(res, a), or the Taylor value for |a| < 0.126; cosine: res
    const double rs = ax < 0.126 ? rt : copysign_(res, a);
    return want_cos ? res : rs;
}

// do_sin and do_cos evaluated side by side and selected: every lane of a warp runs the same instruction stream whatever
// quadrant its argument fell into (on the GPU the two-way branch cost a warp both paths anyway, plus the divergence
// bookkeeping).  Each flavour's arithmetic is exactly glibc's do_sin / do_cos above, the small-argument Taylor branch of
// do_sin included.
template <class ST>
DEQ_HD double sincos_dual(double a, double da, int n, const ST& tab) {
    const bool want_cos = (n & 1) != 0;
    const double ax = fabs(a);
    const double u = ax + sc::big_();
    const double xr = ax - (u - sc::big_());
    const int k = (int)(uint32_t)asu64(u) * 4;
    const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
    // do_sin flavour
    const double dxs = (a <= 0) ? -da : da;
    const double xxs = xr * xr;
    const double ss = xr + fma_(xr * xxs, fma_(xxs, sc::sn5_(), sc::sn3_()), dxs);
    const double csn = fma_(xr, dxs, xxs * fma_(xxs, fma_(xxs, sc::cs6_(), sc::cs4_()), sc::cs2_()));
    const double rs = copysign_(sn + fma_(ss, cs, fma_(-csn, sn, fma_(ss, ccs, ssn))), a);
    // do_cos flavour
    const double dxc = (a < 0) ? -da : da;
    const double xc = xr + dxc;
    const double xxc = xc * xc;
    const double sc_ = fma_(xc * xxc, fma_(xxc, sc::sn5_(), sc::sn3_()), xc);
    const double cc = xxc * fma_(xxc, fma_(xxc, sc::cs6_(), sc::cs4_()), sc::cs2_());
    const double rc = cs + fma_(-sc_, sn, fma_(-cc, cs, fma_(-sc_, ssn, ccs)));
#ifdef DEQ_SINCOS_BRANCHY
    if (!want_cos && ax < 0.126) return taylor_sin(a, da);
    return want_cos ? rc : rs;
#else
    const double rt = taylor_sin(a, da);   // do_sin's small-argument branch, evaluated alongside and selected
    return want_cos ? rc : (ax < 0.126 ? rt : rs);
#endif
}

// Range dispatch of s_sin.c `__sin` / `__cos`: every range first produces (a, da, n) and then ONE sincos_dual evaluation
// follows.  The arithmetic of each range is exactly glibc's.  The three ordinary ranges are evaluated side by side and
// selected (DEQ_SINCOS_BRANCHY restores the branches): a warp's lanes land in different ranges almost always (theta in
// [-pi, pi], Omega t in [0, 6.7] on BASELINE config 4), so the branches cost a warp every path plus the divergence bookkeeping.
template <class ST>
DEQ_HD double sin_glibc(double x, const ST& tab) {
    const uint32_t k = (uint32_t)(asu64(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;
    if (k >= 0x419921FBu) return sincos_large(x, 0, tab);  // outside reduce_sincos' range: branred, or x / x for inf and NaN
    double a, da;
    int n;
#ifdef DEQ_SINCOS_BRANCHY
    bool from_cos = false;
    if (k < 0x3feb6000u) { a = x; da = 0.0; n = 0; }
    else if (k < 0x400368fdu) { a = sc::hp0_() - fabs(x); da = sc::hp1_(); n = 1; from_cos = true; }
    else n = reduce_sincos(x, a, da);
#else
    const bool ra = k < 0x3feb6000u, rb = k < 0x400368fdu, from_cos = rb && !ra;
    n = reduce_sincos(x, a, da);
    const double ab = sc::hp0_() - fabs(x);
    a = ra ? x : rb ? ab : a;
    da = ra ? 0.0 : rb ? sc::hp1_() : da;
    n = ra ? 0 : rb ? 1 : n;
#endif
#ifdef DEQ_SINCOS_BRANCHY
    const double r = sincos_dual(a, da, n, tab);
    if (from_cos) return copysign_(r, x);
    return (n & 2) ? -r : r;
#else
    const double r = sincos_unified(a, da, (n & 1) != 0, tab);
    // middle range (taken through do_cos): copysign(r, x); otherwise negate in quadrants 2 and 3
    return from_cos ? copysign_(r, x) : xor_sign_(r, (uint64_t)(n & 2) << 62);
#endif
}

template <class ST>
DEQ_HD double cos_glibc(double x, const ST& tab) {
    const uint32_t k = (uint32_t)(asu64(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e400000u) return 1.0;
    if (k >= 0x419921FBu) return sincos_large(x, 1, tab);
    double a, da;
    int n;
#ifdef DEQ_SINCOS_BRANCHY
    if (k < 0x3feb6000u) { a = x; da = 0.0; n = 1; }
    else if (k < 0x400368fdu) {
        const double y = sc::hp0_() - fabs(x);
        a = y + sc::hp1_();
        da = (y - a) + sc::hp1_();
        n = 0;
    } else n = reduce_sincos(x, a, da) + 1;
#else
    const bool ra = k < 0x3feb6000u, rb = k < 0x400368fdu;
    n = reduce_sincos(x, a, da) + 1;
    const double y = sc::hp0_() - fabs(x);
    const double ab = y + sc::hp1_();
    const double dab = (y - ab) + sc::hp1_();
    a = ra ? x : rb ? ab : a;
    da = ra ? 0.0 : rb ? dab : da;
    n = ra ? 1 : rb ? 0 : n;
#endif
#ifdef DEQ_SINCOS_BRANCHY
    const double r = sincos_dual(a, da, n, tab);
    return (n & 2) ? -r : r;
#else
    return xor_sign_(sincos_unified(a, da, (n & 1) != 0, tab), (uint64_t)(n & 2) << 62);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// sin(xs) and cos(xc) in ONE straight-line evaluation (the right-hand side of the driven pendulum needs exactly this pair
// thirteen times per step attempt of Fehlberg 7(8)).  Same operations in the same order as sin_glibc / cos_glibc above for
// each argument — verified bit for bit on the host against the running libm (tests/cpp/math_check.cpp) — but shaped for the
// GPU's issue port (every instruction, FP64 or not, costs at least one issue slot):
//   * one range test for the pair: both arguments in [2^-26, 105414350) (resp. 2^-27 for cos); anything else goes through
//     the two general routines;
//   * `if (x <= 0) dx = -dx` (do_sin) and `if (x < 0) dx = -dx` (do_cos) are a sign transfer, an XOR on the high word
//     instead of a compare and two selects each.  The two conditions differ only for x == +-0, and there the result does
//     not depend on the sign of dx: do_sin returns its Taylor branch (which reads the original dx) for |x| < 0.126, and
//     do_cos at table entry 0 (sn = ssn = 0) multiplies every odd power of dx by zero;
//   * negations by (n & 2) are XORs on the sign bit instead of an FP64 negate and two selects.
// ---------------------------------------------------------------------------------------------------------------

template <class ST>
DEQ_HD double sincos_dual2(double a, double da, bool want_cos, const ST& tab) {
    const double ax = fabs(a);
    const double u = ax + sc::big_();
    const double xr = ax - (u - sc::big_());
    const int k = (int)(uint32_t)asu64(u) * 4;
    const double sn = tab.at(k), ssn = tab.at(k + 1), cs = tab.at(k + 2), ccs = tab.at(k + 3);
    const double dx = xor_sign_(da, asu64(a));
    // do_sin flavour
    const double xxs = xr * xr;
    const double ss = xr + fma_(xr * xxs, fma_(xxs, sc::sn5_(), sc::sn3_()), dx);
    const double csn = fma_(xr, dx, xxs * fma_(xxs, fma_(xxs, sc::cs6_(), sc::cs4_()), sc::cs2_()));
    const double rs = copysign_(sn + fma_(ss, cs, fma_(-csn, sn, fma_(ss, ccs, ssn))), a);
    // do_cos flavour
    const double xc = xr + dx;
    const double xxc = xc * xc;
    const double sc_ = fma_(xc * xxc, fma_(xxc, sc::sn5_(), sc::sn3_()), xc);
    const double cc = xxc * fma_(xxc, fma_(xxc, sc::cs6_(), sc::cs4_()), sc::cs2_());
    const double rc = cs + fma_(-sc_, sn, fma_(-cc, cs, fma_(-sc_, ssn, ccs)));
    const double rt = taylor_sin(a, da);   // do_sin's small-argument branch
    return want_cos ? rc : (ax < 0.126 ? rt : rs);
}

template <class ST>
DEQ_HD void sin_cos_glibc(double xs, double xc, const ST& tab, double& so, double& co) {
    const uint32_t ks = (uint32_t)(asu64(xs) >> 32) & 0x7fffffffu, kc = (uint32_t)(asu64(xc) >> 32) & 0x7fffffffu;
    if ((ks - 0x3e500000u) >= (0x419921FBu - 0x3e500000u) || (kc - 0x3e400000u) >= (0x419921FBu - 0x3e400000u)) {
        so = sin_glibc(xs, tab);
        co = cos_glibc(xc, tab);
        return;
    }
    {   // __sin
        const bool ra = ks < 0x3feb6000u, rb = ks < 0x400368fdu;
        double a, da;
        int n = reduce_sincos(xs, a, da);
        const double ab = sc::hp0_() - fabs(xs);
        a = ra ? xs : rb ? ab : a;
        da = ra ? 0.0 : rb ? sc::hp1_() : da;
        n = ra ? 0 : rb ? 1 : n;
        const double r = sincos_unified(a, da, (n & 1) != 0, tab);
        // middle range (taken through do_cos): copysign(r, x); otherwise negate in quadrants 2 and 3
        so = (rb && !ra) ? copysign_(r, xs) : xor_sign_(r, (uint64_t)(n & 2) << 62);
    }
    {   // __cos
        const bool ra = kc < 0x3feb6000u, rb = kc < 0x400368fdu;
        double a, da;
        int n = reduce_sincos(xc, a, da) + 1;
        const double y = sc::hp0_() - fabs(xc);
        const double ab = y + sc::hp1_();
        const double dab = (y - ab) + sc::hp1_();
        a = ra ? xc : rb ? ab : a;
        da = ra ? 0.0 : rb ? dab : da;
        n = ra ? 1 : rb ? 0 : n;
        const double r = sincos_unified(a, da, (n & 1) != 0, tab);
        co = xor_sign_(r, (uint64_t)(n & 2) << 62);
    }
}

}  // namespace deqm
