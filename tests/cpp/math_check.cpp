// Host-side check of diffeq_b200/csrc/deq_math.cuh against the running glibc (the libm the reference calls).
// Build: g++ -O2 -ffp-contract=off -mfma tests/cpp/math_check.cpp -o /tmp/math_check -lm ; run: math_check <n>
// Prints "MISMATCHES <k>" (k must be 0).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <initializer_list>
#include "../../diffeq_b200/csrc/deq_math.cuh"
#include "../../diffeq_b200/csrc/deq_sincos.cuh"
static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double u01() { return (rnd() >> 11) * 0x1p-53; }
static inline bool same(double a, double b) { return deqm::asu64(a) == deqm::asu64(b) || (a != a && b != b); }
int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    const deqm::Tables T = deqm::host();
    long bad = 0;
    const double ys[] = {0.5, 0.2, 0.125, 1.0 / 3.0, 0.5, 1.0 / 6.0, 0.25, 1.0 / 8.0};
    for (long i = 0; i < n; ++i) {
        double x = exp((u01() * 2 - 1) * (i % 4 == 0 ? 700.0 : 46.0));
        if (i % 7 == 0) x = 1.0 + (u01() * 2 - 1) * 0.5 * pow(2.0, -(double)(i % 50));
        const double y = ys[i % 8];
        volatile double vx = x, vy = y;
        if (!same(pow(vx, vy), deqm::pow_glibc(x, y, T))) { if (bad < 5) printf("pow x=%a y=%a\n", x, y); ++bad; }
        double p = (u01() * 2 - 1) * (i % 3 == 0 ? 330.0 : 30.0);
        volatile double v10 = 10.0, vp = p;
        if (!same(pow(v10, vp), deqm::pow_glibc(10.0, p, T))) { if (bad < 5) printf("pow10 p=%a\n", p); ++bad; }
        if (!same(log10(vx), deqm::log10_glibc(x, T))) { if (bad < 5) printf("log10 x=%a\n", x); ++bad; }
        if (!same(log(vx), deqm::log_glibc(x, T))) { if (bad < 5) printf("log x=%a\n", x); ++bad; }
    }
    const double sp[] = {0.0, INFINITY, NAN, 1.0, 0x1p-1074, 0x1p-1030, 0x1p-1022, 1e300, 1e-300, 2.0, 4.0, 1e308, 0.5, 0x1.fffffffffffffp-1, 0x1.0000000000001p0};
    const double spy[] = {0.5, 0.2, 0.125, 1.0 / 3.0, -0.5, 0.0, 1.0, 2.0, 1e-70, 1e70, -1e70, INFINITY, -INFINITY, NAN, 700.0, -700.0};
    for (double x : sp) {
        for (double y : spy) {
            volatile double vx = x, vy = y;
            if (!same(pow(vx, vy), deqm::pow_glibc(x, y, T))) { printf("special pow x=%a y=%a glibc=%a mine=%a\n", x, y, pow(vx, vy), deqm::pow_glibc(x, y, T)); ++bad; }
        }
        volatile double vx = x;
        if (!same(log10(vx), deqm::log10_glibc(x, T))) { printf("special log10 x=%a\n", x); ++bad; }
        volatile double nx = -x;
        if (!same(log10(nx), deqm::log10_glibc(-x, T))) { printf("special log10 x=%a\n", -x); ++bad; }
    }
    // exp10 replica ("LLVM-folded" libm flavour of hinit)
    for (long i = 0; i < n; ++i) {
        const double p = (u01() * 2 - 1) * (i % 3 ? 30.0 : 320.0);
        volatile double vp = p;
        if (!same(exp10(vp), deqm::exp10_glibc(p, T))) { if (bad < 5) printf("exp10 p=%a\n", p); ++bad; }
    }
    for (double p : {0.0, -0.0, 1e-300, 308.0, 308.3, 309.0, -307.0, -308.0, -320.0, -324.0, -400.0, (double)INFINITY, -(double)INFINITY}) {
        volatile double vp = p;
        if (!same(exp10(vp), deqm::exp10_glibc(p, T))) { printf("exp10 special p=%a glibc=%a mine=%a\n", p, exp10(vp), deqm::exp10_glibc(p, T)); ++bad; }
    }
    // sin / cos replicas (pendulum right-hand side)
    const deqm::SinCosHost ST;
    for (long i = 0; i < n; ++i) {
        double x; const int mm = (int)(i % 7);
        if (mm == 0) x = (u01() * 2 - 1) * 0.9; else if (mm == 1) x = (u01() * 2 - 1) * 2.5; else if (mm == 2) x = (u01() * 2 - 1) * 100;
        else if (mm == 3) x = (u01() * 2 - 1) * 1e6; else if (mm == 4) x = (u01() * 2 - 1) * pow(2.0, -(double)(i % 60));
        else if (mm == 5) x = (u01() * 2 - 1) * 8; else x = (u01() * 2 - 1) * 1e8;
        volatile double vx = x;
        if (!same(sin(vx), deqm::sin_glibc(x, ST))) { if (bad < 5) printf("sin x=%a\n", x); ++bad; }
        if (!same(cos(vx), deqm::cos_glibc(x, ST))) { if (bad < 5) printf("cos x=%a\n", x); ++bad; }
        // the fused pair (pendulum right-hand side): sin of this argument, cos of the previous one
        static double xprev = 0.3;
        volatile double vp = xprev;
        double fs, fc;
        deqm::sin_cos_glibc(x, xprev, ST, fs, fc);
        if (!same(sin(vx), fs) || !same(cos(vp), fc)) { if (bad < 5) printf("sin_cos pair xs=%a xc=%a\n", x, xprev); ++bad; }
        xprev = x;
    }
    for (double x : {0.0, -0.0, 0.126, -0.126, 0.855469, -0.855469, 2.426265, -2.426265, 105414349.0, 1e-300, 3.141592653589793, 1.5707963267948966,
                     0x1p-26, 0x1p-27, 0x1.fffffffffffffp-27, 0x1.fffffffffffffp-28, 0.7853981633974483, 2.356194490192345, (double)INFINITY})
        for (double z : {0.0, -0.0, 0.126, 0.855469, 2.426265, 105414349.0, 1e-300, 6.283185307179586, 4.71238898038469, (double)INFINITY}) {
            volatile double vx = x, vz = z;
            double fs, fc;
            deqm::sin_cos_glibc(x, z, ST, fs, fc);
            if (!same(sin(vx), fs) || !same(cos(vz), fc)) { printf("sin_cos pair special xs=%a xc=%a\n", x, z); ++bad; }
        }
    for (double x : {0.0, -0.0, 0.126, -0.126, 0.855469, 2.426265, 105414349.0, 1e-300, 3.141592653589793, 1.5707963267948966, (double)INFINITY}) {
        volatile double vx = x;
        if (!same(sin(vx), deqm::sin_glibc(x, ST)) || !same(cos(vx), deqm::cos_glibc(x, ST))) { printf("sincos special x=%a\n", x); ++bad; }
    }
    // huge arguments: glibc's __branred (|x| >= 105414350), every binade up to DBL_MAX, and the range boundary
    for (long i = 0; i < n; ++i) {
        const uint64_t r = rnd();
        const uint64_t e = 0x419u + (r >> 52) % (0x7ffu - 0x419u);
        double x = deqm::asf64((r & 0x800fffffffffffffULL) | (e << 52));
        if (i % 16 == 0) x = (u01() * 2 - 1) * 4e8;
        if (i % 1000 == 1) x = 105414350.0 + (double)(i % 7) - 3.0;
        volatile double vx = x;
        if (!same(sin(vx), deqm::sin_glibc(x, ST))) { if (bad < 5) printf("sin huge x=%a glibc=%a mine=%a\n", x, sin(vx), deqm::sin_glibc(x, ST)); ++bad; }
        if (!same(cos(vx), deqm::cos_glibc(x, ST))) { if (bad < 5) printf("cos huge x=%a\n", x); ++bad; }
    }
    for (double x : {1e300, -1e300, 1e22, 0x1.fffffffffffffp1023, -0x1.fffffffffffffp1023, 105414350.0, 105414357.85197645, 0x1.921fb54442d18p+600,
                     (double)INFINITY, -(double)INFINITY, (double)NAN}) {
        volatile double vx = x;
        double fs, fc;
        deqm::sin_cos_glibc(x, x, ST, fs, fc);
        if (!same(sin(vx), deqm::sin_glibc(x, ST)) || !same(cos(vx), deqm::cos_glibc(x, ST)) || !same(sin(vx), fs) || !same(cos(vx), fc)) { printf("sincos huge special x=%a\n", x); ++bad; }
    }
    // FAST-mode step factor: relative error of pow_fast_pos against libm (not bit-exact by design)
    double worst = 0.0;
    for (long i = 0; i < n; ++i) {
        const double x = exp((u01() * 2 - 1) * (i % 4 == 0 ? 700.0 : 40.0));
        const double y = -0.5 / (double)(2 + i % 7);
        const double a = pow(x, y), b = deqm::pow_fast_pos(x, y, T);
        const double rel = fabs(a - b) / a;
        if (rel > worst) worst = rel;
    }
    printf("POW_FAST_MAX_REL %.3e\n", worst);
    if (!(worst < 1e-12)) ++bad;
    if (deqm::pow_fast_pos(0.0, -0.1, T) != INFINITY || deqm::pow_fast_pos(INFINITY, -0.1, T) != 0.0) { printf("pow_fast specials\n"); ++bad; }
    printf("MISMATCHES %ld\n", bad);
    return bad != 0;
}
