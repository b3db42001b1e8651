// tableaus.cuh — Butcher tableaus as compile-time constants for the device steppers.
//
// Mirrors /root/reference/src/ode/runge_kutta.rs:231-817 (`ButcherTableau::{feuler,midpoint,heun,rk21,rk23,rk4,
// rk45,dopri5,feh78}`).  Coefficients are written as the same f64 quotient expressions as the reference (each is one
// correctly rounded IEEE division).  Two sets exist:
//   REFERENCE — bug-for-bug what the crate computes with (rk21 weights transposed, rk23 b*[0] = 1/9, dopri5
//               c[3] = 0.75, feh78 row 5 / weights scrambled); used for parity against the oracle;
//   TEXTBOOK  — the corrected Heun–Euler, Bogacki–Shampine, Dormand–Prince and Fehlberg 7(8) coefficients.
// Layout: a[row][col] (strictly lower triangular is all that is read), b[s][0] = stepping weights,
// b[s][1] = error-comparison weights (reference `Weights::Adaptive` column 0 / column 1, runge_kutta.rs:114-117),
// c[s] nodes.  `ORDER_MIN` is `RKOrder::min()` (:43-55) — the controller exponent is 1/(ORDER_MIN+1).
// `FSAL` reproduces `is_first_same_as_last` (:185-193): last row of a == column 0 of b and c[S-1] == 1.
#pragma once

namespace deqt {

template <int S_>
struct Data {
    double a[S_][S_];
    double b[S_][2];
    double c[S_];
};

template <int S_>
constexpr bool fsal_of(const Data<S_>& d) {
    for (int col = 0; col < S_; ++col)
        if (d.a[S_ - 1][col] != d.b[col][0]) return false;
    return d.c[S_ - 1] == 1.0;
}

// ---- fixed-step (Weights::Explicit; only b[s][0] is meaningful) ------------------------------------------------
struct Feuler {  // runge_kutta.rs:238-249
    static constexpr int S = 1, ORDER_MIN = 1;
    static constexpr bool ADAPTIVE = false;
    static constexpr Data<1> data() { return {{{0.}}, {{1., 0.}}, {0.}}; }
};
struct Midpoint {  // :261-272
    static constexpr int S = 2, ORDER_MIN = 2;
    static constexpr bool ADAPTIVE = false;
    static constexpr Data<2> data() { return {{{0., 0.}, {0.5, 0.}}, {{0., 0.}, {1., 0.}}, {0., 0.5}}; }
};
struct Heun {  // :281-292
    static constexpr int S = 2, ORDER_MIN = 2;
    static constexpr bool ADAPTIVE = false;
    static constexpr Data<2> data() { return {{{0., 0.}, {1., 0.}}, {{0.5, 0.}, {0.5, 0.}}, {0., 1.}}; }
};
struct Rk4 {  // :370-383
    static constexpr int S = 4, ORDER_MIN = 4;
    static constexpr bool ADAPTIVE = false;
    static constexpr Data<4> data() {
        return {{{0., 0., 0., 0.}, {0.5, 0., 0., 0.}, {0., 0.5, 0., 0.}, {0., 0., 1., 0.}},
                {{1. / 6., 0.}, {1. / 3., 0.}, {1. / 3., 0.}, {1. / 6., 0.}},
                {0., 0.5, 0.5, 1.}};
    }
};

// ---- adaptive (Weights::Adaptive) ------------------------------------------------------------------------------
template <bool TEXTBOOK>
struct Rk21 {  // :294-305  Adaptive(2,1)
    static constexpr int S = 2, ORDER_MIN = 1;
    static constexpr bool ADAPTIVE = true;
    static constexpr Data<2> data() {
        // reference: Matrix2::new(0.5, 0.5, 1., 0.) is row-major => col0 = [0.5, 1], col1 = [0.5, 0]
        return TEXTBOOK ? Data<2>{{{0., 0.}, {1., 0.}}, {{0.5, 1.}, {0.5, 0.}}, {0., 1.}}
                        : Data<2>{{{0., 0.}, {1., 0.}}, {{0.5, 0.5}, {1., 0.}}, {0., 1.}};
    }
};
template <bool TEXTBOOK>
struct Rk23 {  // :319-357  Adaptive(2,3)
    static constexpr int S = 4, ORDER_MIN = 2;
    static constexpr bool ADAPTIVE = true;
    static constexpr Data<4> data() {
        return {{{0., 0., 0., 0.}, {0.5, 0., 0., 0.}, {0., 0.75, 0., 0.}, {2. / 9., 1. / 3., 4. / 9., 0.}},
                {{7. / 24., TEXTBOOK ? 2. / 9. : 1. / 9.}, {0.25, 1. / 3.}, {1. / 3., 4. / 9.}, {0.125, 0.}},
                {0., 0.5, 0.75, 1.}};
    }
};
struct Rk45 {  // Fehlberg 4(5) :401-463  Adaptive(4,5) — identical in both sets
    static constexpr int S = 6, ORDER_MIN = 4;
    static constexpr bool ADAPTIVE = true;
    static constexpr Data<6> data() {
        return {{{0., 0., 0., 0., 0., 0.},
                 {0.25, 0., 0., 0., 0., 0.},
                 {0.09375, 0.28125, 0., 0., 0., 0.},
                 {1932. / 2197., -7200. / 2197., 7296. / 2197., 0., 0., 0.},
                 {439. / 216., -8., 3680. / 513., -845. / 4104., 0., 0.},
                 {-8. / 27., 2., -3544. / 2565., 1859. / 4104., -0.275, 0.}},
                {{25. / 216., 16. / 135.}, {0., 0.}, {1408. / 2565., 6656. / 12825.}, {2197. / 4104., 28561. / 56430.},
                 {-0.2, -0.18}, {0., 2. / 55.}},
                {0., 0.25, 0.375, 12. / 13., 1., 0.5}};
    }
};
template <bool TEXTBOOK>
struct Dopri5 {  // :478-562  Adaptive(5,4)
    static constexpr int S = 7, ORDER_MIN = 4;
    static constexpr bool ADAPTIVE = true;
    static constexpr Data<7> data() {
        return {{{0., 0., 0., 0., 0., 0., 0.},
                 {0.2, 0., 0., 0., 0., 0., 0.},
                 {0.075, 0.225, 0., 0., 0., 0., 0.},
                 {44. / 45., -56. / 15., 32. / 9., 0., 0., 0., 0.},
                 {19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729., 0., 0., 0.},
                 {9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656., 0., 0.},
                 {35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 0.}},
                {{35. / 384., 5179. / 57600.}, {0., 0.}, {500. / 1113., 7571. / 16695.}, {125. / 192., 393. / 640.},
                 {-2187. / 6784., -92097. / 339200.}, {11. / 84., 187. / 2100.}, {0., 0.025}},
                {0., 0.2, 0.3, TEXTBOOK ? 4. / 5. : 0.75, 8. / 9., 1., 1.}};
    }
};
template <bool TEXTBOOK>
struct Feh78 {  // :582-816  Adaptive(7,8)
    static constexpr int S = 13, ORDER_MIN = 7;
    static constexpr bool ADAPTIVE = true;
    static constexpr Data<13> data() {
        Data<13> d = {{{0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.},
                       {2. / 27., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.},
                       {1. / 36., 1. / 12., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.},
                       {1. / 24., 0., 0.125, 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.},
                       {5. / 12., 0., -25. / 16., 25. / 16., 0., 0., 0., 0., 0., 0., 0., 0., 0.},
                       {0.5, 0., 0., 0.75, 0.2, 0., 0., 0., 0., 0., 0., 0., 0.},
                       {-25. / 108., 0., 0., 125. / 108., -65. / 27., 125. / 54., 0., 0., 0., 0., 0., 0., 0.},
                       {31. / 300., 0., 0., 0., 61. / 225., -2. / 9., 13. / 900., 0., 0., 0., 0., 0., 0.},
                       {2., 0., 0., -53. / 6., 704. / 45., -107. / 9., 67. / 90., 3., 0., 0., 0., 0., 0.},
                       {-91. / 108., 0., 0., 23. / 108., -976. / 135., 311. / 54., -19. / 60., 17. / 6., -1. / 12., 0., 0., 0., 0.},
                       {2383. / 4100., 0., 0., -341. / 164., 4496. / 1025., -301. / 82., 2133. / 4100., 45. / 82., 45. / 164., 18. / 41., 0., 0., 0.},
                       {3. / 205., 0., 0., 0., 0., -6. / 41., -3. / 205., -3. / 41., 3. / 41., 6. / 41., 0., 0., 0.},
                       {-1777. / 4100., 0., 0., -341. / 164., 4496. / 1025., -289. / 82., 2193. / 4100., 51. / 82., 33. / 164., 12. / 41., 0., 1., 0.}},
                      // reference weights as stored (from_row_slice U13xU2): col0 = [41/840, 0 x9, 34/105, 34/105, 9/35]
                      {{41. / 840., 9. / 35.}, {0., 9. / 35.}, {0., 9. / 35.}, {0., 9. / 280.}, {0., 9. / 280.},
                       {0., 9. / 280.}, {0., 9. / 280.}, {0., 41. / 840.}, {0., 0.}, {0., 0.}, {34. / 105., 41. / 840.},
                       {34. / 105., 0.}, {9. / 35., 41. / 840.}},
                      {0., 2. / 27., 1. / 9., 1. / 6., 5. / 12., 0.5, 5. / 6., 1. / 6., 2. / 3., 1. / 3., 1., 0., 1.}};
        if (TEXTBOOK) {
            d.a[5][0] = 1. / 20.; d.a[5][3] = 1. / 4.; d.a[5][4] = 1. / 5.;
            const double b7[13] = {41. / 840., 0., 0., 0., 0., 34. / 105., 9. / 35., 9. / 35., 9. / 280., 9. / 280., 41. / 840., 0., 0.};
            const double b8[13] = {0., 0., 0., 0., 0., 34. / 105., 9. / 35., 9. / 35., 9. / 280., 9. / 280., 0., 41. / 840., 41. / 840.};
            for (int s = 0; s < 13; ++s) { d.b[s][0] = b7[s]; d.b[s][1] = b8[s]; }
        }
        return d;
    }
};

}  // namespace deqt
