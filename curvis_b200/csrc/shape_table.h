// shape_table.h — piecewise-polynomial table of the Interstellar (DNEG) shape function used by
// CURVIS_PRECISION_F64_FAST (render_f64_fast.cu: FastInterstellar).
//
// InterstellarMetric::r / r_derivative (reference src/metrics.rs:461-485) need, per Euler step,
//     F(x) = x atan x - ln(1 + x^2)/2      (r  = rho + m F)
//     G(x) = (2/pi) atan x = (2/pi) F'(x)  (r' = sign(l) G),          x = 2(|l| - a)/(pi m) > 0,
// two library transcendentals worth ~80 fp64-pipe instructions — two thirds of that metric's step.
// Both are parameter-free functions of x, so ONE table serves every metric setting:
// x in [2^kShapeTabEmin, 2^kShapeTabEmax) is cut into 2^kShapeTabK equal intervals per binade (the
// interval index is a shift of x's high word, its midpoint c a mask), and on each interval F and G
// are degree-5 polynomials in t = x - c (exact subtraction).  Coefficients: Chebyshev interpolation
// evaluated in x87 long double on the host (shape_table.cpp), rounded once to double; approximation
// error < 1e-19, so the result is the rounding of the Horner evaluation: <= ~1 ulp, the class of the
// CUDA library's atan/log (tests/test_gpu_fast64.py checks both against long double).
// Outside the table range the kernel calls atan/log.
#pragma once
#include <stddef.h>

namespace curvis {

constexpr int kShapeTabK = 7;                 // 2^7 intervals per binade
constexpr int kShapeTabEmin = -10;            // first binade: [2^-10, 2^-9)
constexpr int kShapeTabEmax = 16;             // x < 2^16
constexpr int kShapeTabDegree = 5;
constexpr int kShapeTabDoubles = 2 * (kShapeTabDegree + 1);   // per interval: F a0..a5, then G b0..b5
constexpr size_t kShapeTabIntervals = (size_t)(kShapeTabEmax - kShapeTabEmin) << kShapeTabK;
constexpr unsigned kShapeTabShift = 20 - kShapeTabK;          // high-word bits below the interval index
constexpr unsigned kShapeTabBase = (unsigned)(1023 + kShapeTabEmin) << kShapeTabK;   // index of the first interval

// Fills out[kShapeTabIntervals * kShapeTabDoubles]; pure host arithmetic.
void build_interstellar_shape_table(double* out);

// The per-metric edition the default fast kernel (fast_variant 1) reads: what the regrouped step needs from the shape
// function is 1/r(l)^2 (theta and phi advance by it) and r'(l)/r(l)^3 (the radial force), so for a given (rho, m) the table
// holds exactly those two, as functions of z = |l| - a (the distance from the throat's plateau, ONE fp64 subtraction in the
// step; x = 2 z / (pi m) is folded into the coefficients):
//     U(z) = 1 / (rho + m F(x))^2      and      H(z) = G(x) / (rho + m F(x))^3,
// degree 5 each, on 2^kInvTabK = 128 intervals per binade of z in [2^kInvTabEmin, 2^kInvTabEmax).  The step then takes its one
// reciprocal of sin^2 theta alone (w = U / sin^2, r'/r^3 = sign(l) H): seven fp64 instructions fewer than going through r, and
// the metric parameters leave the loop.  One extra CONSTANT row (U = 1/rho^2, H = 0: the plateau |l| <= a of the throat,
// metrics.rs:470 / :482) receives every z below the range — zero, negative, denormal — through an unsigned min of the index, so the
// step has no branch and no call for the plateau (at z = 2^-44 the neglected m F is < 1e-20 m for every m >= 1e-3); z >=
// 2^kInvTabEmax is kept out of the loop by the step's radius gate (interstellar_table_l_limit).
// Accuracy and interval width: U behaves like z^-2 and H like z^-3 for large z, whose seventh Taylor coefficients are 7 and 28
// (1/z: 1), so on 2^-7-wide intervals the degree-5 interpolation error reaches 8 resp. 28 units of 2^-53 (relative; 14 / 48 in
// the transition zone x ~ 1 of a large-m metric) at the START of a binade, falls 64-fold towards its end, and is 0.8 resp. 2.9
// units r.m.s. — the size of the Horner evaluation's own rounding.  2^-8-wide intervals keep both below 2 units everywhere,
// but the hot part of the table (z in [2^-6, 2^7): 320 KB) then no longer fits the L1: hit rate 99.8 % -> 91 %, 4K frame
// 53.1 -> 55.9 ms.  What the guard band has to cover does not depend on the choice (tools/guard_study_interstellar.py, four
// scenes, rays with stiffness < 1: direction 9.6e-13 against 9.6e-13, l 6.0e-9 against 6.1e-9): the deviation from the
// operation-for-operation kernel comes from the regrouped roundings of 2000 steps, not from the table.  Checked against long
// double in tests/test_abi_host.py (host) and tests/test_gpu_fast64.py (device, bit-identical to the host evaluation).
// (Round-2 history: the first per-metric table held Y = 1/r and G = |r'| as functions of x; u = Y^2 and r'/r^3 = G Y^3 cost
// three more multiplications per step, and x = fma(|l|, xscale, xoff) needed its addend re-loaded into a vector register
// every step.)
constexpr int kInvTabK = 7;                   // 2^7 intervals per binade
constexpr int kInvTabEmin = -44;
constexpr int kInvTabEmax = 14;               // z < 16384
constexpr unsigned kInvTabShift = 20 - kInvTabK;
constexpr size_t kInvTabConstRow = (size_t)(kInvTabEmax - kInvTabEmin) << kInvTabK;   // index of the constant row
constexpr size_t kInvTabSelfRow = kInvTabConstRow + 1;   // one more row: its first 8 bytes hold the device address of the table itself (fast_f64.cuh)
constexpr size_t kInvTabIntervals = kInvTabConstRow + 2;
constexpr unsigned kInvTabBase = (unsigned)(1023 + kInvTabEmin) << kInvTabK;
void build_interstellar_inverse_table(double rho, double m, double* out);   // out[kInvTabIntervals * kShapeTabDoubles]
// |l| below which z = |l| - a stays inside the table (with a margin of one part in 2^20)
double interstellar_table_l_limit(double m, double a);

// atan and ln for the operation-for-operation kernel (CURVIS_PRECISION_F64, geodesic_f64.cuh: ShapeInterstellar::eval_fast).  The
// reference evaluates r(l) as rho + m (x atan x - ln(1 + x^2)/2), operation by operation (metrics.rs:467-468); so does that
// kernel, and two thirds of its Interstellar step were the CUDA library's atan and log (branches, ~80 fp64-pipe instructions, a
// 370 ns dependency chain).  These tables give the same two functions to <= 1.5 ulp of the exact values — the class of the CUDA
// library (2 ulp for atan, 1 for log); no libm is bit-identical to another anyway — in one table row and one degree-5 Horner
// chain each:
//   * atan(x) on x in [2^kShapeTabEmin, 2^kShapeTabEmax), the intervals of the F/G table above;
//   * ln(y) on y in [1, 2^kLogTabEmax), 2^kShapeTabK intervals per binade (y = 1 + x^2 < 2^33 for every x of the atan table).
//     Near y = 1, where ln -> 0, the polynomial's ABSOLUTE error (< 2^-55) is what holds: ln(1 + x^2) enters r as m ln/2 beside
//     rho, and the reference's own 1 + x*x has already rounded x^2 to 2^-53 there.
// Outside those ranges the kernel calls the library.
constexpr int kLogTabEmax = 33;
constexpr size_t kAtanTabIntervals = kShapeTabIntervals;
constexpr size_t kLogTabIntervals = (size_t)kLogTabEmax << kShapeTabK;
constexpr unsigned kLogTabBase = (unsigned)1023 << kShapeTabK;
constexpr int kFnTabDoubles = kShapeTabDegree + 1;            // per interval: a0..a5
void build_atan_table(double* out);   // out[kAtanTabIntervals * kFnTabDoubles]
void build_log_table(double* out);    // out[kLogTabIntervals * kFnTabDoubles]

// The fp32 edition for CURVIS_PRECISION_F32 (render_f32.cu): 2^4 intervals per binade of the same range,
// degree-3 polynomials, 8 floats per interval (F a0..a3, then G b0..b3): two 128-bit loads and six FFMA
// replace atanf + logf.  Relative error ~1e-7 (the fp32 rounding floor).
constexpr int kShapeTab32K = 4;
constexpr int kShapeTab32Degree = 3;
constexpr int kShapeTab32Floats = 2 * (kShapeTab32Degree + 1);
constexpr size_t kShapeTab32Intervals = (size_t)(kShapeTabEmax - kShapeTabEmin) << kShapeTab32K;
constexpr unsigned kShapeTab32Shift = 23 - kShapeTab32K;      // float bits below the interval index
constexpr unsigned kShapeTab32Base = (unsigned)(127 + kShapeTabEmin) << kShapeTab32K;
void build_interstellar_shape_table_f32(float* out);

}  // namespace curvis
