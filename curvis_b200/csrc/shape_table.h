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
// function is 1/r(l) and r'(l), so for a given (rho, m) the table holds, on the same intervals,
//     Y(x) = 1 / (rho + m F(x))      and      G(x),
// degree 5 each.  The step then takes its one reciprocal of sin^2 theta alone (u = Y^2, w = u / sin^2, r'/r^3 = G Y^3):
// four fp64 instructions fewer than going through r, and the metric parameters leave the loop.  The range reaches down
// to 2^kInvTabEmin, and one extra CONSTANT row (Y = 1/rho, G = 0: the plateau |l| <= a of the throat, metrics.rs:470 / :482)
// receives every x below it — zero, negative, denormal — through an unsigned min of the index, so the step has no
// branch and no call for the plateau; x >= 2^kShapeTabEmax is kept out of the loop by the step's radius gate.
// 1/r behaves like 1/x for large x: the degree-5 interpolation error is ~2^-53 (the t^6 coefficient of 1/(1+t) is 1),
// checked against long double in tests/test_abi_host.py (host) and tests/test_gpu_fast64.py (device): <= 2.5 ulp.
constexpr int kInvTabEmin = -40;
constexpr size_t kInvTabConstRow = (size_t)(kShapeTabEmax - kInvTabEmin) << kShapeTabK;   // index of the constant row
constexpr size_t kInvTabIntervals = kInvTabConstRow + 1;
constexpr unsigned kInvTabBase = (unsigned)(1023 + kInvTabEmin) << kShapeTabK;
void build_interstellar_inverse_table(double rho, double m, double* out);   // out[kInvTabIntervals * kShapeTabDoubles]
// |l| below which x = (|l| - a) 2/(pi m) stays inside the table (with a margin of one part in 2^20)
double interstellar_table_l_limit(double m, double a);

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
