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
