// trig_f64.cuh — sin/cos for the Euler step, fp64, ~1 ulp (same class as the CUDA math
// library's sincos: <= 2 ulp).  Written out here because the library version costs ~35
// non-fp64 instructions per call in constant materialisation and guards (UMOV/IMAD.MOV/BRA,
// see profiles/r01_f64_v0_ncu_summary.txt); here the coefficients come straight from the
// constant bank as FMA operands and the large-argument guard is part of the step's single
// merged check.  Coefficients: tools/gen_trig_coeffs.py (Remez on [0,(1.02*pi/4)^2]).
//
// The reference calls the platform libm (glibc) here; no GPU implementation is bit-identical
// to it (glibc's own result depends on the CPU's FMA support), so the parity bar for the
// photon STATE is a tolerance while RGB / steps / texels must be identical.
#pragma once
#include <cuda_runtime.h>

namespace curvis {

static __device__ __constant__ double kSinPoly[6] = {
    -0.16666666666666663, 0.00833333333333043, -0.00019841269835988552,
    2.7557315710929657e-06, -2.5051051817332214e-08, 1.5912475864762696e-10};
static __device__ __constant__ double kCosPoly[6] = {
    0.041666666666666664, -0.0013888888888887072, 2.4801587298283765e-05,
    -2.755731702664308e-07, 2.0876096187138334e-09, -1.1379094621237813e-11};

constexpr double kTwoOverPi = 0.6366197723675814;
constexpr double kPio2Hi = 1.5707963267948966;
constexpr double kPio2Mid = 6.123233995736766e-17;
constexpr double kPio2Lo = -1.4973849048591698e-33;
constexpr double kRoundMagic = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to an integer in the low word
constexpr double kTrigFastLimit = 1073741824.0;      // 2^30: beyond it fall back to ::sincos (Payne-Hanek)

// sin and cos of x for |x| < kTrigFastLimit (caller guarantees it; NaN/Inf excluded).  CB = true reads the reduction's
// constants from the constant bank as FMA operands instead of immediates (two UMOV per use): fewer issue slots in the
// operation-for-operation kernel, where sincos runs every step; the regrouped kernel, which only re-derives (sin, cos) once
// per window, keeps the immediates (its step loop is tuned to the uniform registers it has).  Same arithmetic either way.
static __device__ __constant__ double kTrigReduce[5] = {0.6366197723675814, 6755399441055744.0, 1.5707963267948966,
                                                        6.123233995736766e-17, -1.4973849048591698e-33};

template <bool CB = false>
__device__ __forceinline__ void sincos_fast(double x, double& s, double& c) {
    const double magic = CB ? kTrigReduce[1] : kRoundMagic;
    const double t = fma(x, CB ? kTrigReduce[0] : kTwoOverPi, magic);
    const int k = __double2loint(t);                  // nearest integer to x*2/pi
    const double q = t - magic;
    double r = fma(-q, CB ? kTrigReduce[2] : kPio2Hi, x);
    r = fma(-q, CB ? kTrigReduce[3] : kPio2Mid, r);
    r = fma(-q, CB ? kTrigReduce[4] : kPio2Lo, r);
    const double u = r * r;
    double sp = fma(u, kSinPoly[5], kSinPoly[4]);
    double cp = fma(u, kCosPoly[5], kCosPoly[4]);
    sp = fma(u, sp, kSinPoly[3]);
    cp = fma(u, cp, kCosPoly[3]);
    sp = fma(u, sp, kSinPoly[2]);
    cp = fma(u, cp, kCosPoly[2]);
    sp = fma(u, sp, kSinPoly[1]);
    cp = fma(u, cp, kCosPoly[1]);
    sp = fma(u, sp, kSinPoly[0]);
    cp = fma(u, cp, kCosPoly[0]);
    const double sr = fma(r * u, sp, r);              // sin(r)
    const double cr = fma(u * u, cp, fma(u, -0.5, 1.0));  // cos(r)
    // quadrant: sin(x) = [sr, cr, -sr, -cr][k&3], cos(x) = [cr, -sr, -cr, sr][k&3]
    const double a = (k & 1) ? cr : sr;
    const double b = (k & 1) ? sr : cr;
    // sign flips on the integer pipe (bit 1 of k, resp. k+1, moved onto the sign bit)
    s = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
    c = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
}

// The three constants of sincos_fast that meet another constant in the same FMA (an fp64 instruction takes ONE uniform /
// constant-bank operand, so the other must sit in a vector register): the rounding magic and the two leading coefficients.
// Left to itself ptxas re-loads them (LDC) every step — six issue slots of the operation-for-operation kernel's step; read once
// per kernel through a plain global load they stay in registers (the same trick as fast_f64.cuh: kPinned).
// `half` = 0.5: cos r's fma(u, -0.5, 1.0) and the square root's fma(e, 0.375, 0.5) each hold two constants; pinned (an option:
// two more registers), the second one costs no instruction (ptxas otherwise builds it with two moves per use).
static __device__ double kTrigPinned[4] = {6755399441055744.0, 1.5912475864762696e-10, -1.1379094621237813e-11, 0.5};   // magic, kSinPoly[5], kCosPoly[5], 1/2

struct TrigPins {
    double magic, sin5, cos5, half;
    __device__ __forceinline__ void load(bool pin_half = false) {
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(magic) : "l"(kTrigPinned));
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(sin5) : "l"(kTrigPinned + 1));
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(cos5) : "l"(kTrigPinned + 2));
        half = 0.5;                                                  // (a literal to the compiler unless pinned)
        if (pin_half) asm volatile("ld.global.f64 %0, [%1];" : "=d"(half) : "l"(kTrigPinned + 3));
    }
};

// sincos_fast<true> with those three in registers: the same operations on the same values.
__device__ __forceinline__ void sincos_fast_pinned(const TrigPins& tp, double x, double& s, double& c) {
    const double t = fma(x, kTrigReduce[0], tp.magic);
    const int k = __double2loint(t);
    const double q = t - tp.magic;
    double r = fma(-q, kTrigReduce[2], x);
    r = fma(-q, kTrigReduce[3], r);
    r = fma(-q, kTrigReduce[4], r);
    const double u = r * r;
    double sp = fma(u, tp.sin5, kSinPoly[4]);
    double cp = fma(u, tp.cos5, kCosPoly[4]);
    sp = fma(u, sp, kSinPoly[3]);
    cp = fma(u, cp, kCosPoly[3]);
    sp = fma(u, sp, kSinPoly[2]);
    cp = fma(u, cp, kCosPoly[2]);
    sp = fma(u, sp, kSinPoly[1]);
    cp = fma(u, cp, kCosPoly[1]);
    sp = fma(u, sp, kSinPoly[0]);
    cp = fma(u, cp, kCosPoly[0]);
    const double sr = fma(r * u, sp, r);
    const double cr = fma(u * u, cp, fma(u, -tp.half, 1.0));
    const double a = (k & 1) ? cr : sr;
    const double b = (k & 1) ? sr : cr;
    s = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
    c = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
}

struct TrigFast {
    static __device__ __forceinline__ void sincos(double x, double& s, double& c) {
        if (fabs(x) < kTrigFastLimit) sincos_fast(x, s, c);
        else ::sincos(x, &s, &c);
    }
    static __device__ __forceinline__ double sin(double x) {
        double s, c;
        sincos(x, s, c);
        return s;
    }
};

}  // namespace curvis
