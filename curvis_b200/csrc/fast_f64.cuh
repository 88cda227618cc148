// fast_f64.cuh — arithmetic primitives of CURVIS_PRECISION_F64_FAST (render_f64_fast.cu): a
// <= 1 ulp reciprocal without the correction step, and sin^2 / sin*cos from one range reduction.
// Shared with the op-level test hook (curvis_debug_eval ops 10-12, render_f64.cu).
#pragma once
#include <cuda_runtime.h>
#include "ieee_f64.cuh"
#include "trig_f64.cuh"
#include "shape_table.h"

namespace curvis {

// 1/d to <= 1 ulp for d in the safe window: seed (~2^-20) + one cubic Newton step.
__device__ __forceinline__ double rcp_1ulp(double d) {
    const double y = rcp_seed(d);
    double e = fma(-d, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
}

// An fp64 instruction takes at most ONE uniform (constant-bank) operand, so the constants that
// meet another constant in the same FMA — 2/pi with the rounding magic, the leading coefficient
// of each polynomial — must sit in vector registers.  Left to itself ptxas re-loads them (LDC)
// every step; read through ld.global they stay in registers for the whole kernel.  The remaining
// constants come from the constant bank into uniform registers once per window.
static __device__ __constant__ double kReduce[2] = {1.5707963267948966, 6.123233995736766e-17};   // pi/2 hi, mid
static __device__ double kPinned[3] = {0.6366197723675814, 1.5912475864762696e-10, -1.1379094621237813e-11};   // 2/pi, kSinPoly[5], kCosPoly[5]

struct TrigRegs {
    double two_over_pi, sin5, cos5;
    __device__ __forceinline__ void load() {
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(two_over_pi) : "l"(kPinned));
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(sin5) : "l"(kPinned + 1));
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(cos5) : "l"(kPinned + 2));
    }
};

// sin^2(x) and sin(x)cos(x) for |x| < 2^30.  With x = k*pi/2 + r, |r| <= pi/4:
//   sin^2 x = (k odd) ? cos^2 r : sin^2 r,   sin x cos x = (-1)^k sin r cos r.
// Two-term Cody-Waite reduction (|k| is small: theta stays within a few multiples of pi).
__device__ __forceinline__ void sin2_sincos(const TrigRegs& tr, double x, double& s2, double& cs) {
    const double t = fma(x, tr.two_over_pi, kRoundMagic);
    const int k = __double2loint(t);
    const double q = t - kRoundMagic;
    double r = fma(-q, kReduce[0], x);
    r = fma(-q, kReduce[1], r);
    const double u = r * r;
    double sp = fma(u, tr.sin5, kSinPoly[4]);
    double cp = fma(u, tr.cos5, kCosPoly[4]);
    sp = fma(u, sp, kSinPoly[3]);
    cp = fma(u, cp, kCosPoly[3]);
    sp = fma(u, sp, kSinPoly[2]);
    cp = fma(u, cp, kCosPoly[2]);
    sp = fma(u, sp, kSinPoly[1]);
    cp = fma(u, cp, kCosPoly[1]);
    sp = fma(u, sp, kSinPoly[0]);
    cp = fma(u, cp, kCosPoly[0]);
    const double sr = fma(r * u, sp, r);                    // sin r
    const double cr = fma(u, fma(u, cp, -0.5), 1.0);        // cos r
    const double a = (k & 1) ? cr : sr;
    s2 = a * a;
    const double m = sr * cr;
    cs = __hiloint2double(__double2hiint(m) ^ (k << 31), __double2loint(m));
}

// ---- incremental trigonometry (fast_variant 1, the default of render_f64_fast.cu)
// theta changes by d = delta * p_theta / r^2 per step, and |d| < 2^-10 on 87 % of all steps of the
// 4K Ellis frame (< 2^-4 on 99.99 %).  Instead of a full range reduction + two degree-5 polynomials
// per step, (sin theta, cos theta) are carried along and rotated by the small angle, in the
// three-shear form of a rotation (every update in place, no temporaries):
//     c -= t s;   s += sd c;   c -= t s;        t = tan(d/2),  sd = sin d
// with sin d = d (1 + v S(v)) and tan(d/2) = d (1/2 + v T(v)), v = d*d; S and T are degree-2 Remez
// fits on |d| < 2^-4 (tools/gen_rot_coeffs.py: error 2.0e-17 resp. 3.1e-16 relative to d, i.e.
// <= 2e-17 absolute).  12 fp64 instructions instead of 21.  The pair is re-derived from theta itself
// (sincos_fast) at the start of every window of steps and after any step with |d| >= 2^-4, so its
// rounding drift is bounded by one window (<= 128 steps * ~1.5e-16) instead of growing along the ray.
// (16-byte aligned: each pair is ONE LDCU.128 in the step loop, wherever the linker puts the other constants)
static __device__ __constant__ __align__(16) double kRotSin[2] = {-0.16666666666666152, 0.008333333309682096};
static __device__ __constant__ __align__(16) double kRotTan[2] = {0.04166666666674629, 0.00416666629980884};
static __device__ double kRotPinned[2] = {-0.00019839655223880117, 0.0004218773803051587};   // S2, T2 (see kPinned)

struct RotRegs {
    double sin2, tan2;                    // leading coefficients: vector registers (see kPinned)
    __device__ __forceinline__ void load() {
        // The address carries a per-thread zero the compiler cannot see through (the lane id times the top bit of the cycle counter), so the
        // loads are not promoted to the uniform datapath: a coefficient living in a uniform register would force its
        // partner constant into a vector register by two moves EVERY step (a DFMA takes one uniform operand).
        unsigned lane_id;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_id));
        const unsigned long long zero = (unsigned long long)lane_id * ((unsigned long long)clock64() >> 63);
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(sin2) : "l"(kRotPinned + zero));
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(tan2) : "l"(kRotPinned + 1 + zero));
    }
};

// (s, c) = (sin, cos)(theta)  ->  (sin, cos)(theta + d), |d| < 2^-4.
// v_out = d*d (the caller's |d| < 2^-4 test reads its high word: v >= 0 needs no sign mask)
__device__ __forceinline__ void rotate_sincos(const RotRegs& rr, double d, double& s, double& c, double& v_out) {
    const double v = d * d;
    v_out = v;
    double sp = fma(v, rr.sin2, kRotSin[1]);
    double tp = fma(v, rr.tan2, kRotTan[1]);
    sp = fma(v, sp, kRotSin[0]);
    tp = fma(v, tp, kRotTan[0]);
    const double sd = d * fma(v, sp, 1.0);    // sin d
    const double t = d * fma(v, tp, 0.5);     // tan(d/2)
    c = fma(-t, s, c);
    s = fma(sd, c, s);
    c = fma(-t, s, c);
}
__device__ __forceinline__ void rotate_sincos(const RotRegs& rr, double d, double& s, double& c) {
    double v;
    rotate_sincos(rr, d, s, c, v);
}

// ---- Interstellar shape function from the per-metric table (shape_table.h: build_interstellar_inverse_table):
// U = 1/r(l)^2 and H = |r'(l)|/r(l)^3 at z = |l| - a, two degree-5 Horner chains on twelve coefficients.  Every z below the table
// (the plateau |l| <= a, z <= 0 included) reads the constant row through an unsigned min — no branch, no call; z beyond the
// table must be kept out by the caller (FrameParams::fast_l_limit).
//
// The coefficients are CACHED in the lane's registers and fetched again only when the step has left the interval.  Fetching
// them every step made the L1 data pipe the bound of the Interstellar kernel, not the fp64 pipe: 96 bytes per lane and step
// are 24 wavefronts of register write-back per warp-step whatever the addresses (profiles/r02_ncu_f64_fast_interstellar_
// v1_summary.txt: l1tex data-pipe 87 % of peak, fp64 pipe 65 %).  A photon advances |l| by ~0.05 per step and the intervals are
// z/128 wide, so beyond z ~ 6 it stays several steps in one interval (a third of all lane-steps fetch) — and the lanes of a
// warp, claimed from neighbouring pixels, change intervals on nearly the same steps, so whole quarter-warps skip the fetch.
struct InverseShapeCache {
    double a0, a1, a2, a3, a4, a5, b0, b1, b2, b3, b4, b5;
    const double2* tab;
    unsigned idx;
    // The table's address is pinned in a register pair: it is read from the table's own last row (where the host wrote it,
    // shape_table.h: kInvTabSelfRow) with a plain global load, so the compiler cannot know that it equals the kernel parameter.
    // Left to itself ptxas re-loads the parameter from the constant bank, under the fetch's predicate, in the middle of the
    // index -> address -> load chain of every step (an empty asm or an opaque zero offset do not stop it).
    __device__ __forceinline__ void reset(const double2* table) {
        unsigned long long self;
        asm volatile("ld.global.u64 %0, [%1];" : "=l"(self) : "l"(table + kInvTabSelfRow * (kShapeTabDoubles / 2)));
        tab = (const double2*)self;
        idx = 0xffffffffu;                                                 // (no interval has this index)
    }
};

__device__ __forceinline__ void interstellar_inverse_lookup(double a, double l, InverseShapeCache& k, double& U, double& H) {
    const double x = fabs(l) - a;        // z
    const unsigned hi = (unsigned)__double2hiint(x);
    const unsigned idx = min((hi >> kInvTabShift) - kInvTabBase, (unsigned)kInvTabConstRow);
    const double c = __hiloint2double((int)(hi & ~((1u << kInvTabShift) - 1u)), 0);   // the interval's lower edge: z's low mantissa bits cleared
    const double t = x - c;
    if (idx != k.idx) {
        // three 256-bit loads (sm_100: LDG.E.256; a row is 96 bytes, the table 256-byte aligned)
        const double2* e = k.tab + idx * (unsigned)(kShapeTabDoubles / 2);   // (32-bit offset: the table is < 2 MB)
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k.a0), "=d"(k.a1), "=d"(k.a2), "=d"(k.a3) : "l"(e));
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k.a4), "=d"(k.a5), "=d"(k.b0), "=d"(k.b1) : "l"(e + 2));
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k.b2), "=d"(k.b3), "=d"(k.b4), "=d"(k.b5) : "l"(e + 4));
        k.idx = idx;
    }
    U = fma(t, fma(t, fma(t, fma(t, fma(t, k.a5, k.a4), k.a3), k.a2), k.a1), k.a0);
    H = fma(t, fma(t, fma(t, fma(t, fma(t, k.b5, k.b4), k.b3), k.b2), k.b1), k.b0);
}

// The same lookup without a cache (one-off evaluations: test hook).
__device__ __forceinline__ void interstellar_inverse_lookup(const double2* __restrict__ tab, double a, double l, double& U, double& H) {
    InverseShapeCache k;
    k.reset(tab);
    interstellar_inverse_lookup(a, l, k, U, H);
}

// d >= 0 by construction (a product of squares and a positive radius), so the high word is
// its own magnitude key: finite, normal and in [2^-300, 2^300) <=> one unsigned compare.
// NaN (either sign) and Inf fall outside.
__device__ __forceinline__ bool in_window_nonneg(double d) {
    return ((unsigned)__double2hiint(d) - pow2_hi(-300)) < (pow2_hi(300) - pow2_hi(-300));
}

}  // namespace curvis
