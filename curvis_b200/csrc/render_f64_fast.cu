// render_f64_fast.cu — CURVIS_PRECISION_F64_FAST: the same forward-Euler scheme, in fp64, with
// the right-hand side of update_relativistic_object (reference src/metrics.rs:223-270)
// regrouped for the B200's fp64 pipe.
//
// The parity kernel (render_f64.cu) keeps one rounding per reference operation: six correctly
// rounded divisions, one correctly rounded square root and sincos cost ~100 fp64-pipe
// instructions per step, and that pipe issues one warp instruction per 2-3 cycles per scheduler
// (profiles/r01_microbench_fp64_pipe.txt) — it is the bound.  Here the same quantities are
// computed with ~45:
//   * ONE reciprocal per step, w = 1/(r^2 sin^2 theta) (MUFU.RCP64H seed + one cubic Newton step,
//     <= 1 ulp); the others follow by multiplication: 1/r^2 = w sin^2, 1/sin^2 = w r^2,
//     cos/(r^2 sin^3) = cos sin * w * (1/sin^2);
//   * Ellis: r'/r^3 = l/r^4 — no square root, no division;
//   * the trigonometry delivers sin^2 theta and sin theta cos theta directly (the only
//     combinations the right-hand side reads), which removes the quadrant selects of cos;
//   * delta is folded into per-ray constants and the five state updates are single FMAs.
// Every operation is accurate to <= 1 ulp, but the rounding points differ from the reference's,
// so the final photon state agrees with the oracle to ~1e-13 relative instead of ~1e-15; escape
// side, step count, texel and RGB8 agree wherever a 1e-13 perturbation does not cross a decision
// boundary (tests/test_gpu_fast64.py states the bar; bench.py reports the measured deviation).
// Operands outside the window in which the unguarded sequences are exact (rays grazing the
// coordinate poles, NaN/Inf states, huge angles) take the parity kernel's step instead, so the
// exotic cases (Flat metric NaN rays, NotEscaped) behave exactly as in parity mode.
//
// Execution model: identical to render_rows_f64_lean (persistent grid, one ray per lane,
// windowed ballot refill from one work queue).  Compiled with -fmad=false like every TU that
// includes geodesic_f64.cuh (ray generation and the escaped-photon epilogue are the parity code);
// every fused operation below is an explicit fma().
#include "geodesic_f64.cuh"
#include "fast_f64.cuh"
#include "launch.h"
#include "shape_table.h"

namespace curvis {

namespace {

constexpr int kBlockFast = 128;
constexpr double kLongRaySin = 0.03;   // refill: near-critical rays whose orbit comes within asin(0.03) of the polar axis are claimed first
constexpr unsigned kFullFast = 0xffffffffu;

// Shape policies of the fast step.  factors() returns, for the current l and sin^2 theta:
//   w = 1/(r^2 sin^2), u = 1/r^2, v = 1/sin^2, fd = delta * r'(l)/r(l)^3; false when an operand
// left the safe window (the caller then takes parity steps).
struct FastEllis {   // metrics.rs:417-421 : r^2 = rho^2 + l^2, r' = l/r  =>  r'/r^3 = l/r^4
    using Shape64 = ShapeEllis;
    static __device__ __forceinline__ bool factors(const FrameParams& p, double l, double s2, double& w, double& u, double& v, double& ud, double& fd) {
        const double r2 = fma(l, l, p.d_rho2);
        const double d = r2 * s2;
        if (!in_window_nonneg(d)) return false;
        w = rcp_1ulp(d);
        u = w * s2;
        v = w * r2;
        ud = u * p.delta;
        fd = l * (u * ud);
        return true;
    }
    // The same quantities split around the reciprocal (fast_variant 1): prepare() returns the
    // divisor d = r^2 sin^2, finish() turns y0 = 1/d into w, u, v and f = r'/r^3 (no delta: the
    // momenta are pre-scaled, see fast_window_scaled).
    struct Pre { double r2; };
    struct Cache { __device__ __forceinline__ void reset(const double2*) {} };
    static __device__ __forceinline__ double prepare(const FrameParams& p, double l, double s2, Pre& pre, Cache&) {
        pre.r2 = fma(l, l, p.d_rho2);
        return pre.r2 * s2;
    }
    static __device__ __forceinline__ void finish(const FrameParams&, const Pre& pre, Cache&, double y0, double l, double s2, double& w, double& u, double& v, double& f) {
        w = y0;
        u = w * s2;
        v = w * pre.r2;
        f = l * (u * u);
    }
    static __device__ __forceinline__ bool beyond(const FrameParams&, double) { return false; }
    // Radius gate of the step loop on r^2 = l^2 + rho^2, which prepare() has just formed for the next step: positive, so its
    // high word compares without the sign mask |l| needs.  fma(l, l, rho^2) is monotone in |l|: |l| > R_gate implies
    // r2 >= fma(R_gate, R_gate, rho^2), so no escape is missed; the few false alarms (same high word, smaller value) fall
    // through the caller's exact tests.  NaN l gives NaN r2 (high word above every threshold).
    static __device__ __forceinline__ unsigned gate_key(const FrameParams& p, double R_gate) {
        return (R_gate >= 0.0) ? (unsigned)__double2hiint(fma(R_gate, R_gate, p.d_rho2)) : 0u;
    }
    static __device__ __forceinline__ bool at_gate(const Pre& pre, unsigned key) { return (unsigned)__double2hiint(pre.r2) >= key; }
    static constexpr bool kGateFromSquares = true;
    // longest-first refill in launches of any size ("longest_first" = 2): with the listed rays in the favoured warp slots a whole 4K
    // frame gains 1.1 % (35.29 -> 34.89 ms); the Interstellar kernel loses 0.7 % there (its listed rays run the slow pole-crossing
    // code all at once) and keeps the rule "at most 64 rays per lane"
    static constexpr bool kLongFirstWholeFrames = true;
};

// r(l) > 0 and r'(l) given: one reciprocal of r*sin^2 yields 1/r and 1/sin^2.
__device__ __forceinline__ bool factors_from_r(const FrameParams& p, double r, double rp, double s2, double& w, double& u, double& v, double& ud, double& fd) {
    const double d = r * s2;
    if (!in_window_nonneg(d)) return false;   // also catches NaN l, r <= 0
    const double y0 = rcp_1ulp(d);
    const double y = y0 * s2;      // 1/r
    v = y0 * r;                    // 1/sin^2
    u = y * y;
    w = u * v;
    ud = u * p.delta;
    fd = rp * (y * ud);
    return true;
}

struct PreFromR { double r, rp; };
__device__ __forceinline__ void finish_from_r(const PreFromR& pre, double y0, double s2, double& w, double& u, double& v, double& f) {
    const double y = y0 * s2;      // 1/r
    v = y0 * pre.r;                // 1/sin^2
    u = y * y;
    w = u * v;
    f = pre.rp * (y * u);          // r'/r^3
}

// F(x) = x atan x - ln(1+x^2)/2 and G(x) = (2/pi) atan x for x > 0 from the piecewise degree-5 table
// (shape_table.h): interval index = a shift of x's high word, t = x - midpoint (exact), two Horner
// chains on coefficients fetched with six 128-bit loads (neighbouring rays sit in the same or the
// next interval, so the loads are L1 hits).  Outside [2^-10, 2^16): the library functions; x <= 0 (the
// plateau |l| <= a of the throat, metrics.rs:470 / :482) and NaN give F = G = 0, i.e. r = rho, r' = 0.
__device__ __noinline__ double2 shape_fg_library(double x) {   // x outside the table: rare, out of line
    if (!(x > 0.0)) return make_double2(0.0, 0.0);
    const double at = atan(x);
    return make_double2(fma(x, at, -0.5 * log(fma(x, x, 1.0))), (2.0 / CURVIS_PI) * at);
}

__device__ __forceinline__ void shape_fg(const FrameParams& p, double x, double& F, double& G) {
    const unsigned hi = (unsigned)__double2hiint(x);
    const unsigned idx = (hi >> kShapeTabShift) - kShapeTabBase;
    if (idx < (unsigned)kShapeTabIntervals) {
        const double c = __hiloint2double((int)((hi & ~((1u << kShapeTabShift) - 1u)) | (1u << (kShapeTabShift - 1))), 0);
        const double t = x - c;
        const double2* e = p.shape_tab + (size_t)idx * (kShapeTabDoubles / 2);
        const double2 a01 = __ldg(e), a23 = __ldg(e + 1), a45 = __ldg(e + 2);
        const double2 b01 = __ldg(e + 3), b23 = __ldg(e + 4), b45 = __ldg(e + 5);
        F = fma(t, fma(t, fma(t, fma(t, fma(t, a45.y, a45.x), a23.y), a23.x), a01.y), a01.x);
        G = fma(t, fma(t, fma(t, fma(t, fma(t, b45.y, b45.x), b23.y), b23.x), b01.y), b01.x);
    } else {
        const double2 fg = shape_fg_library(x);
        F = fg.x; G = fg.y;
    }
}

struct FastInterstellar {   // metrics.rs:461-485 with the uniform divisor pi*m folded into d_xscale
    using Shape64 = ShapeInterstellar;
    // fast_variant 0: r and r' through the parameter-free table of F and G (library functions outside its range).
    // One path for the whole l axis: x <= 0 on the plateau |l| <= a fails the table's range test and comes back as
    // F = G = 0 from the library branch (a ray spends at most a step or two there), so the step carries no branch on l.
    static __device__ __forceinline__ void shape(const FrameParams& p, double l, double& r, double& rp) {
        const double x = (fabs(l) - p.a) * p.d_xscale;
        double F, G;
        shape_fg(p, x, F, G);
        r = fma(p.m, F, p.rho);
        rp = copysign(G, l);
    }
    static __device__ __forceinline__ bool factors(const FrameParams& p, double l, double s2, double& w, double& u, double& v, double& ud, double& fd) {
        double r, rp;
        shape(p, l, r, rp);
        return factors_from_r(p, r, rp, s2, w, u, v, ud, fd);
    }
    // fast_variant 1 (default): U = 1/r^2 and H = |r'|/r^3 from the per-metric table (shape_table.h) — two degree-5 Horner chains on
    // coefficients cached in registers while the photon stays in one interval (fast_f64.cuh); the step's one reciprocal is then
    // 1/sin^2 theta alone.  No call: every x below the table (the plateau, x <= 0 included) reads the constant row through an
    // unsigned min; x beyond it never gets here (beyond(): the kernel's radius gate).  41 fp64-pipe instructions per step.
    // The lookup sits in finish(), i.e. at the head of the step it serves: nothing but the photon is carried from one trip of the
    // loop to the next (looked up a step ahead, in prepare(), U and H changed registers by four moves per step).
    struct Pre {};
    using Cache = InverseShapeCache;
    static __device__ __forceinline__ double prepare(const FrameParams&, double, double s2, Pre&, Cache&) { return s2; }
    static __device__ __forceinline__ void finish(const FrameParams& p, const Pre&, Cache& cache, double y0, double l, double, double& w, double& u, double& v, double& f) {
        double H;
        interstellar_inverse_lookup(p.a, l, cache, u, H);   // u = 1/r^2
        v = y0;                        // 1/sin^2
        w = u * v;
        f = copysign(H, l);            // r'/r^3
    }
    static __device__ __forceinline__ bool beyond(const FrameParams& p, double l) { return !(fabs(l) < p.fast_l_limit); }
    static __device__ __forceinline__ unsigned gate_key(const FrameParams&, double) { return 0u; }
    static __device__ __forceinline__ bool at_gate(const Pre&, unsigned) { return false; }
    static constexpr bool kGateFromSquares = false;
    static constexpr bool kLongFirstWholeFrames = false;
};

struct FastFlat {   // metrics.rs:501-505: r = l, r' = 1 (r may be negative: take the parity step then)
    using Shape64 = ShapeFlat;
    static __device__ __forceinline__ bool factors(const FrameParams& p, double l, double s2, double& w, double& u, double& v, double& ud, double& fd) {
        return factors_from_r(p, l, 1.0, s2, w, u, v, ud, fd);
    }
    using Pre = PreFromR;
    struct Cache { __device__ __forceinline__ void reset(const double2*) {} };
    static __device__ __forceinline__ double prepare(const FrameParams&, double l, double s2, Pre& pre, Cache&) {
        pre.r = l; pre.rp = 1.0;
        return l * s2;
    }
    static __device__ __forceinline__ void finish(const FrameParams&, const Pre& pre, Cache&, double y0, double, double s2, double& w, double& u, double& v, double& f) {
        finish_from_r(pre, y0, s2, w, u, v, f);
    }
    static __device__ __forceinline__ bool beyond(const FrameParams&, double) { return false; }
    static __device__ __forceinline__ unsigned gate_key(const FrameParams&, double) { return 0u; }
    static __device__ __forceinline__ bool at_gate(const Pre&, unsigned) { return false; }
    static constexpr bool kGateFromSquares = false;
    static constexpr bool kLongFirstWholeFrames = false;
};

// One forward-Euler step (metrics.rs:283-297) with the regrouped right-hand side.  Returns
// false, leaving the state untouched, when an operand is outside the safe window.
template <class Fast>
__device__ __forceinline__ bool fast_step(const FrameParams& p, const TrigRegs& tr, Ray& q) {
    double s2, cs, w, u, v, ud, fd;
    sin2_sincos(tr, q.th, s2, cs);
    if (!Fast::factors(p, q.l, s2, w, u, v, ud, fd)) return false;
    const double pv = q.pph2 * v;                           // p_phi^2 / sin^2
    const double b2 = fma(q.pth, q.pth, pv);                // metrics.rs:257
    const double wd = w * p.delta;
    q.l = fma(q.pl, p.delta, q.l);                          // :238, :295
    q.th = fma(q.pth, ud, q.th);                            // :239
    q.ph = fma(q.pph, wd, q.ph);                            // :240
    q.pl = fma(b2, fd, q.pl);                               // :261, :296
    q.pth = fma(pv * cs, wd, q.pth);                        // :262  p_phi^2 cos / (r^2 sin^3)
    return true;
}

// fast_variant 1 (default): up to n steps of one window with
//   * (sin theta, cos theta) carried along and rotated by the step's dtheta (fast_f64.cuh), re-derived
//     from theta at the start of the window and after a step with |dtheta| >= 2^-4;
//   * the momenta pre-scaled by delta, P = delta * p (the caller keeps q.pl, q.pth, q.pph, q.pph2 in
//     that form).  The geodesic equations are homogeneous in the momenta — this is the same Euler
//     iteration with the affine parameter rescaled to unit steps — and delta disappears from the loop:
//         l += P_l;  theta += P_theta u;  phi += P_phi w;
//         P_l += (P_theta^2 + P_phi^2 v) r'/r^3;  P_theta += P_phi^2 v (sin cos) w
//     (u = 1/r^2, v = 1/sin^2, w = u v): 33 fp64 instructions per Ellis step instead of 41;
//   * ONE exit branch per step: the four rarely-true conditions (step budget used up, |l| within reach
//     of the escape radius or NaN, |dtheta| too large for the rotation, next divisor outside the safe
//     window) are OR-ed on the integer pipe and sorted out after the loop.
// Returns the number of steps taken; `near` = the last step took |l| to the radius gate (the high word of R: |l| within
// 2^-20 R of the radius or beyond it; for Interstellar also the end of the shape table) or made it NaN: the caller runs the
// escape test of systems.rs:129-134 and records how close the step came to +-R on either side — `last_b2`, `last_f` (the
// factors of that step's p_l increment) let it reconstruct the l the step started from; `slow` = the remaining steps of
// the window need the parity step; `wmax_hi` = running maximum of the high word of w = 1/(r^2 sin^2 theta): (P_phi w)^2 is
// the stiffness of curvis_ray_record.
template <class Fast>
__device__ __forceinline__ uint32_t fast_window_scaled(const FrameParams& p, const RotRegs& rr, Ray& q, uint32_t n, unsigned gate, unsigned gate_key,
                                                       bool& near, bool& slow, unsigned& wmax_hi, double& wsum_out, double& last_b2, double& last_f) {
    uint32_t left = n;   // steps still allowed (a down-counter: one instruction per step)
    // phi += P_phi * w every step (:240): the w's are summed (a two-operand DADD issues faster than a DFMA with three
    // distinct registers) and folded into phi once, by the caller, when the window is left
    double wsum = 0.0;
    typename Fast::Cache cache;   // Interstellar: the shape table's coefficients of the interval the photon is in
    cache.reset(p.inv_tab);
    for (;;) {
        if (abs_hi(q.th) >= pow2_hi(30)) { slow = true; break; }
        double sn, cn;
        sincos_fast(q.th, sn, cn);
        typename Fast::Pre pre;
        double s2 = sn * sn;
        double d = Fast::prepare(p, q.l, s2, pre, cache);
        if (!in_window_nonneg(d)) { slow = true; break; }
        double dth;
        for (;;) {
            double w, u, v, f;
            Fast::finish(p, pre, cache, rcp_1ulp(d), q.l, s2, w, u, v, f);
            const double cs = sn * cn;
            dth = q.pth * u;                                    // metrics.rs:239
            const double pv = q.pph2 * v;                       // p_phi^2 / sin^2
            const double b2 = fma(q.pth, q.pth, pv);            // :257
            q.l = q.l + q.pl;                                   // :238, :295
            q.th = q.th + dth;
            wsum = wsum + w;                                    // :240, folded below
            wmax_hi = max(wmax_hi, (unsigned)__double2hiint(w));   // stiffness monitor (w > 0: the high word orders it)
            q.pl = fma(b2, f, q.pl);                            // :261, :296
            q.pth = fma(pv * cs, w, q.pth);                     // :262
            last_b2 = b2; last_f = f;                           // read after the loop only (the step that reached the radius)
            double dth2 = 0.0;
            if (Fast::kGateFromSquares) rotate_sincos(rr, dth, sn, cn, dth2);
            else rotate_sincos(rr, dth, sn, cn);
            --left;
            asm("" : "+r"(left));   // one induction variable (the optimiser otherwise keeps two copies of the counter)
            s2 = sn * sn;
            d = Fast::prepare(p, q.l, s2, pre, cache);
            if (Fast::kGateFromSquares) {
                // Ellis: |dtheta| >= 2^-4 read off dtheta^2 >= 2^-8 and |l| at the gate off r^2 = l^2 + rho^2 (FastEllis::at_gate) —
                // positive numbers, whose high words compare without a sign mask: 44 instructions per step instead of 46
                if ((left == 0u) | Fast::at_gate(pre, gate_key) | ((unsigned)__double2hiint(dth2) >= pow2_hi(-8)) | !in_window_nonneg(d)) break;
            } else {
                // (the Interstellar loop gains nothing from those forms: ptxas' schedule of it is 2.5 % slower with them)
                if ((left == 0u) | (abs_hi(q.l) >= gate) | (abs_hi(dth) >= pow2_hi(-4)) | !in_window_nonneg(d)) break;
            }
        }
        if (abs_hi(q.l) >= gate) { near = true; break; }        // |l| >= R (1 - 2^-20), or past the shape table, or NaN: the caller's business
        if (left == 0u) break;
        if (!(abs_hi(dth) >= pow2_hi(-4)) && !in_window_nonneg(d)) { slow = true; break; }
        // |dtheta| too large for the rotation: re-derive (sin, cos) and go on
    }
    wsum_out = wsum;   // the caller folds it: phi += P_phi * wsum (phi and P_phi live in shared memory, not in registers)
    return n - left;
}

// Parity steps for a lane whose operands left the safe window: plain operators, reference
// arithmetic (euler_step_lean with the guards off).  Out of line so that the hot loop keeps its
// registers and uniform constants to itself; rays that come here graze a coordinate pole or
// carry NaN/Inf.
struct SlowResult { double l, th, ph, pl, pth; uint32_t steps; bool stop; };

template <class Shape64>
__device__ __noinline__ SlowResult parity_steps(const FrameParams& p, Ray q, uint32_t n, unsigned gate) {
    const double R = p.max_radius;
    uint32_t k = 0;
    bool stop = false;
    while (k < n) {
        euler_step_lean<Shape64>(p, q, false);
        ++k;
        if (abs_hi(q.l) >= gate) {
            if ((q.l > R) || (q.l < -R) || (q.l != q.l)) { stop = true; break; }   // systems.rs:129-134
        }
    }
    SlowResult r;
    r.l = q.l; r.th = q.th; r.ph = q.ph; r.pl = q.pl; r.pth = q.pth; r.steps = k; r.stop = stop;
    return r;
}

// The photon of fast_variant 1 (momenta scaled by delta) in the reference's units, for the parity
// steps and the epilogue.  p_phi is conserved, so it is regenerated from the ray index rather than
// divided back (bit-exact, and it spares the hot loop two registers); a ray that has not moved is
// regenerated whole.
__device__ __noinline__ Ray unscaled_ray(const FrameParams& p, const Ray& q, unsigned long long ray, unsigned long long tile_rays, bool untouched) {
    Ray o;
    new_photon_for_ray(p, ray, tile_rays, o);
    if (!untouched) {
        o.l = q.l; o.th = q.th; o.ph = q.ph;
        o.pl = q.pl / p.delta;
        o.pth = q.pth / p.delta;
    }
    return o;
}

// Guard band of the fast kernel.  The regrouped arithmetic reproduces the photon state of the operation-for-operation
// kernel to ~1e-13 relative (per-step rounding differences of a few 1e-16, carried (sin, cos) drift <= 2e-14 per window),
// multiplied by the trajectory's own error amplification, which explicit Euler in (theta, phi) coordinates makes large
// only where a step's azimuth advance is not small: kappa = max (delta dphi/dlambda)^2, the `stiffness` of
// curvis_ray_record, tracked as the running maximum of w's high word (one integer instruction per step).
// Measured (tools/guard_study.py, profiles/r02_guard_study.json: eight scenes, 10.4 M rays), fast kernel against the
// operation-for-operation kernel, as a function of kappa:
//     kappa < 1     direction of the lookup vector / end-state factor <= 2.3e-12,  |delta l| <= 1.8e-9
//     kappa >= 1    up to 1.2e-6 and 1e-3 |p_l| (|p_l| itself up to 1e31: the ray has been kicked by a coordinate pole)
// kappa < 1: the ray's integers (step count, texel) are accepted when every decision was taken farther from its boundary
// than guard_rel = 1e-9 (x the end-state factor for the direction, x (1 + R) for l: factors 430 and 56 over the maxima
// above); otherwise its index goes to the redo list and the parity kernel re-integrates it.  kappa >= 1 ("kicked", 2.5 % of
// the default frame): counted; re-integrated only with "guard" = 2 (their strict re-integration costs 20 % of the frame
// time: near-critical rays of 10^4 steps dominate a second launch), else kept (1 differing pixel measured in 21 M).
__device__ __forceinline__ double guard_eps(const FrameParams& p, double kappa) {
    if (kappa < 1.0) return p.guard_rel;
    return p.guard_kicked ? __longlong_as_double(0x7ff0000000000000ll) : 0.0;   // kappa NaN counts as kicked
}

// Epilogue of a finished ray of fast_variant 1, out of line (once per ray; the step loop keeps its registers): the photon
// back in the reference's units, the guard-band test, then either the common epilogue (finish_ray) or the redo list.
template <class Shape64>
__device__ __noinline__ void fast_epilogue(const FrameParams& p, const Ray& q, unsigned long long ray, unsigned long long tile_rays,
                                           uint32_t steps, float margin, unsigned wmax_hi, bool guard, RayTally& tally) {
    const double R = p.max_radius;
    const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);   // systems.rs:129-134 on the final state
    const bool untouched = (steps == 0);
    const Ray qe = unscaled_ray(p, q, ray, tile_rays, untouched);
    // stiffness = max (P_phi w)^2 over the steps (w's high word rounded up to the end of its bucket)
    const double wmax = __hiloint2double((int)wmax_hi, (int)0xffffffffu);
    const double sphi = q.pph * wmax;
    const RayDiag diag = {__longlong_as_double(0x7ff8000000000000ll), untouched ? 0.0 : sphi * sphi};
    if (!guard) {
        finish_ray<Shape64, TrigFast, false>(p, qe, side, steps, ray, tally, diag, 0.0);
        return;
    }
    const double eps = guard_eps(p, diag.stiffness);
    if (!(diag.stiffness < 1.0)) atomicAdd(&p.counters->n_kicked, 1ull);
    // the step count: every l visited near the radius stayed farther than eps * (1 + R) from +-R (eps = 0: a kicked ray
    // that is kept as integrated)
    bool accepted = (eps == 0.0) || ((double)margin > eps * (1.0 + fabs(R)));
    if (accepted) accepted = finish_ray<Shape64, TrigFast, true>(p, qe, side, steps, ray, tally, diag, eps);
    if (accepted) return;
    const unsigned long long slot = atomicAdd(&p.counters->n_reintegrated, 1ull);
    if (slot < p.redo_capacity) {
        p.redo_list[slot] = ray;
        return;
    }
    // list full (never with the default capacity of one slot per ray): re-integrate here, in line
    Ray qs;
    new_photon_for_ray(p, ray, tile_rays, qs);
    const SlowResult sr = parity_steps<Shape64>(p, qs, p.max_iterations, (R >= 0.0) ? abs_hi(R) : 0u);
    qs.l = sr.l; qs.th = sr.th; qs.ph = sr.ph; qs.pl = sr.pl; qs.pth = sr.pth;
    const int side2 = (qs.l > R) ? 1 : ((qs.l < -R) ? -1 : 0);
    finish_ray<Shape64, TrigFast, false>(p, qs, side2, (qs.l != qs.l) ? p.max_iterations : sr.steps, ray, tally, diag, 0.0);
}

// MinBlocks: resident CTAs per SM the register allocation is held to — 5 (96 registers, the default: five warps per scheduler
// hide more of the step's dependency chains than four) or 4 (128 registers), kept for A/B ("fast_regs").
// LongFirst: the longest-first refill (below) is compiled into its own instantiation — with the list handling present, even
// unused, the whole-frame kernel was 2 % slower (same step loop, different code around it: 36.65 against 35.87 ms per 4K frame).
template <class Fast, int Variant, int MinBlocks, bool LongFirst>
__global__ void __launch_bounds__(kBlockFast, MinBlocks) render_rows_f64_fast(const __grid_constant__ FrameParams p) {
    using Shape64 = typename Fast::Shape64;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const double R = p.max_radius;
    // escape test: |l| > R needs abs_hi(l) >= hi(R) when R >= 0; for negative or NaN R the gate is open.  Variant 1 also
    // closes it at the end of the Interstellar shape table (+inf for the other metrics).
    const double R_gate = (Variant == 1) ? fmin(R, p.fast_l_limit) : R;
    unsigned gate = (R_gate >= 0.0) ? abs_hi(R_gate) : 0u;
    unsigned gate_key = Fast::gate_key(p, R_gate);         // (Ellis: the same gate on r^2)
    // (code generation only: plain values from here on, else the select above is re-evaluated, as a DSETP, in the Ellis step loop)
    if (Fast::kGateFromSquares) asm volatile("" : "+r"(gate), "+r"(gate_key));
    const bool guard = (Variant == 1) && p.redo_list != nullptr;
    const float finf = __int_as_float(0x7f800000);
    // longest-first list (written by collect_long_rays earlier on the stream); a list that overflowed is ignored
    unsigned long long n_long = 0;
    if (LongFirst) {
        n_long = p.counters->n_long;
        if (n_long > p.long_capacity) n_long = 0;
    }
    // The warp schedulers do not share a saturated fp64 pipe evenly: a warp's share falls with its hardware slot (%warpid).
    // Over a whole 4K frame the five warps of a scheduler executed 1.65 / 1.51 / 1.08 / 0.55 / 0.21 of the mean
    // (tools/scheduler_shares.py, profiles/r02_scheduler_shares.json) — a 20,000-step ray takes 2.7 ms in a warp of the first
    // resident CTA of its SM and 20 ms in one of the fifth.  The listed rays are therefore claimed by the favoured slots first;
    // the other warps take from the list only once the index walk is exhausted (so the list is consumed whatever slots the
    // launch got).
    bool favoured = false;
    if (LongFirst) {
        unsigned warpid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
        favoured = warpid < p.favoured_slots;
    }
    bool long_drained = (n_long == 0);

    // Per-ray state the step loop never reads lives in shared memory (fast_variant 1): the kernel sits at the 96-register
    // limit of five resident CTAs per SM, and every register the loop does not need is one constant it can keep pinned.
    struct Cold {
        double ph, pph;             // phi and P_phi: touched once per window
        unsigned long long ray;     // ray index: read by the epilogue
        float margin;               // smallest distance of an l visited near the radius to +-R (guard band of the step count); 0 = unknown
        unsigned pad;
    };
    __shared__ Cold cold_all[kBlockFast];
    Cold& cold = cold_all[threadIdx.x];

    TrigRegs tr;
    RotRegs rr;
    if (Variant == 0) tr.load();
    else rr.load();
    Ray q;
    int state = 0;            // 0 idle, 1 integrating, 2 finished (epilogue pending)
    bool drained = false;
    uint32_t remaining = 0;
    unsigned long long ray = 0;   // Variant 0 only
    unsigned wmax_hi = 0;     // running max of the high word of w (stiffness monitor)
    RayTally tally;

    for (;;) {
        if (state == 2) {
            const uint32_t steps = p.max_iterations - remaining;
            if (Variant == 1) {
                q.ph = cold.ph; q.pph = cold.pph;
                fast_epilogue<Shape64>(p, q, cold.ray, tile_rays, steps, cold.margin, wmax_hi, guard, tally);
            } else {
                const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);   // systems.rs:129-134 on the final state
                const RayDiag nodiag = {__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll)};
                finish_ray<Shape64, TrigFast, false>(p, q, side, steps, ray, tally, nodiag, 0.0);
            }
            state = 0;
        }

        // ---- refill, longest first.  A pre-pass kernel (collect_long_rays, below) has listed the rays predicted to be long:
        // near-critical photons whose orbit plane almost contains the polar axis (ray_predicted_long, geodesic_f64.cuh) — the
        // rows next to the image's central row take 10x the time of any other (profiles/r02_latency_probe.json).  One queue
        // (counters->long_next) hands out that list, the other (next_ray) walks the ray indices and skips the listed ones; which
        // a warp draws from first depends on its hardware slot (`favoured`, above).  A 20,000-step ray needs 2.7 ms in a favoured
        // slot whenever it starts, so on a small tile (one 4K frame over 8 GPUs: 4.5 ms) it must start at once.  A skipped ticket costs the prediction only (the pixel's unnormalised direction: ~40 instructions), and the
        // lane takes another.  The list must stay SHORT (here 0.3 % of a frame): listing every pole-grazing ray (2.5 %) and
        // starting them all at once cost 2-4 % of the frame — for two generations every warp of the GPU was in the slow,
        // divergent pole-crossing code at the same time, with no regular warps to hide its latency behind.
        unsigned idle = __ballot_sync(kFullFast, state == 0);
        if (idle) {
            if (LongFirst) {
                while (idle && !(drained && long_drained)) {
                    // favoured slots: the list first, then the index walk; the others: the index walk, then what is left of the list
                    const bool from_list = !long_drained && (favoured || drained);
                    unsigned long long* const queue = from_list ? &p.counters->long_next : &p.counters->next_ray;
                    const unsigned long long queue_end = from_list ? n_long : launch_rays;
                    const int leader = __ffs(idle) - 1;
                    unsigned long long base = 0;
                    if ((int)lane == leader) base = atomicAdd(queue, (unsigned long long)__popc(idle));
                    base = __shfl_sync(kFullFast, base, leader);
                    if (state == 0) {
                        const unsigned long long ticket = base + (unsigned long long)__popc(idle & lt_mask);
                        if (ticket < queue_end) {
                            unsigned long long idx = ticket;
                            bool take = true;
                            if (from_list) idx = p.long_list[ticket];
                            else if (n_long) take = !ray_predicted_long(p, idx, tile_rays, kLongRaySin * kLongRaySin);   // (listed: somebody's list ticket)
                            if (take) {
                                new_photon_for_ray(p, idx, tile_rays, q);
                                q.pl = q.pl * p.delta; q.pth = q.pth * p.delta; q.pph = q.pph * p.delta;   // (Variant 1 only)
                                q.pph2 = q.pph * q.pph;
                                cold.ph = q.ph; cold.pph = q.pph; cold.ray = idx; cold.margin = finf;
                                remaining = p.max_iterations;
                                wmax_hi = 0;
                                state = (remaining == 0) ? 2 : 1;
                            }
                        }
                    }
                    if (base + (unsigned long long)__popc(idle) >= queue_end) {
                        if (from_list) long_drained = true;
                        else drained = true;
                    }
                    idle = __ballot_sync(kFullFast, state == 0);
                }
            } else if (!drained) {                 // rays in index order
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(&p.counters->next_ray, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFullFast, base, leader);
                if (state == 0) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < launch_rays) {
                        new_photon_for_ray(p, idx, tile_rays, q);
                        if (Variant == 1) {            // momenta pre-scaled by delta (fast_window_scaled)
                            q.pl = q.pl * p.delta; q.pth = q.pth * p.delta; q.pph = q.pph * p.delta;
                            q.pph2 = q.pph * q.pph;
                            cold.ph = q.ph; cold.pph = q.pph; cold.ray = idx; cold.margin = finf;
                        } else {
                            ray = idx;
                        }
                        remaining = p.max_iterations;
                        wmax_hi = 0;
                        state = (remaining == 0) ? 2 : 1;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFullFast, state != 0) == 0u) break;
        }

        // ---- up to `window` Euler steps (escape_photon's loop body, systems.rs:126-135).  Lanes
        // leave the loop when they escape, run out of iterations or need the parity step; they
        // reconverge after it.
        if (state == 1) {
            const uint32_t n = min(p.window, remaining);
            uint32_t k = 0;
            bool stop = false;
            // huge angles (outside the reduction's range) and non-finite p_theta / p_phi^2 take parity steps
            bool slow = !(abs_hi(q.th) < pow2_hi(30) && abs_hi(q.pth) < pow2_hi(200) && abs_hi(q.pph2) < pow2_hi(200));
            if (!slow && Variant == 0) {
                do {
                    if (!fast_step<Fast>(p, tr, q)) { slow = true; break; }
                    ++k;
                    if (abs_hi(q.l) >= gate) {                                   // within 2^-20 of the radius, or NaN
                        if ((q.l > R) || (q.l < -R) || (q.l != q.l)) { stop = true; break; }   // :129-134
                    }
                } while (k < n);
            }
            if (!slow && Variant == 1) {
                bool near = false;
                double wsum = 0.0, b2 = 0.0, f = 0.0;
                k = fast_window_scaled<Fast>(p, rr, q, n, gate, gate_key, near, slow, wmax_hi, wsum, b2, f);
                cold.ph = fma(cold.pph, wsum, cold.ph);                            // :240 for every step of the window
                if (near) {
                    // The step just taken brought |l| to the radius gate.  Escape test (systems.rs:129-134), and the guard
                    // band of the step count: how far the step landed beyond +-R, and how far inside it started —
                    // l_before = l - P_l_before, P_l_before = P_l - b2 f (the fma of :261 undone to an ulp).  A step that
                    // ends inside the gate's sliver (|l| in [R (1 - 2^-20), R]) records its distance and goes on.
                    const double after = fabs(q.l) - R;
                    if (q.l != q.l) { stop = true; cold.margin = 0.f; }
                    else if (after > 0.0) {
                        const double before = R - fabs(q.l - (q.pl - b2 * f));
                        stop = true;
                        cold.margin = fminf(cold.margin, __double2float_rd(fmin(before, after)));
                    } else {
                        cold.margin = fminf(cold.margin, __double2float_rd(-after));
                        if (Fast::beyond(p, q.l)) slow = true;                     // past the shape table: parity steps
                    }
                }
            }
            if (slow) {
                if (Variant == 1) { cold.margin = 0.f; q.ph = cold.ph; q.pph = cold.pph; }   // a ray that needed parity steps is re-integrated whole
                const SlowResult sr = parity_steps<Shape64>(p, Variant == 1 ? unscaled_ray(p, q, cold.ray, tile_rays, false) : q, n - k,
                                                            (R >= 0.0) ? abs_hi(R) : 0u);
                q.l = sr.l; q.th = sr.th; q.ph = sr.ph; q.pl = sr.pl; q.pth = sr.pth;
                if (Variant == 1) { q.pl = q.pl * p.delta; q.pth = q.pth * p.delta; cold.ph = sr.ph; }
                k += sr.steps;
                stop = sr.stop;
            }
            remaining -= k;
            // A NaN l never compares true and never recovers: the reference would spin through all
            // remaining iterations and return NotEscaped.  Same result, same step count, no spinning.
            if (q.l != q.l) remaining = 0;
            if (stop || remaining == 0) state = 2;                               // :137
        }
        __syncwarp();
    }

    flush_tally(p, tally, lane);
}

// Pre-pass of the longest-first refill: the indices of the rays predicted to graze a coordinate pole, appended to
// FrameParams::long_list in no particular order (one aggregated atomic per warp); counters->n_long counts every such ray,
// listed or not (the render kernel ignores a list that overflowed).
__global__ void __launch_bounds__(256) collect_long_rays(const __grid_constant__ FrameParams p) {
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    // (whole warps iterate together: the bound is rounded up to a multiple of 32)
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < ((launch_rays + 31ull) & ~31ull); i += stride) {
        const bool is_long = i < launch_rays && ray_predicted_long(p, i, tile_rays, kLongRaySin * kLongRaySin);
        const unsigned m = __ballot_sync(kFullFast, is_long);
        if (m) {
            unsigned long long base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(&p.counters->n_long, (unsigned long long)__popc(m));
            base = __shfl_sync(kFullFast, base, leader);
            const unsigned long long slot = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
            if (is_long && slot < p.long_capacity) p.long_list[slot] = i;
        }
    }
}

// When the pre-pass runs ("longest_first" = 2, the default): launches of at least 2^15 rays (the efficient renderer's table skips
// it) and of at most 64 rays per lane of the grid — or of any size where the metric's policy says so (`whole_frames`:
// FastEllis::kLongFirstWholeFrames).  In a launch of 11 rays per lane (a 4K frame over 8 GPUs, 4.5 ms) a 20,000-step straggler must
// start first and in a favoured slot (tile time 4.47 .. 5.83 ms -> 4.50 .. 4.53).  In a launch of 87 rays per lane (a whole 4K frame
// on one GPU) the index order claims it half-way through: in an unfavoured slot it then outlasts the Ellis kernel (35.29 -> 34.89 ms
// with the list), while the Interstellar kernel, 50 % longer, absorbs it and only pays for 27,000 slow rays starting at once (+0.4 ms).
constexpr unsigned long long kLongestFirstMinRays = 1ull << 15;
constexpr unsigned long long kLongestFirstMaxRaysPerLane = 64;

__host__ bool longest_first_wanted(const FrameParams& p, int mode, int sm_count, bool whole_frames) {
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    if (!p.long_list || mode == 0 || rays < kLongestFirstMinRays) return false;
    return mode == 1 || whole_frames || rays <= kLongestFirstMaxRaysPerLane * (unsigned long long)sm_count * 5ull * kBlockFast;
}

template <class Fast, int Variant, int MinBlocks>
cudaError_t launch_fast_variant(const FrameParams& p, int sm_count, int blocks_per_sm_override, int longest_first, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;
    if (blocks_per_sm_auto == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_auto, render_rows_f64_fast<Fast, Variant, MinBlocks, false>, kBlockFast, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm_auto < 1) blocks_per_sm_auto = 1;
    }
    int blocks_per_sm = blocks_per_sm_auto;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < blocks_per_sm) blocks_per_sm = blocks_per_sm_override;
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    unsigned long long want = (rays + kBlockFast - 1) / kBlockFast;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    if (Variant == 1 && longest_first_wanted(p, longest_first, sm_count, Fast::kLongFirstWholeFrames)) {
        const unsigned long long blocks = (rays + 255) / 256;
        collect_long_rays<<<(unsigned)(blocks < 8ull * sm_count ? blocks : 8ull * sm_count), 256, 0, stream>>>(p);
        render_rows_f64_fast<Fast, Variant, MinBlocks, Variant == 1><<<grid, kBlockFast, 0, stream>>>(p);
    } else {
        render_rows_f64_fast<Fast, Variant, MinBlocks, false><<<grid, kBlockFast, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

template <class Fast>
cudaError_t launch_fast(const FrameParams& p, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    // variant 1 scales the momenta by delta: it needs a finite, non-zero step of ordinary magnitude
    const double ad = p.delta < 0.0 ? -p.delta : p.delta;
    if (t.fast_variant == 0 || !(ad >= 0x1p-100 && ad <= 0x1p100)) return launch_fast_variant<Fast, 0, 5>(p, sm_count, t.blocks_per_sm, 0, stream);   // trigonometry from theta every step
    const int regs = t.fast_regs ? t.fast_regs : 96;
    if (regs == 96) return launch_fast_variant<Fast, 1, 5>(p, sm_count, t.blocks_per_sm, t.longest_first, stream);         // default: rotated (sin, cos)
    return launch_fast_variant<Fast, 1, 4>(p, sm_count, t.blocks_per_sm, t.longest_first, stream);
}

}  // namespace

__global__ void debug_shape_kernel(const double2* tab, int which, const double* x, double* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FrameParams p;
    p.shape_tab = tab;
    double F, G;
    shape_fg(p, x[i], F, G);
    out[i] = which ? G : F;
}

__global__ void debug_inverse_shape_kernel(const double2* tab, const double* x, double* y, double* g, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // the lookup reads z = |l| - a: feed |l| = z, a = 0; negative z (the plateau) cannot be expressed by |l| and goes through a
    const double xi = x[i];
    double U, H;
    if (xi >= 0.0 || xi != xi) interstellar_inverse_lookup(tab, 0.0, xi, U, H);
    else interstellar_inverse_lookup(tab, -xi, 0.0, U, H);
    y[i] = U; g[i] = H;
}

cudaError_t launch_debug_inverse_shape(const double2* tab, const double* x, double* y, double* g, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    debug_inverse_shape_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tab, x, y, g, n);
    return cudaGetLastError();
}

cudaError_t launch_debug_shape(const double2* tab, int which, const double* x, double* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    debug_shape_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tab, which, x, out, n);
    return cudaGetLastError();
}

// The longest-first pre-pass for the operation-for-operation kernel (render_f64.cu: the same list, the same rule).
bool longest_first_prepass_wanted(const FrameParams& p, int mode, int sm_count, bool whole_frames) {
    return longest_first_wanted(p, mode, sm_count, whole_frames);
}

cudaError_t launch_collect_long_rays(const FrameParams& p, int sm_count, cudaStream_t stream) {
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    const unsigned long long blocks = (rays + 255) / 256;
    collect_long_rays<<<(unsigned)(blocks < 8ull * sm_count ? blocks : 8ull * sm_count), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// Whether launch_render_f64_fast launches the longest-first pre-pass in front of the render kernel (the launch counter's business).
bool render_f64_fast_has_prepass(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count) {
    const double ad = p.delta < 0.0 ? -p.delta : p.delta;
    return t.fast_variant == 1 && (ad >= 0x1p-100 && ad <= 0x1p100) && longest_first_wanted(p, t.longest_first, sm_count, metric_kind == CURVIS_METRIC_ELLIS /* FastEllis::kLongFirstWholeFrames */);
}

cudaError_t launch_render_f64_fast(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch_fast<FastEllis>(p, t, sm_count, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch_fast<FastInterstellar>(p, t, sm_count, stream);
    case CURVIS_METRIC_FLAT: return launch_fast<FastFlat>(p, t, sm_count, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace curvis
