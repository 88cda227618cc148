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

namespace curvis {

namespace {

constexpr int kBlockFast = 128;
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

struct FastInterstellar {   // metrics.rs:461-485 with the uniform divisor pi*m folded into d_xscale
    using Shape64 = ShapeInterstellar;
    static __device__ __forceinline__ bool factors(const FrameParams& p, double l, double s2, double& w, double& u, double& v, double& ud, double& fd) {
        const double al = fabs(l);
        double r = p.rho, rp = 0.0;
        if (al > p.a) {
            const double x = (al - p.a) * p.d_xscale;
            const double at = atan(x);
            r = fma(p.m, fma(x, at, -0.5 * log(fma(x, x, 1.0))), p.rho);
            rp = copysign((2.0 / CURVIS_PI) * at, l);
        }
        return factors_from_r(p, r, rp, s2, w, u, v, ud, fd);
    }
};

struct FastFlat {   // metrics.rs:501-505: r = l, r' = 1 (r may be negative: take the parity step then)
    using Shape64 = ShapeFlat;
    static __device__ __forceinline__ bool factors(const FrameParams& p, double l, double s2, double& w, double& u, double& v, double& ud, double& fd) {
        return factors_from_r(p, l, 1.0, s2, w, u, v, ud, fd);
    }
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

// The same step with (sin theta, cos theta) carried as state (fast_f64.cuh: rotate_sincos): the
// pair is rotated by the step's dtheta unconditionally; the caller re-derives it from theta when
// |dtheta| >= 2^-4 (`dth` is returned for that test).  Returns false, leaving the state untouched,
// when an operand is outside the safe window.
template <class Fast>
__device__ __forceinline__ bool fast_step_rot(const FrameParams& p, const RotRegs& rr, Ray& q, double& s, double& c, double& dth) {
    const double s2 = s * s, cs = s * c;
    double w, u, v, ud, fd;
    if (!Fast::factors(p, q.l, s2, w, u, v, ud, fd)) return false;
    dth = q.pth * ud;                                       // :239 times delta
    const double pv = q.pph2 * v;                           // p_phi^2 / sin^2
    const double b2 = fma(q.pth, q.pth, pv);                // metrics.rs:257
    const double wd = w * p.delta;
    q.l = fma(q.pl, p.delta, q.l);                          // :238, :295
    q.th = q.th + dth;
    q.ph = fma(q.pph, wd, q.ph);                            // :240
    q.pl = fma(b2, fd, q.pl);                               // :261, :296
    q.pth = fma(pv * cs, wd, q.pth);                        // :262
    rotate_sincos(rr, dth, s, c);
    return true;
}

// (sin, cos)(theta) for the rare re-derivation inside the loop; out of line, by value.
__device__ __noinline__ double2 sincos_pair(double th) {
    double s, c;
    sincos_fast(th, s, c);
    return make_double2(s, c);
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

template <class Fast, int Variant>
__global__ void __launch_bounds__(kBlockFast) render_rows_f64_fast(const __grid_constant__ FrameParams p) {
    using Shape64 = typename Fast::Shape64;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const double R = p.max_radius;
    // |l| > R needs abs_hi(l) >= hi(R) when R >= 0; for negative or NaN R the gate is open.
    const unsigned gate = (R >= 0.0) ? abs_hi(R) : 0u;

    TrigRegs tr;
    RotRegs rr;
    if (Variant == 0) tr.load();
    else rr.load();
    Ray q;
    int state = 0;            // 0 idle, 1 integrating, 2 finished (epilogue pending)
    bool drained = false;
    uint32_t remaining = 0;
    unsigned long long ray = 0;
    RayTally tally;

    for (;;) {
        if (state == 2) {
            const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);   // systems.rs:129-134 on the final state
            finish_ray<Shape64, TrigFast>(p, q, side, p.max_iterations - remaining, ray, tally);
            state = 0;
        }

        const unsigned idle = __ballot_sync(kFullFast, state == 0);
        if (idle) {
            if (!drained) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(&p.counters->next_ray, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFullFast, base, leader);
                if (state == 0) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < launch_rays) {
                        ray = idx;
                        new_photon_for_ray(p, idx, tile_rays, q);
                        remaining = p.max_iterations;
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
                double sn, cn, dth;
                sincos_fast(q.th, sn, cn);                                       // once per window, then rotated
                do {
                    if (!fast_step_rot<Fast>(p, rr, q, sn, cn, dth)) { slow = true; break; }
                    ++k;
                    // one rarely-taken branch for both per-step tests: near the escape radius (or NaN), and
                    // a dtheta too large for the rotation (or NaN)
                    const bool big = abs_hi(dth) >= pow2_hi(-4);
                    if (big || abs_hi(q.l) >= gate) {
                        if ((q.l > R) || (q.l < -R) || (q.l != q.l)) { stop = true; break; }   // :129-134
                        if (big) {
                            if (abs_hi(q.th) >= pow2_hi(30)) { slow = (k < n); break; }
                            const double2 sc = sincos_pair(q.th);
                            sn = sc.x; cn = sc.y;
                        }
                    }
                } while (k < n);
            }
            if (slow) {
                const SlowResult sr = parity_steps<Shape64>(p, q, n - k, gate);
                q.l = sr.l; q.th = sr.th; q.ph = sr.ph; q.pl = sr.pl; q.pth = sr.pth;
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

template <class Fast, int Variant>
cudaError_t launch_fast_variant(const FrameParams& p, int sm_count, int blocks_per_sm_override, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;
    if (blocks_per_sm_auto == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_auto, render_rows_f64_fast<Fast, Variant>, kBlockFast, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm_auto < 1) blocks_per_sm_auto = 1;
    }
    int blocks_per_sm = blocks_per_sm_auto;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < blocks_per_sm) blocks_per_sm = blocks_per_sm_override;
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    unsigned long long want = (rays + kBlockFast - 1) / kBlockFast;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    render_rows_f64_fast<Fast, Variant><<<grid, kBlockFast, 0, stream>>>(p);
    return cudaGetLastError();
}

template <class Fast>
cudaError_t launch_fast(const FrameParams& p, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    if (t.fast_variant == 0) return launch_fast_variant<Fast, 0>(p, sm_count, t.blocks_per_sm, stream);   // trigonometry from theta every step
    return launch_fast_variant<Fast, 1>(p, sm_count, t.blocks_per_sm, stream);                            // default: rotated (sin, cos)
}

}  // namespace

cudaError_t launch_render_f64_fast(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch_fast<FastEllis>(p, t, sm_count, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch_fast<FastInterstellar>(p, t, sm_count, stream);
    case CURVIS_METRIC_FLAT: return launch_fast<FastFlat>(p, t, sm_count, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace curvis
