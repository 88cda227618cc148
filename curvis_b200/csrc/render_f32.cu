// render_f32.cu — CURVIS_PRECISION_F32, the opt-in fast mode (extension; NOT the parity path).
//
// Same ODE, same forward-Euler scheme, same execution model as render_f64.cu (persistent grid,
// one ray per lane, windowed ballot refill), but the right-hand side of
// update_relativistic_object (reference src/metrics.rs:223-270) is evaluated in fp32 and
// regrouped like CURVIS_PRECISION_F64_FAST (render_f64_fast.cu):
//   * ONE MUFU.RCP per step, w = 1/(r^2 sin^2 theta); 1/r^2 = w sin^2, 1/sin^2 = w r^2; Ellis needs no
//     square root (r'/r^3 = l/r^4);
//   * momenta pre-scaled by delta (P = delta p — the same Euler iteration with the affine parameter
//     rescaled to unit steps), so delta leaves the loop;
//   * (sin theta, cos theta) carried along and rotated by the step's dtheta in the three-shear form
//     (c -= t s; s += sd c; c -= t s with sd = sin dtheta, t = tan(dtheta/2), two-term series for
//     |dtheta| < 2^-4), re-derived from the compensated theta at the start of every window (32
//     steps) and after a larger dtheta, so the rotation's fp32 rounding never accumulates along the ray;
//   * the five state variables (l, theta, phi, P_l, P_theta) are accumulated with Kahan
//     compensation (sum + carry in fp32), so 2000 increments of ~0.05 do not random-walk the
//     low bits of l ~ 100 — this is what keeps the step count and the end direction close to
//     the fp64 path;
//   * one exit branch per step (step budget, |l| near the radius or NaN, |dtheta| >= 2^-4), sorted
//     out after the loop; the exact escape test runs in fp64 on the compensated l;
//   * ray generation and the escaped-photon epilogue (direction, acos/atan2, texel index) reuse
//     the fp64 code of geodesic_f64.cuh: they run once per ray.
// An fp32 instruction issues every cycle per scheduler where an fp64 one holds the dispatch port
// for two or more (profiles/r01_microbench_fp64_pipe.txt).
//
// Results are NOT bit-comparable with the reference: tests/test_gpu_fast_mode.py states the
// tolerance (escape side identical, end direction within 1e-5 rad on >= 99 % of rays, texel equal
// or adjacent) and bench.py reports the measured deviation next to the throughput.
#include "geodesic_f64.cuh"
#include "launch.h"
#include "shape_table.h"

namespace curvis {

namespace {

constexpr int kBlock32 = 128;
constexpr unsigned kFull32 = 0xffffffffu;

struct Kahan {  // value ~ s - c
    float s, c;
    __device__ __forceinline__ void set(double v) { s = (float)v; c = (float)((double)s - v); }
    __device__ __forceinline__ void add(float inc) {
        const float y = inc - c;
        const float t = s + y;
        c = (t - s) - y;
        s = t;
    }
    __device__ __forceinline__ double value() const { return (double)s - (double)c; }
};

__device__ __forceinline__ float rcp_fast(float x) {   // one MUFU.RCP, ~1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Exact escape test on the compensated radial coordinate; out of line so the fp64 compares stay
// off the per-step instruction stream (it runs only within a few steps of the escape radius).
__device__ __noinline__ bool escaped_exact(float ls, float lc, double R) {
    const double lv = (double)ls - (double)lc;
    return (lv > R) || (lv < -R);
}

// sin/cos for |x| < 2^16 (the caller guarantees it; larger or non-finite angles go to sincos_library).
__device__ __forceinline__ void sincos_f32(float x, float& s, float& c) {
    const float t = fmaf(x, 0.63661975f, 12582912.0f);   // 1.5*2^23: integer in the low mantissa bits
    const int k = __float_as_int(t);
    const float q = t - 12582912.0f;
    float r = fmaf(-q, 1.5707963705062866f, x);
    r = fmaf(-q, -4.371138828673793e-08f, r);
    r = fmaf(-q, -1.7151245100058819e-15f, r);
    const float u = r * r;
    float sp = fmaf(u, 2.723765874179662e-06f, -0.0001983999100048095f);
    float cp = fmaf(u, 2.4537857825635e-05f, -0.001388825592584908f);
    sp = fmaf(u, sp, 0.008333331905305386f);
    cp = fmaf(u, cp, 0.0416666641831398f);
    sp = fmaf(u, sp, -0.1666666716337204f);
    const float sr = fmaf(r * u, sp, r);
    const float cr = fmaf(u * u, cp, fmaf(u, -0.5f, 1.0f));
    const float a = (k & 1) ? cr : sr;
    const float b = (k & 1) ? sr : cr;
    s = __int_as_float(__float_as_int(a) ^ ((k & 2) << 30));
    c = __int_as_float(__float_as_int(b) ^ (((k + 1) & 2) << 30));
}

// Large or non-finite angles: the library (Payne-Hanek), out of line.
__device__ __noinline__ float2 sincos_library(float x) {
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

// Shape policies, split around the reciprocal like render_f64_fast.cu: prepare() returns the divisor
// d (r^2 sin^2 for Ellis, r sin^2 otherwise), finish() turns y0 = 1/d into w = 1/(r^2 sin^2),
// u = 1/r^2, v = 1/sin^2 and f = r'(l)/r(l)^3.
struct Shape32Ellis {   // metrics.rs:417-421
    using Shape64 = ShapeEllis;
    struct Pre { float r2; };
    static __device__ __forceinline__ float prepare(const FrameParams& p, float l, float s2, Pre& pre) {
        pre.r2 = fmaf(l, l, p.f_rho2);
        return pre.r2 * s2;
    }
    static __device__ __forceinline__ void finish(const Pre& pre, float y0, float l, float s2, float& w, float& u, float& v, float& f) {
        w = y0;
        u = w * s2;
        v = w * pre.r2;
        f = l * (u * u);             // (l/r) / r^3
    }
};

struct Pre32FromR { float r, rp; };
__device__ __forceinline__ void finish32_from_r(const Pre32FromR& pre, float y0, float s2, float& w, float& u, float& v, float& f) {
    const float y = y0 * s2;         // 1/r
    v = y0 * pre.r;                  // 1/sin^2
    u = y * y;
    w = u * v;
    f = pre.rp * (y * u);            // r'/r^3
}

// F(x) = x atan x - ln(1+x^2)/2 and G(x) = (2/pi) atan x, x > 0 (shape_table.h): inside [2^-10, 2^16) two degree-3
// polynomials in t = x - interval midpoint, coefficients from two 128-bit loads; outside, the library; x <= 0 (the
// plateau |l| <= a) and NaN give F = G = 0, i.e. r = rho, r' = 0.
__device__ __noinline__ float2 shape32_library(float x) {
    if (!(x > 0.0f)) return make_float2(0.0f, 0.0f);
    const float at = atanf(x);
    return make_float2(fmaf(x, at, -0.5f * log1pf(x * x)), 0.63661975f * at);
}

__device__ __forceinline__ void shape32_fg(const float4* tab, float x, float& F, float& G) {
    const unsigned bits = (unsigned)__float_as_int(x);
    const unsigned idx = (bits >> kShapeTab32Shift) - kShapeTab32Base;
    if (idx < (unsigned)kShapeTab32Intervals) {
        const float c = __int_as_float((int)((bits & ~((1u << kShapeTab32Shift) - 1u)) | (1u << (kShapeTab32Shift - 1))));
        const float t = x - c;
        const float4 a = __ldg(tab + 2 * idx), b = __ldg(tab + 2 * idx + 1);
        F = fmaf(t, fmaf(t, fmaf(t, a.w, a.z), a.y), a.x);
        G = fmaf(t, fmaf(t, fmaf(t, b.w, b.z), b.y), b.x);
    } else {
        const float2 fg = shape32_library(x);
        F = fg.x; G = fg.y;
    }
}

struct Shape32Interstellar {   // metrics.rs:461-485
    using Shape64 = ShapeInterstellar;
    using Pre = Pre32FromR;
    static __device__ __forceinline__ float prepare(const FrameParams& p, float l, float s2, Pre& pre) {
        const float x = (fabsf(l) - p.f_a) * p.f_xscale;      // <= 0 on the plateau: F = G = 0 from the library branch
        float F, G;
        shape32_fg(p.shape_tab32, x, F, G);
        pre.r = fmaf(p.f_m, F, p.f_rho);
        pre.rp = copysignf(G, l);
        return pre.r * s2;
    }
    static __device__ __forceinline__ void finish(const Pre& pre, float y0, float, float s2, float& w, float& u, float& v, float& f) {
        finish32_from_r(pre, y0, s2, w, u, v, f);
    }
};

struct Shape32Flat {   // metrics.rs:501-505
    using Shape64 = ShapeFlat;
    using Pre = Pre32FromR;
    static __device__ __forceinline__ float prepare(const FrameParams&, float l, float s2, Pre& pre) {
        pre.r = l; pre.rp = 1.0f;
        return l * s2;
    }
    static __device__ __forceinline__ void finish(const Pre& pre, float y0, float, float s2, float& w, float& u, float& v, float& f) {
        finish32_from_r(pre, y0, s2, w, u, v, f);
    }
};

// Compensated state of one ray; momenta scaled by delta (see the file header).
struct Ray32 {
    Kahan l, th, ph, pl, pth;
    float pph, pph2;
};

// Up to n steps of one window.  Returns the number of steps taken; `stop` = the ray left the radius
// (systems.rs:129-134, tested in fp64 on the compensated l) or its l became NaN.
template <class Shape32>
__device__ __forceinline__ uint32_t window_f32(const FrameParams& p, Ray32& q, uint32_t n, float near_radius, double R, bool& stop) {
    uint32_t k = 0;
    for (;;) {
        // (sin, cos) of the compensated angle
        float sn, cn;
        if (fabsf(q.th.s) < 65536.0f) sincos_f32(q.th.s, sn, cn);
        else { const float2 sc = sincos_library(q.th.s); sn = sc.x; cn = sc.y; }
        { const float s0 = sn; sn = fmaf(-cn, q.th.c, sn); cn = fmaf(s0, q.th.c, cn); }      // theta = th.s - th.c
        typename Shape32::Pre pre;
        float s2 = sn * sn;
        float d = Shape32::prepare(p, q.l.s, s2, pre);
        float dth;
        // one step; true = leave the loop (one of the rarely-true conditions holds)
        auto step = [&]() -> bool {
            float w, u, v, f;
            Shape32::finish(pre, rcp_fast(d), q.l.s, s2, w, u, v, f);
            const float cs = sn * cn;
            dth = q.pth.s * u;                                  // metrics.rs:239 (times delta)
            const float pv = q.pph2 * v;                        // p_phi^2 / sin^2
            const float b2 = fmaf(q.pth.s, q.pth.s, pv);        // :257
            q.l.add(q.pl.s);                                    // :238, :295 (old p_l)
            q.th.add(dth);
            q.ph.add(q.pph * w);                                // :240
            q.pl.add(b2 * f);                                   // :261, :296
            q.pth.add((pv * cs) * w);                           // :262  p_phi^2 cos / (r^2 sin^3)
            // rotate (sin, cos) by dth: three shears, sd = sin dth, t = tan(dth/2)
            const float v2 = dth * dth;
            const float sd = dth * fmaf(v2, fmaf(v2, 0.0083333333f, -0.16666667f), 1.0f);
            const float t = dth * fmaf(v2, fmaf(v2, 0.0041666667f, 0.041666667f), 0.5f);
            cn = fmaf(-t, sn, cn);
            sn = fmaf(sd, cn, sn);
            cn = fmaf(-t, sn, cn);
            ++k;
            s2 = sn * sn;
            d = Shape32::prepare(p, q.l.s, s2, pre);
            return (k >= n) | !(fabsf(q.l.s) < near_radius) | !(fabsf(dth) < 0.0625f);
        };
        for (;;) {               // unrolled by two: the Kahan sums alternate registers instead of copying them
            if (step()) break;
            if (step()) break;
        }
        if (!(fabsf(q.l.s) < near_radius)) {                    // near the radius, or NaN
            if (q.l.s != q.l.s || escaped_exact(q.l.s, q.l.c, R)) { stop = true; return k; }
        }
        if (k >= n) return k;
        // |dtheta| too large for the rotation (or NaN), or a false alarm near the radius: re-derive (sin, cos) and go on
    }
}

template <class Shape32>
__global__ void __launch_bounds__(kBlock32) render_rows_f32(const __grid_constant__ FrameParams p) {
    using Shape64 = typename Shape32::Shape64;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const double R = p.max_radius;
    const float near_radius = p.f_near_radius;   // below it no escape test is needed

    Ray32 q;
    q.pph = 0.f; q.pph2 = 0.f;
    int state = 0;
    bool drained = false;
    uint32_t remaining = 0;
    unsigned long long ray = 0;
    RayTally tally;

    for (;;) {
        if (state == 2) {
            // epilogue in fp64 on the compensated state: same code path as the parity kernel.  p_phi is
            // conserved: it is regenerated from the ray index (exact) instead of un-scaled.
            Ray e;
            new_photon_for_ray(p, ray, tile_rays, e);
            if (remaining != p.max_iterations && p.delta != 0.0) {   // (a zero step never moves the photon)
                e.l = q.l.value(); e.th = q.th.value(); e.ph = q.ph.value();
                e.pl = q.pl.value() / p.delta; e.pth = q.pth.value() / p.delta;
            }
            const int side = (e.l > R) ? 1 : ((e.l < -R) ? -1 : 0);
            const RayDiag nodiag = {__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll)};
            finish_ray<Shape64, TrigFast, false>(p, e, side, p.max_iterations - remaining, ray, tally, nodiag, 0.0);
            state = 0;
        }

        const unsigned idle = __ballot_sync(kFull32, state == 0);
        if (idle) {
            if (!drained) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(&p.counters->next_ray, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFull32, base, leader);
                if (state == 0) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < launch_rays) {
                        ray = idx;
                        Ray e;
                        new_photon_for_ray(p, idx, tile_rays, e);   // fp64 ray generation (once per ray)
                        q.l.set(e.l); q.th.set(e.th); q.ph.set(e.ph);
                        q.pl.set(e.pl * p.delta); q.pth.set(e.pth * p.delta);
                        const double pphd = e.pph * p.delta;
                        q.pph = (float)pphd; q.pph2 = (float)(pphd * pphd);
                        remaining = p.max_iterations;
                        state = (remaining == 0) ? 2 : 1;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFull32, state != 0) == 0u) break;
        }

        if (state == 1) {
            bool stop = false;
            const uint32_t k = window_f32<Shape32>(p, q, min(p.window, remaining), near_radius, R, stop);
            remaining -= k;
            // a NaN l never escapes: NotEscaped with all iterations counted (as the reference would, after spinning)
            if (q.l.s != q.l.s) remaining = 0;
            if (stop || remaining == 0) state = 2;
        }
        __syncwarp();
    }

    flush_tally(p, tally, lane);
}

template <class Shape32>
cudaError_t launch32(const FrameParams& p, int sm_count, int blocks_per_sm_override, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;
    if (blocks_per_sm_auto == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_auto, render_rows_f32<Shape32>, kBlock32, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm_auto < 1) blocks_per_sm_auto = 1;
    }
    int blocks_per_sm = blocks_per_sm_auto;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < blocks_per_sm) blocks_per_sm = blocks_per_sm_override;
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    unsigned long long want = (rays + kBlock32 - 1) / kBlock32;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    render_rows_f32<Shape32><<<grid, kBlock32, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

__global__ void debug_shape32_kernel(const float4* tab, int which, const double* x, double* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float F, G;
    shape32_fg(tab, (float)x[i], F, G);
    out[i] = (double)(which ? G : F);
}

cudaError_t launch_debug_shape32(const float4* tab, int which, const double* x, double* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    debug_shape32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tab, which, x, out, n);
    return cudaGetLastError();
}

cudaError_t launch_render_f32(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch32<Shape32Ellis>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch32<Shape32Interstellar>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_FLAT: return launch32<Shape32Flat>(p, sm_count, t.blocks_per_sm, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace curvis
