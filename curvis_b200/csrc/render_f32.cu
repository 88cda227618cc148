// render_f32.cu — CURVIS_PRECISION_F32, the opt-in fast mode (extension; NOT the parity path).
//
// Same ODE, same forward-Euler scheme, same execution model as render_f64.cu (persistent grid,
// one ray per lane, windowed ballot refill), but the right-hand side of
// update_relativistic_object (reference src/metrics.rs:223-270) is evaluated in fp32 and
// algebraically regrouped for the hardware:
//   * Ellis needs only two reciprocals per step (1/r^2 and 1/sin^2 theta, MUFU.RCP): r'/r^3 = l/r^4,
//     cos/(r^2 sin^3) = cos*sin/(r^2 sin^4); no square root, no division;
//   * sin/cos: Cody-Waite reduction + degree-9/8 minimax polynomials (tools/gen_trig_coeffs_f32.py);
//   * the five state variables (l, theta, phi, p_l, p_theta) are accumulated with Kahan
//     compensation (sum + carry in fp32), so 2000 increments of ~0.05 do not random-walk the
//     low bits of l ~ 100 — this is what keeps the step count and the end direction close to
//     the fp64 path;
//   * ray generation and the escaped-photon epilogue (direction, acos/atan2, texel index) reuse
//     the fp64 code of geodesic_f64.cuh: they run once per ray.
// An fp32 instruction issues every cycle per scheduler where an fp64 one holds the dispatch port
// for two or more (profiles/r01_microbench_fp64_pipe.txt), hence the ~4x.
//
// Results are NOT bit-comparable with the reference: tests/test_gpu_fast_mode.py states the
// tolerance (escape side identical, end direction within 1e-5 rad on >= 99 % of rays, texel equal
// or adjacent) and bench.py reports the measured deviation next to the throughput.
#include "geodesic_f64.cuh"
#include "launch.h"

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

// sin/cos for |x| < 2^16 (fast path), otherwise the library.
__device__ __forceinline__ void sincos_f32(float x, float& s, float& c) {
    if (fabsf(x) < 65536.0f) {
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
    } else {
        sincosf(x, &s, &c);
    }
}

// Shape functions in fp32: 1/r^2 and the radial-force coefficient r'(l)/r(l)^3.
struct Shape32Ellis {
    using Shape64 = ShapeEllis;
    __device__ __forceinline__ void init(const FrameParams&) {}
    __device__ __forceinline__ void eval(const FrameParams& p, float l, float& inv_r2, float& force) const {
        inv_r2 = rcp_fast(fmaf(l, l, p.f_rho2));
        force = l * (inv_r2 * inv_r2);   // (l/r) / r^3
    }
};

struct Shape32Interstellar {
    using Shape64 = ShapeInterstellar;
    __device__ __forceinline__ void init(const FrameParams&) {}
    __device__ __forceinline__ void eval(const FrameParams& p, float l, float& inv_r2, float& force) const {
        const float al = fabsf(l);
        float r = p.f_rho, rp = 0.0f;
        if (al > p.f_a) {
            const float x = (al - p.f_a) * p.f_xscale;
            const float at = atanf(x);
            r = fmaf(p.f_m, fmaf(x, at, -0.5f * __logf(fmaf(x, x, 1.0f))), p.f_rho);
            rp = copysignf(0.63661975f * at, l);
        }
        const float inv_r = rcp_fast(r);
        inv_r2 = inv_r * inv_r;
        force = rp * (inv_r2 * inv_r);
    }
};

struct Shape32Flat {
    using Shape64 = ShapeFlat;
    __device__ __forceinline__ void init(const FrameParams&) {}
    __device__ __forceinline__ void eval(const FrameParams&, float l, float& inv_r2, float& force) const {
        const float inv_r = rcp_fast(l);
        inv_r2 = inv_r * inv_r;
        force = inv_r2 * inv_r;
    }
};

template <class Shape32>
__global__ void __launch_bounds__(kBlock32) render_rows_f32(const __grid_constant__ FrameParams p) {
    using Shape64 = typename Shape32::Shape64;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const double R = p.max_radius;
    const float delta = p.f_delta;
    const float near_radius = p.f_near_radius;   // below it no escape test is needed
    Shape32 shape;
    shape.init(p);

    Kahan l, th, ph, pl, pth;
    float pph = 0.f, pph2 = 0.f;
    double pph_exact = 0.0;
    int state = 0;
    bool drained = false;
    uint32_t remaining = 0;
    unsigned long long ray = 0;
    RayTally tally;

    for (;;) {
        if (state == 2) {
            // epilogue in fp64 on the compensated state: same code path as the parity kernel
            Ray q;
            q.l = l.value(); q.th = th.value(); q.ph = ph.value(); q.pl = pl.value(); q.pth = pth.value();
            q.pph = pph_exact; q.pph2 = pph_exact * pph_exact;
            const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);
            finish_ray<Shape64, TrigFast>(p, q, side, p.max_iterations - remaining, ray, tally);
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
                        Ray q;
                        new_photon_for_ray(p, idx, tile_rays, q);   // fp64 ray generation (once per ray)
                        l.set(q.l); th.set(q.th); ph.set(q.ph); pl.set(q.pl); pth.set(q.pth);
                        pph_exact = q.pph;
                        pph = (float)q.pph; pph2 = (float)q.pph2;
                        remaining = p.max_iterations;
                        state = (remaining == 0) ? 2 : 1;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFull32, state != 0) == 0u) break;
        }

#pragma unroll 2
        for (uint32_t k = 0; k < p.window; ++k) {
            if (state == 1) {
                // ---- one forward-Euler step, fp32 right-hand side (metrics.rs:223-270 regrouped)
                float s, c;
                sincos_f32(th.s, s, c);
                float inv_r2, force;
                shape.eval(p, l.s, inv_r2, force);
                const float inv_s2 = rcp_fast(s * s);
                const float ir2d = inv_r2 * delta;                   // delta folded into the shared factors
                const float w = pph2 * inv_s2;                       // p_phi^2 / sin^2
                const float b2 = fmaf(pth.s, pth.s, w);              // :257
                const float inc_l = pl.s * delta;                    // :295 (old p_l)
                const float inc_th = pth.s * ir2d;                   // :239
                const float inc_ph = pph * (ir2d * inv_s2);          // :240
                const float inc_pl = b2 * (force * delta);           // :261
                const float inc_pth = (w * inv_s2) * (c * s) * ir2d; // :262  p_phi^2 cos / (r^2 sin^3)
                l.add(inc_l);
                th.add(inc_th);
                ph.add(inc_ph);
                pl.add(inc_pl);                                      // :296
                pth.add(inc_pth);
                --remaining;
                bool done = (remaining == 0);
                if (!(fabsf(l.s) < near_radius)) {                  // near the radius, or NaN
                    done = done || escaped_exact(l.s, l.c, R);
                    if (l.s != l.s) { remaining = 0; done = true; }   // NaN never escapes: NotEscaped with all iterations counted
                }
                if (done) state = 2;
            }
        }
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

cudaError_t launch_render_f32(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch32<Shape32Ellis>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch32<Shape32Interstellar>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_FLAT: return launch32<Shape32Flat>(p, sm_count, t.blocks_per_sm, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace curvis
