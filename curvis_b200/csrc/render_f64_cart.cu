// render_f64_cart.cu — CURVIS_COORDINATES_CARTESIAN ("pole-safe" extension, no reference counterpart; its oracle is
// the CPU restatement oracle_escape_photon_cart of the test oracle, same operation order).
//
// The reference integrates the angular motion in (theta, phi) with momenta (p_theta, p_phi): the right-hand side
// (src/metrics.rs:257-262) carries 1/sin^2 theta and cos theta / sin^3 theta, and explicit Euler at the default step
// turns every ray that passes near theta = 0 or pi — ~11 % of the default frame, SURVEY.md 7a — into a "kicked" ray
// whose end state is an artefact of the coordinate pole.  The metric is spherically symmetric, so the same geodesic
// equations can be written without any chart of the sphere: with n the unit position vector and J = r^2 n x dn/dlambda
// the (conserved) angular-momentum vector,
//     dl/dlambda = p_l          dp_l/dlambda = |J|^2 r'(l) / r(l)^3          dn/dlambda = (J x n) / r(l)^2
// (the first two are metrics.rs:238 and :261 with b^2 = p_theta^2 + p_phi^2/sin^2 theta = |J|^2; the third is
// :239-:240 in vector form).  Same explicit Euler scheme, same step, same escape test (systems.rs:126-135); no
// trigonometry and no pole in the loop: 3 + 1 + 1 state variables, ~25 flop per step.  The end state is converted back
// to the reference's variables (theta, phi, p_theta = r t.e_theta, p_phi = J_z) and goes through the common epilogue
// (finish_ray), so every frame / sampling mode applies unchanged.
//
// Compiled with -fmad=false like the other fp64 TUs: with IEEE +,-,*,/ and sqrt only in the loop, the Ellis photon
// state equals the oracle's bit for bit.
#include "geodesic_f64.cuh"
#include "fast_f64.cuh"
#include "launch.h"

namespace curvis {

namespace {

constexpr int kBlockCart = 128;
constexpr unsigned kFullCart = 0xffffffffu;

struct RayCart {
    double l, pl;
    double nx, ny, nz;
    double jx, jy, jz, l2;   // conserved
};

// The photon of camera_pixels_x_y_to_photon (systems.rs:531-534) in chart-free form.  The tangent-space direction
// d = (d_l, d_theta, d_phi) (metrics.rs:320 normalises it) at the camera position: velocity d_l along n, tangential
// velocity t = d_theta e_theta + d_phi e_phi; J = r n x t  (|J|^2 = r^2 (d_theta^2 + d_phi^2) = p_theta^2 + p_phi^2/sin^2).
__device__ __forceinline__ void new_photon_cart(const CameraBlock& cam, double dx, double dy, double dz, RayCart& q) {
    const double n = norm3(dx, dy, dz);
    dx = dx / n; dy = dy / n; dz = dz / n;
    const double tx = dy * cam.cam_eth[0] + dz * cam.cam_eph[0];
    const double ty = dy * cam.cam_eth[1] + dz * cam.cam_eph[1];
    const double tz = dy * cam.cam_eth[2] + dz * cam.cam_eph[2];
    q.l = cam.cam_pos[1];
    q.pl = dx;
    q.nx = cam.cam_n[0]; q.ny = cam.cam_n[1]; q.nz = cam.cam_n[2];
    q.jx = cam.cam_r * (q.ny * tz - q.nz * ty);
    q.jy = cam.cam_r * (q.nz * tx - q.nx * tz);
    q.jz = cam.cam_r * (q.nx * ty - q.ny * tx);
    q.l2 = (q.jx * q.jx + q.jy * q.jy) + q.jz * q.jz;
}

__device__ __forceinline__ void new_photon_cart_for_ray(const FrameParams& p, unsigned long long idx, unsigned long long tile_rays, RayCart& q) {
    double dx, dy, dz;
    if (p.ray_dirs) {
        new_photon_cart(p.cam, p.ray_dirs[3 * idx], p.ray_dirs[3 * idx + 1], p.ray_dirs[3 * idx + 2], q);
    } else {
        unsigned long long f; uint32_t px, py, k;
        split_ray_index(p, idx, tile_rays, f, px, k);
        tile_pixel(p, k, px, px, py);
        const CameraBlock& cam = camera_of_frame(p, f);
        outward_vector_on_world_space(cam, p.frame_width, p.height, px, py, dx, dy, dz);
        new_photon_cart(cam, dx, dy, dz, q);
    }
}

// One explicit Euler step: every derivative at the old state, then x += dx * delta.
template <class Shape>
__device__ __forceinline__ void euler_step_cart(const FrameParams& p, RayCart& q) {
    double r, r2, rp;
    Shape::eval(p, q.l, r, r2, rp);
    const double u = 1.0 / r2;
    const double cx = q.jy * q.nz - q.jz * q.ny;
    const double cy = q.jz * q.nx - q.jx * q.nz;
    const double cz = q.jx * q.ny - q.jy * q.nx;
    const double dpl = (q.l2 * rp) / ((r * r) * r);
    q.nx = q.nx + (cx * u) * p.delta;
    q.ny = q.ny + (cy * u) * p.delta;
    q.nz = q.nz + (cz * u) * p.delta;
    q.l = q.l + q.pl * p.delta;
    q.pl = q.pl + dpl * p.delta;
}

// The end state in the reference's variables (what finish_ray reads).
template <class Shape>
__device__ __forceinline__ void to_spherical(const FrameParams& p, const RayCart& q, Ray& o) {
    const double nn = norm3(q.nx, q.ny, q.nz);
    const double hx = q.nx / nn, hy = q.ny / nn, hz = q.nz / nn;
    double th = acos(hz), ph = atan2(hy, hx);
    normalize_theta_phi(th, ph);
    const double s = sqrt(hx * hx + hy * hy);
    double r, r2, rp;
    Shape::eval(p, q.l, r, r2, rp);
    // tangential velocity t = (J x n^) / r; p_theta = r (t . e_theta), e_theta = (hz hx / s, hz hy / s, -s)
    const double cx = q.jy * hz - q.jz * hy, cy = q.jz * hx - q.jx * hz, cz = q.jx * hy - q.jy * hx;
    o.l = q.l; o.th = th; o.ph = ph;
    o.pl = q.pl;
    o.pth = (cx * (hz * hx / s) + cy * (hz * hy / s)) + cz * (-s);
    o.pph = q.jz;                          // p_phi is the z component of the angular momentum
    o.pph2 = o.pph * o.pph;
}

template <class Shape>
__global__ void __launch_bounds__(kBlockCart) render_rows_f64_cart(const __grid_constant__ FrameParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const double R = p.max_radius;
    const unsigned gate = (R >= 0.0) ? abs_hi(R) : 0u;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    RayCart q;
    int state = 0;
    bool drained = false;
    uint32_t remaining = 0;
    unsigned long long ray = 0;
    RayTally tally;

    for (;;) {
        if (state == 2) {
            const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);
            Ray e;
            to_spherical<Shape>(p, q, e);
            const RayDiag nodiag = {qnan, qnan};
            finish_ray<Shape, TrigFast, false>(p, e, side, p.max_iterations - remaining, ray, tally, nodiag, 0.0);
            state = 0;
        }
        const unsigned idle = __ballot_sync(kFullCart, state == 0);
        if (idle) {
            if (!drained) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(&p.counters->next_ray, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFullCart, base, leader);
                if (state == 0) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < launch_rays) {
                        ray = idx;
                        new_photon_cart_for_ray(p, idx, tile_rays, q);
                        remaining = p.max_iterations;
                        state = (remaining == 0) ? 2 : 1;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFullCart, state != 0) == 0u) break;
        }
#pragma unroll 1
        for (uint32_t k = 0; k < p.window; ++k) {
            if (state == 1) {
                euler_step_cart<Shape>(p, q);
                --remaining;
                bool done = (remaining == 0);
                if (abs_hi(q.l) >= gate) {
                    done = done || (q.l > R) || (q.l < -R);
                    if (q.l != q.l) { remaining = 0; done = true; }
                }
                if (done) state = 2;
            }
        }
    }
    flush_tally(p, tally, lane);
}

// ---------------------------------------------------------------- the same scheme regrouped (CURVIS_PRECISION_F64_FAST)
// The chart-free step has no trigonometry and one divisor, r^2, so regrouped like render_f64_fast.cu it is the cheapest
// integrator of the library: with the momenta pre-scaled by the step (P_l = delta p_l, Jd = delta J, K = |Jd|^2)
//     u = 1/r^2 (MUFU seed + one cubic Newton step);   n += (Jd x n) u;   l += P_l;   P_l += K r'/r^3
// — 17 fp64-pipe instructions for Ellis (r'/r^3 = l u^2, no square root), 29 for Interstellar (1/r and r' from the per-metric
// table, u = Y^2, r'/r^3 = G Y^3) against 33.5 / 44 for the regrouped (theta, phi) step — one exit branch per step (step
// budget OR |l| at the radius gate), no windows to re-derive anything.  Every operation <= 1 ulp; the state agrees with the
// operation-for-operation chart-free kernel (and its oracle) to ~1e-13, the integers wherever that does not cross a
// decision boundary (tests/test_gpu_extensions.py states the bar: there is no guard band here — an extension's extension).
struct CartEllis {
    using Shape64 = ShapeEllis;
    struct Cache { __device__ __forceinline__ void reset(const double2*) {} };
    static __device__ __forceinline__ void factors(const FrameParams& p, double l, Cache&, double& u, double& f) {
        u = rcp_1ulp(fma(l, l, p.d_rho2));
        f = l * (u * u);                         // r'/r^3 = l / r^4
    }
};
struct CartInterstellar {
    using Shape64 = ShapeInterstellar;
    using Cache = InverseShapeCache;    // the table's coefficients stay in registers while the photon stays in one interval
    static __device__ __forceinline__ void factors(const FrameParams& p, double l, Cache& cache, double& u, double& f) {
        double H;
        interstellar_inverse_lookup(p.a, l, cache, u, H);
        f = copysign(H, l);
    }
};

template <class Fast>
__global__ void __launch_bounds__(kBlockCart, 5) render_rows_cart_fast(const __grid_constant__ FrameParams p) {
    using Shape = typename Fast::Shape64;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);
    const double R = p.max_radius;
    const double R_gate = fmin(R, p.fast_l_limit);            // (the end of the Interstellar table, else +inf)
    const unsigned gate = (R_gate >= 0.0) ? abs_hi(R_gate) : 0u;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    RayCart q;                 // here: pl = delta p_l, (jx, jy, jz) = delta J, l2 = |delta J|^2
    int state = 0;
    bool drained = false;
    uint32_t remaining = 0;
    unsigned long long ray = 0;
    RayTally tally;

    for (;;) {
        if (state == 2) {
            const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);
            RayCart o = q;                                     // back to the reference's units for the epilogue
            if (remaining != p.max_iterations) {
                o.pl = q.pl / p.delta; o.jx = q.jx / p.delta; o.jy = q.jy / p.delta; o.jz = q.jz / p.delta;
            } else {
                new_photon_cart_for_ray(p, ray, tile_rays, o); // a ray that never moved: regenerated whole (bit-exact)
            }
            Ray e;
            to_spherical<Shape>(p, o, e);
            const RayDiag nodiag = {qnan, qnan};
            finish_ray<Shape, TrigFast, false>(p, e, side, p.max_iterations - remaining, ray, tally, nodiag, 0.0);
            state = 0;
        }
        const unsigned idle = __ballot_sync(kFullCart, state == 0);
        if (idle) {
            if (!drained) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(&p.counters->next_ray, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFullCart, base, leader);
                if (state == 0) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < launch_rays) {
                        ray = idx;
                        new_photon_cart_for_ray(p, idx, tile_rays, q);
                        q.pl = q.pl * p.delta; q.jx = q.jx * p.delta; q.jy = q.jy * p.delta; q.jz = q.jz * p.delta;
                        q.l2 = fma(q.jx, q.jx, fma(q.jy, q.jy, q.jz * q.jz));
                        remaining = p.max_iterations;
                        state = (remaining == 0) ? 2 : 1;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFullCart, state != 0) == 0u) break;
        }
        if (state == 1) {
            uint32_t left = min(p.window, remaining);
            const uint32_t n = left;
            // (a ray already past the Interstellar table skips the regrouped steps: handled below)
            const bool near = (abs_hi(q.l) >= gate) && !(fabs(q.l) < p.fast_l_limit);
            if (!near) {
                typename Fast::Cache cache;
                cache.reset(p.inv_tab);
#pragma unroll 1
                do {
                    double u, f;
                    Fast::factors(p, q.l, cache, u, f);
                    const double cx = fma(q.jy, q.nz, -(q.jz * q.ny));
                    const double cy = fma(q.jz, q.nx, -(q.jx * q.nz));
                    const double cz = fma(q.jx, q.ny, -(q.jy * q.nx));
                    q.nx = fma(cx, u, q.nx);
                    q.ny = fma(cy, u, q.ny);
                    q.nz = fma(cz, u, q.nz);
                    q.l = q.l + q.pl;
                    q.pl = fma(q.l2, f, q.pl);
                    --left;
                    asm("" : "+r"(left));   // one induction variable (see render_f64_fast.cu)
                } while ((left != 0u) & (abs_hi(q.l) < gate));
            }
            remaining -= n - left;
            bool done = (remaining == 0);
            if (abs_hi(q.l) >= gate) {
                done = done || (q.l > R) || (q.l < -R);
                if (q.l != q.l) { remaining = 0; done = true; }
                // inside R but past the Interstellar table: the rest of the ray with the operation-for-operation step
                if (!done && !(fabs(q.l) < p.fast_l_limit)) {
                    RayCart o = q;
                    o.pl = q.pl / p.delta; o.jx = q.jx / p.delta; o.jy = q.jy / p.delta; o.jz = q.jz / p.delta;
                    o.l2 = (o.jx * o.jx + o.jy * o.jy) + o.jz * o.jz;
                    while (remaining) {
                        euler_step_cart<Shape>(p, o);
                        --remaining;
                        if ((o.l > R) || (o.l < -R)) break;
                        if (o.l != o.l) { remaining = 0; break; }
                    }
                    q.l = o.l; q.nx = o.nx; q.ny = o.ny; q.nz = o.nz; q.pl = o.pl * p.delta;
                    done = true;
                }
            }
            if (done) state = 2;
        }
        __syncwarp();
    }
    flush_tally(p, tally, lane);
}

template <class Fast>
cudaError_t launch_cart_fast(const FrameParams& p, int sm_count, int blocks_per_sm_override, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;
    if (blocks_per_sm_auto == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_auto, render_rows_cart_fast<Fast>, kBlockCart, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm_auto < 1) blocks_per_sm_auto = 1;
    }
    int blocks_per_sm = blocks_per_sm_auto;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < blocks_per_sm) blocks_per_sm = blocks_per_sm_override;
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    unsigned long long want = (rays + kBlockCart - 1) / kBlockCart;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    render_rows_cart_fast<Fast><<<grid, kBlockCart, 0, stream>>>(p);
    return cudaGetLastError();
}

template <class Shape>
cudaError_t launch_cart(const FrameParams& p, int sm_count, int blocks_per_sm_override, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;
    if (blocks_per_sm_auto == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_auto, render_rows_f64_cart<Shape>, kBlockCart, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm_auto < 1) blocks_per_sm_auto = 1;
    }
    int blocks_per_sm = blocks_per_sm_auto;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < blocks_per_sm) blocks_per_sm = blocks_per_sm_override;
    const unsigned long long rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    unsigned long long want = (rays + kBlockCart - 1) / kBlockCart;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    render_rows_f64_cart<Shape><<<grid, kBlockCart, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_render_cart_fast(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    const double ad = p.delta < 0.0 ? -p.delta : p.delta;
    if (!(ad >= 0x1p-100 && ad <= 0x1p100)) return launch_render_cart(p, metric_kind, t, sm_count, stream);   // the scaling needs an ordinary step
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch_cart_fast<CartEllis>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch_cart_fast<CartInterstellar>(p, sm_count, t.blocks_per_sm, stream);
    default: return launch_render_cart(p, metric_kind, t, sm_count, stream);        // Flat: r = l changes sign; the plain kernel
    }
}

cudaError_t launch_render_cart(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch_cart<ShapeEllis>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch_cart<ShapeInterstellar>(p, sm_count, t.blocks_per_sm, stream);
    case CURVIS_METRIC_FLAT: return launch_cart<ShapeFlat>(p, sm_count, t.blocks_per_sm, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace curvis
