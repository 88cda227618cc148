// geodesic_f64.cuh — device-side fp64 restatement of the per-ray work of
// RelativisticSystem::render_image (reference src/systems.rs:307-330), in the reference's
// operation order.  This translation unit MUST be compiled with -fmad=false: Rust never
// contracts a*b+c, so a contracted FMA here would change the low bit of the Euler update and
// break bit parity with the reference.  IEEE double +,-,*,/ and sqrt are correctly rounded on
// sm_100a, so with contraction off every line below is bit-identical to the CPU evaluation;
// the only operations that can differ from the reference's platform libm are the
// transcendentals, isolated behind the Trig / shape-function policies.
#pragma once
#include <cuda_runtime.h>
#include "frame_params.h"
#include "ieee_f64.cuh"
#include "trig_f64.cuh"
#include "shape_table.h"

namespace curvis {

#define CURVIS_PI 3.14159265358979323846264338327950288  // std::f64::consts::PI

// ---------------------------------------------------------------- photon state
// Position (t, l, theta, phi) + covariant momentum (p_t, p_l, p_theta, p_phi)
// (vectors.rs:135-139).  t is never read, p_t and p_phi never change (their derivatives are
// the literals 0.0, metrics.rs:259-264; x + 0.0*delta == x exactly), so a ray is five live
// doubles plus two per-ray constants.
struct Ray {
    double l, th, ph;    // x(1), x(2), x(3)
    double pl, pth;      // p(1), p(2)
    double pph, pph2;    // p(3) and p(3).powi(2) (same operands every step -> hoisted bit-safely)
};

// ---------------------------------------------------------------- trig policies
struct TrigCuda {  // CUDA math library (<= 2 ulp; not bit-identical to glibc)
    static __device__ __forceinline__ void sincos(double x, double& s, double& c) { ::sincos(x, &s, &c); }
    static __device__ __forceinline__ double sin(double x) { return ::sin(x); }
};

// ---------------------------------------------------------------- shape functions r(l)
// Each returns r(l), r_squared(l), r_derivative(l) exactly as the reference's three trait
// methods would (the reference re-evaluates them several times per step with identical
// arguments; evaluating once is bit-safe).
struct ShapeEllis {  // metrics.rs:417-421
    static constexpr int kind = CURVIS_METRIC_ELLIS;
    static __device__ __forceinline__ void eval(const FrameParams& p, double l, double& r, double& r2, double& rp) {
        r2 = p.rho * p.rho + l * l;  // :419 (and the radicand of :418)
        r = sqrt(r2);                // :418
        rp = l / r;                  // :420
    }
    // Same values through the unguarded correctly-rounded sequences (ieee_f64.cuh); only
    // called when step_operands_safe() holds.
    static __device__ __forceinline__ void eval_fast(const FrameParams& p, double l, double& r, double& r2, double& rp) {
        r2 = p.rho * p.rho + l * l;
        r = sqrt_rn_unguarded(r2);
        rp = div_rn_unguarded(l, r);
    }
    // The same values with the by-products the shared-reciprocal step needs: yr ~ 1/r (<= 2 ulp), from the square root's
    // own Newton iteration; r' = l / r through one correction step on yr.
    static __device__ __forceinline__ void eval_shared(const FrameParams& p, double l, double& r, double& r2, double& rp, double& yr, double half = 0.5) {
        r2 = p.rho * p.rho + l * l;
        r = sqrt_rn_with_rsqrt(r2, yr, half);
        rp = div_corrected(l, r, yr);
    }
    static __device__ __forceinline__ bool params_safe(const FrameParams& p) { return exponent_in(p.rho, -100, 100); }
    static constexpr bool kAheadBranchless = false, kAheadPinHalf = false;   // code-generation choices of the latency-form step loop (look_ahead)
};

struct ShapeInterstellar {  // metrics.rs:461-485
    static constexpr int kind = CURVIS_METRIC_INTERSTELLAR;
    static __device__ __forceinline__ void eval(const FrameParams& p, double l, double& r, double& r2, double& rp) {
        const double al = fabs(l);
        if (al > p.a) {
            const double x = 2.0 * (al - p.a) / (CURVIS_PI * p.m);            // :461
            const double at = atan(x);
            r = p.rho + p.m * (x * at - log(1.0 + x * x) / 2.0);               // :467-468
            const double sg = (l != l) ? l : copysign(1.0, l);                 // f64::signum
            rp = (2.0 / CURVIS_PI) * sg * at;                                  // :479-480
        } else {
            r = p.rho;   // :470
            rp = 0.0;    // :482
        }
        r2 = r * r;      // :474
    }
    // The lean kernels' evaluation (kernel_variant >= 1; operands inside the safe window, params_safe() below): the same
    // operations in the same order, with
    //   * x = 2 (|l| - a) / (pi m) as ONE correctly rounded quotient in three instructions: the divisor is a launch constant, the
    //     host supplies its correctly rounded reciprocal y, and q0 = a y; rem = fma(-b, q0, a); q = fma(rem, y, q0) is RN(a / b)
    //     (the value before the last rounding is within 2^-106 of the quotient, closer than a quotient of two doubles comes to
    //     a rounding boundary);
    //   * atan x and ln(1 + x^2) from tables (shape_table.h: <= 1 ulp of the exact values, like the CUDA library's — whose two
    //     calls were two thirds of this metric's step).  1 + x^2 is formed as the reference forms it (two roundings) and THEN
    //     looked up.  Outside the tables (x < 2^-10: the first 1.6e-4 m beyond the plateau; x >= 2^16) the library is called.
    static __device__ __forceinline__ void atan_log(const FrameParams& p, double x, double& at, double& lg) {
        const unsigned hx = (unsigned)__double2hiint(x);
        const unsigned ix = (hx >> kShapeTabShift) - kShapeTabBase;
        const double y = 1.0 + x * x;
        if (ix < (unsigned)kAtanTabIntervals) {
            const double cx = __hiloint2double((int)((hx & ~((1u << kShapeTabShift) - 1u)) | (1u << (kShapeTabShift - 1))), 0);
            const double tx = x - cx;
            const double2* a = reinterpret_cast<const double2*>(p.atan_tab) + ix * 3u;
            const double2 a01 = __ldg(a), a23 = __ldg(a + 1), a45 = __ldg(a + 2);
            at = fma(tx, fma(tx, fma(tx, fma(tx, fma(tx, a45.y, a45.x), a23.y), a23.x), a01.y), a01.x);
            const unsigned hy = (unsigned)__double2hiint(y);              // 1 <= y < 2^33
            const unsigned iy = (hy >> kShapeTabShift) - kLogTabBase;
            const double cy = __hiloint2double((int)((hy & ~((1u << kShapeTabShift) - 1u)) | (1u << (kShapeTabShift - 1))), 0);
            const double ty = y - cy;
            const double2* b = reinterpret_cast<const double2*>(p.log_tab) + iy * 3u;
            const double2 b01 = __ldg(b), b23 = __ldg(b + 1), b45 = __ldg(b + 2);
            lg = fma(ty, fma(ty, fma(ty, fma(ty, fma(ty, b45.y, b45.x), b23.y), b23.x), b01.y), b01.x);
        } else {
            at = atan(x);
            lg = log(y);
        }
    }
    static __device__ __forceinline__ void eval_fast(const FrameParams& p, double l, double& r, double& r2, double& rp) {
        const double al = fabs(l);
        if (al > p.a) {
            const double x = div_corrected(2.0 * (al - p.a), p.d_pim, p.d_pim_rcp);   // :461
            double at, lg;
            atan_log(p, x, at, lg);
            r = p.rho + p.m * (x * at - lg / 2.0);                             // :467-468
            const double sg = (l != l) ? l : copysign(1.0, l);                 // f64::signum
            rp = (2.0 / CURVIS_PI) * sg * at;                                  // :479-480
        } else {
            r = p.rho;   // :470
            rp = 0.0;    // :482
        }
        r2 = r * r;      // :474
    }
    static __device__ __forceinline__ void eval_shared(const FrameParams& p, double l, double& r, double& r2, double& rp, double& yr, double = 0.5) {
        eval_fast(p, l, r, r2, rp);
        yr = rcp_approx(r);       // r >= rho > 0
    }
    static __device__ __forceinline__ bool params_safe(const FrameParams& p) {
        return exponent_in(p.rho, -100, 100) && exponent_in(p.m, -100, 100) && exponent_in(p.a, -100, 100);
    }
    static constexpr bool kAheadBranchless = true, kAheadPinHalf = true;
};

struct ShapeFlat {  // metrics.rs:501-505
    static constexpr int kind = CURVIS_METRIC_FLAT;
    static __device__ __forceinline__ void eval(const FrameParams&, double l, double& r, double& r2, double& rp) {
        r = l; r2 = l * l; rp = 1.0;
    }
    static __device__ __forceinline__ void eval_fast(const FrameParams& p, double l, double& r, double& r2, double& rp) {
        eval(p, l, r, r2, rp);
    }
    static __device__ __forceinline__ void eval_shared(const FrameParams& p, double l, double& r, double& r2, double& rp, double& yr, double = 0.5) {
        eval(p, l, r, r2, rp);
        yr = rcp_approx(l);       // either sign; |l| is inside the safe window
    }
    static __device__ __forceinline__ bool params_safe(const FrameParams&) { return true; }
    static constexpr bool kAheadBranchless = false, kAheadPinHalf = false;
};

// ---------------------------------------------------------------- nalgebra-order helpers
__device__ __forceinline__ double norm3(double x, double y, double z) { return sqrt((x * x + y * y) + z * z); }

__device__ __forceinline__ void mat3_mul(const double* m, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = (m[0] * x + m[1] * y) + m[2] * z;
    oy = (m[3] * x + m[4] * y) + m[5] * z;
    oz = (m[6] * x + m[7] * y) + m[8] * z;
}

// f64::rem_euclid
__device__ __forceinline__ double rem_euclid(double x, double rhs) {
    const double r = fmod(x, rhs);
    return (r < 0.0) ? r + fabs(rhs) : r;
}

// normalize_theta_phi, algebra.rs:106-116
__device__ __forceinline__ void normalize_theta_phi(double& th, double& ph) {
    if (th < 0.0) { th = fabs(th); ph = ph + CURVIS_PI; }
    ph = rem_euclid(ph, 2.0 * CURVIS_PI);
}

// ---------------------------------------------------------------- ray generation
// camera_pixels_x_y_to_photon (systems.rs:531-534): outward_vector_on_camera_space
// (cameras.rs:150-164), camera_to_world rotation (:169-172), new_photon (metrics.rs:301-334).
// Metric::new_photon (metrics.rs:301-334) at the camera position for a tangent-space direction.
__device__ __forceinline__ void new_photon_from_direction(const CameraBlock& cam, double dx, double dy, double dz, Ray& q) {
    const double n = norm3(dx, dy, dz);
    dx = dx / n; dy = dy / n; dz = dz / n;                       // metrics.rs:320
    q.l = cam.cam_pos[1]; q.th = cam.cam_pos[2]; q.ph = cam.cam_pos[3];
    q.pl = dx;                                                   // :328
    q.pth = dy * cam.cam_r;                                      // :329  direction[1] * r(l)
    q.pph = dz * cam.cam_r * cam.cam_sin_theta;                  // :330  direction[2] * r(l) * sin(theta)
    q.pph2 = q.pph * q.pph;
}

// Camera::outward_vector_on_world_space_from_x_y (cameras.rs:150-172): unit vector in camera
// space, rotated to the tangent space at the camera (NOT re-normalised).
__device__ __forceinline__ void outward_vector_on_world_space(const CameraBlock& cam, uint32_t width, uint32_t height,
                                                              uint32_t px, uint32_t py, double& dx, double& dy, double& dz) {
    const double res_x = (double)width, res_y = (double)height;
    const double h = 0.5 - ((double)py / res_y);
    const double w = ((double)px / res_x) - 0.5;
    double vx = cam.focal_length * 1.0;
    double vy = -cam.sensor_width * w;
    double vz = cam.sensor_height * h;
    const double n = norm3(vx, vy, vz);
    vx = vx / n; vy = vy / n; vz = vz / n;                       // cameras.rs:163
    mat3_mul(cam.cam_to_world, vx, vy, vz, dx, dy, dz);          // cameras.rs:171
}

__device__ __forceinline__ void new_photon_from_camera(const CameraBlock& cam, uint32_t width, uint32_t height,
                                                       uint32_t px, uint32_t py, Ray& q) {
    double dx, dy, dz;
    outward_vector_on_world_space(cam, width, height, px, py, dx, dy, dz);
    new_photon_from_direction(cam, dx, dy, dz, q);
}

// Ray `idx` of a launch -> (frame, pixel column, row of the tile).  idx < 2^53 and the divisors are launch constants, so
// the quotients come from one double multiplication by the host-computed reciprocal plus a one-step correction — the 64-bit
// integer divisions this replaces are ~150 instructions each, and a warp pays them once per refill, not once per ray.
__device__ __forceinline__ void split_ray_index(const FrameParams& p, unsigned long long idx, unsigned long long tile_rays,
                                                unsigned long long& frame, uint32_t& px, uint32_t& k) {
    unsigned long long r = idx;
    frame = 0;
    if (p.n_frames > 1) {
        frame = (unsigned long long)((double)idx * p.inv_tile_rays);
        if (frame * tile_rays > idx) --frame;
        else if ((frame + 1) * tile_rays <= idx) ++frame;
        r = idx - frame * tile_rays;
    }
    k = (uint32_t)((double)r * p.inv_width);
    if ((unsigned long long)k * p.width > r) --k;
    else if ((unsigned long long)(k + 1) * p.width <= r) ++k;
    px = (uint32_t)(r - (unsigned long long)k * p.width);
}

// Row k of the tile, position i in it -> the pixel (px, py) of the frame.  Whole rows: (i, row_begin + k row_stride); block
// tiles (frame_params.h): block v = row_begin + k row_stride is block v % blocks_per_row of frame row v / blocks_per_row.
__device__ __forceinline__ void tile_pixel(const FrameParams& p, uint32_t k, uint32_t i, uint32_t& px, uint32_t& py) {
    const uint32_t v = p.row_begin + k * p.row_stride;
    if (p.blocks_per_row <= 1u) { px = i; py = v; return; }
    uint32_t y = (uint32_t)((double)v * p.inv_blocks_per_row);
    if (y * p.blocks_per_row > v) --y;
    else if ((y + 1u) * p.blocks_per_row <= v) ++y;
    py = y;
    px = (v - y * p.blocks_per_row) * p.width + i;
}

// The camera of frame `frame` of a launch (batched launches hold one per frame).
__device__ __forceinline__ const CameraBlock& camera_of_frame(const FrameParams& p, unsigned long long frame) {
    return (p.ray_dirs || p.n_frames <= 1) ? p.cam : p.cameras[frame];
}

// Will ray `idx` be one of the launch's stragglers?  A scheduling hint (render_f64_fast.cu claims such rays first), computed
// from the pixel's (unnormalised) tangent direction d alone — no normalisation, no square root, no division, never used for a
// result.  The 10^4-step rays of a frame are the near-critical photons (impact parameter b = r_cam sin(angle to the radial
// direction) close to the throat radius rho: they wind around the throat) whose orbit plane almost contains the polar axis
// (sin theta_min = |p_phi| / L small, with L^2 = p_theta^2 + p_phi^2 / sin^2 theta_0, p_theta ~ d_y, p_phi ~ d_z sin theta_0):
// every half turn passes a coordinate pole, and explicit Euler in (theta, phi) kicks them there.  On the 4K default frames
// (oracle records, rows 1060-1100): every ray of more than 6000 steps — 92 Ellis, 176 Interstellar, all within 9 rows of the
// central one — has sin theta_min < 0.03 and 0.7 rho < b < 2.1 rho; the predicate lists 26,876 rays (0.3 % of the frame).
__device__ __forceinline__ bool ray_predicted_long(const FrameParams& p, unsigned long long idx, unsigned long long tile_rays, double limit2) {
    double dx, dy, dz, s0, r0;
    if (p.ray_dirs) {
        dx = p.ray_dirs[3 * idx]; dy = p.ray_dirs[3 * idx + 1]; dz = p.ray_dirs[3 * idx + 2];
        s0 = p.cam.cam_sin_theta; r0 = p.cam.cam_r;
    } else {
        unsigned long long f; uint32_t px, k;
        split_ray_index(p, idx, tile_rays, f, px, k);
        const CameraBlock& cam = camera_of_frame(p, f);
        uint32_t py;
        tile_pixel(p, k, px, px, py);
        const double vx = cam.focal_length;
        const double vy = -cam.sensor_width * (((double)px * p.inv_frame_width) - 0.5);
        const double vz = cam.sensor_height * (0.5 - ((double)py / (double)p.height));
        dx = (cam.cam_to_world[0] * vx + cam.cam_to_world[1] * vy) + cam.cam_to_world[2] * vz;
        dy = (cam.cam_to_world[3] * vx + cam.cam_to_world[4] * vy) + cam.cam_to_world[5] * vz;
        dz = (cam.cam_to_world[6] * vx + cam.cam_to_world[7] * vy) + cam.cam_to_world[8] * vz;
        s0 = cam.cam_sin_theta; r0 = cam.cam_r;
    }
    const double t2 = dy * dy + dz * dz;                       // |tangential part|^2
    const double b2 = (r0 * r0) * t2;                          // b^2 |d|^2
    const double c2 = (p.rho * p.rho) * (dx * dx + t2);        // rho^2 |d|^2
    return ((dz * dz) * (s0 * s0) < limit2 * t2) && (b2 > 0.5 * c2) && (b2 < 4.5 * c2);      // (NaN: false)
}

// Ray `idx` of a launch: frame = idx / tile_rays (batched launches), pixel = idx % tile_rays.
__device__ __forceinline__ void new_photon_for_ray(const FrameParams& p, unsigned long long idx, unsigned long long tile_rays, Ray& q) {
    if (p.ray_dirs) {
        new_photon_from_direction(p.cam, p.ray_dirs[3 * idx], p.ray_dirs[3 * idx + 1], p.ray_dirs[3 * idx + 2], q);
        return;
    }
    unsigned long long f; uint32_t px, k;
    split_ray_index(p, idx, tile_rays, f, px, k);
    uint32_t py;
    tile_pixel(p, k, px, px, py);
    new_photon_from_camera(camera_of_frame(p, f), p.frame_width, p.height, px, py, q);
}

// ---------------------------------------------------------------- one explicit Euler step
// update_relativistic_object (metrics.rs:283-297) = object_position_diff_contr (:223-244) +
// object_momentum_diff_cov (:247-270), evaluated at the old state, then x += dx*delta,
// p += dp*delta.
template <class Shape, class Trig>
__device__ __forceinline__ void euler_step(const FrameParams& p, Ray& q) {
    double s, c;
    Trig::sincos(q.th, s, c);
    double r, r2, rp;
    Shape::eval(p, q.l, r, r2, rp);
    const double s2 = s * s;                                  // sin().powi(2)
    const double g22c = 1.0 / r2;                             // :90 over :61-63
    const double g33c = 1.0 / (r2 * s2);                      // :93 over :66-68
    const double dl = q.pl;                                   // :238  p_l * 1
    const double dth = q.pth * g22c;                          // :239
    const double dph = q.pph * g33c;                          // :240
    const double b2 = q.pth * q.pth + q.pph2 / s2;            // :257
    const double dpl = (b2 * rp) / ((r * r) * r);             // :261
    const double dpth = q.pph2 * (c / (r2 * (s2 * s)));       // :262
    q.l = q.l + dl * p.delta;                                 // :295
    q.th = q.th + dth * p.delta;
    q.ph = q.ph + dph * p.delta;
    q.pl = q.pl + dpl * p.delta;                              // :296
    q.pth = q.pth + dpth * p.delta;
}

// ---------------------------------------------------------------- the same step, tuned
// Identical arithmetic (every rounding of the reference sequence is kept; IEEE division and
// sqrt are correctly rounded either way), but the six divisions and the square root run the
// unguarded Newton-Raphson sequences of ieee_f64.cuh behind ONE merged operand-range check
// instead of one guard + branch each.  `ray_safe` carries the per-ray part of the check
// (p_phi^2 and the frame parameters, constant along a ray).  Outside the window — rays grazing
// the coordinate poles (sin theta -> 0), NaN/Inf states — the plain operators are used.
// p_phi^2 == 0 exactly (the rays of the image's central row, fired along a meridian) is safe as well: every quotient it
// enters is an exact zero in the unguarded sequences as in the plain operators.
__device__ __forceinline__ bool ray_operands_safe(const Ray& q) { return exponent_in(q.pph2, -200, 200) || q.pph2 == 0.0; }

template <class Shape, class Trig>
__device__ __forceinline__ void euler_step_tuned(const FrameParams& p, Ray& q, bool ray_safe) {
    double s, c;
    Trig::sincos(q.th, s, c);
    double dl, dth, dph, dpl, dpth;
    const bool safe = ray_safe && exponent_in(s, -60, 1) && exponent_in(q.l, -100, 100) && (fabs(q.pth) < 0x1p100);
    if (safe) {
        double r, r2, rp;
        Shape::eval_fast(p, q.l, r, r2, rp);
        const double s2 = s * s;
        const double g22c = rcp_rn_unguarded(r2);
        const double g33c = rcp_rn_unguarded(r2 * s2);
        dl = q.pl;
        dth = q.pth * g22c;
        dph = q.pph * g33c;
        const double b2 = q.pth * q.pth + div_rn_unguarded(q.pph2, s2);
        dpl = div_rn_unguarded(b2 * rp, (r * r) * r);
        dpth = q.pph2 * div_rn_unguarded(c, r2 * (s2 * s));
    } else {
        double r, r2, rp;
        Shape::eval(p, q.l, r, r2, rp);
        const double s2 = s * s;
        const double g22c = 1.0 / r2;
        const double g33c = 1.0 / (r2 * s2);
        dl = q.pl;
        dth = q.pth * g22c;
        dph = q.pph * g33c;
        const double b2 = q.pth * q.pth + q.pph2 / s2;
        dpl = (b2 * rp) / ((r * r) * r);
        dpth = q.pph2 * (c / (r2 * (s2 * s)));
    }
    q.l = q.l + dl * p.delta;
    q.th = q.th + dth * p.delta;
    q.ph = q.ph + dph * p.delta;
    q.pl = q.pl + dpl * p.delta;
    q.pth = q.pth + dpth * p.delta;
}

// ---------------------------------------------------------------- the lean step (default kernel)
// euler_step_tuned with the operand check done entirely on the integer pipe (high-word
// exponent compares) BEFORE the trigonometry, so the common case runs sincos_fast + the
// unguarded sequences with two predictable branches and no fp64 compare.  On sm_100a an fp64
// instruction holds the scheduler's dispatch port for two cycles, so every fp64 op removed is
// worth two integer ops (profiles/r01_f64_v2_ncu_summary.txt).  Arithmetic is unchanged.
// Right-hand side of the geodesic equations at a state (metrics.rs:223-270), lean form.
// The shared-reciprocal right-hand side (kernel_variant 4) from its parts: the shape function at l (r, r^2, r', yr ~ 1/r) and
// (sin, cos) of theta.  The same seven roundings as the plain operators: the six divisors' reciprocals are built from TWO seeds —
// yr (a by-product of the square root) and ys ~ 1/sin^2 theta — and one correction step each (ieee_f64.cuh: div_corrected /
// rcp_corrected): 77 fp64-pipe instructions per step instead of 100.
__device__ __forceinline__ void rhs_shared_from(double r, double r2, double rp, double yr, double s, double c, double pth, double pph, double pph2,
                                                double& dth, double& dph, double& dpl, double& dpth) {
    const double s2 = s * s;
    const double yr2 = yr * yr;                                        // ~ 1/r^2
    const double ys = rcp_approx(s2);                                  // ~ 1/sin^2
    const double yrs = yr2 * ys;                                       // ~ 1/(r^2 sin^2)
    const double g22c = rcp_corrected(r2, yr2);                        // metrics.rs:90
    const double g33c = rcp_corrected(r2 * s2, yrs);                   // :93
    dth = pth * g22c;
    dph = pph * g33c;
    const double b2 = pth * pth + div_corrected(pph2, s2, ys);         // :257
    dpl = div_corrected(b2 * rp, (r * r) * r, yr2 * yr);               // :261
    dpth = pph2 * div_corrected(c, r2 * (s2 * s), yrs * (ys * s));     // :262
}

template <class Shape, bool SHARED = true>
__device__ __forceinline__ void rhs_lean(const FrameParams& p, bool ray_safe, double l, double th, double pth, double pph, double pph2,
                                         double& dth, double& dph, double& dpl, double& dpth, double& s, const TrigPins* pins = nullptr) {
    const bool pre = ray_safe && (abs_hi(th) < pow2_hi(30)) && ((abs_hi(l) - pow2_hi(-100)) < (pow2_hi(100) - pow2_hi(-100))) &&
                     (abs_hi(pth) < pow2_hi(100));
    double c;
    if (pre) {
        if (pins) sincos_fast_pinned(*pins, th, s, c);     // (the render kernel's step: constants pinned in registers)
        else sincos_fast<true>(th, s, c);
    } else TrigFast::sincos(th, s, c);
    if (pre && abs_hi(s) >= pow2_hi(-60)) {
        if (SHARED) {
            // kernel_variant 4 (default): rhs_shared_from
            double r, r2, rp, yr;
            Shape::eval_shared(p, l, r, r2, rp, yr, pins ? pins->half : 0.5);
            rhs_shared_from(r, r2, rp, yr, s, c, pth, pph, pph2, dth, dph, dpl, dpth);
        } else {
            double r, r2, rp;
            Shape::eval_fast(p, l, r, r2, rp);
            const double s2 = s * s;
            const double g22c = rcp_rn_unguarded(r2);
            const double g33c = rcp_rn_unguarded(r2 * s2);
            dth = pth * g22c;
            dph = pph * g33c;
            const double b2 = pth * pth + div_rn_unguarded(pph2, s2);
            dpl = div_rn_unguarded(b2 * rp, (r * r) * r);
            dpth = pph2 * div_rn_unguarded(c, r2 * (s2 * s));
        }
    } else {
        double r, r2, rp;
        Shape::eval(p, l, r, r2, rp);
        const double s2 = s * s;
        const double g22c = 1.0 / r2;
        const double g33c = 1.0 / (r2 * s2);
        dth = pth * g22c;
        dph = pph * g33c;
        const double b2 = pth * pth + pph2 / s2;
        dpl = (b2 * rp) / ((r * r) * r);
        dpth = pph2 * (c / (r2 * (s2 * s)));
    }
}

// Trajectory diagnostics of curvis_ray_record (NaN when the kernel does not track them): min |sin theta| over the states
// the right-hand side was evaluated at, and the stiffness max (delta dphi/dlambda)^2 = max delta^2 p_phi^2 / (r^2 sin^2)^2.
struct RayDiag { double min_abs_sin, stiffness; };

template <class Shape, bool TRACK = false, bool SHARED = true>
__device__ __forceinline__ void euler_step_lean(const FrameParams& p, Ray& q, bool ray_safe, RayDiag* diag = nullptr, const TrigPins* pins = nullptr) {
    double dth, dph, dpl, dpth, s;
    rhs_lean<Shape, SHARED>(p, ray_safe, q.l, q.th, q.pth, q.pph, q.pph2, dth, dph, dpl, dpth, s, pins);
    if (TRACK) {
        const double step_phi = dph * p.delta;
        diag->min_abs_sin = fmin(diag->min_abs_sin, fabs(s));
        diag->stiffness = fmax(diag->stiffness, step_phi * step_phi);
    }
    q.l = q.l + q.pl * p.delta;                               // metrics.rs:295 (dl = p_l * 1)
    q.th = q.th + dth * p.delta;
    q.ph = q.ph + dph * p.delta;
    q.pl = q.pl + dpl * p.delta;                              // :296
    q.pth = q.pth + dpth * p.delta;
}

// ---- the same step in LATENCY form (list mode of render_rows_f64_lean: the guard band's re-integration launch).
// That launch holds a few hundred rays — at most one warp per scheduler — so its duration is ONE ray's dependent chain times
// its step count, not the fp64 pipe's throughput.  Inside one step the chain is  theta -> sincos (13 dependent operations) ->
// 1/sin^2 -> quotient -> p_theta  (28 operations, 530 cycles measured), but the NEXT step's two long evaluations depend on
// little of this one: its shape function needs only l + delta p_l (one operation into the step), its sincos only
// theta + delta p_theta / r^2 (six operations).  euler_steps_ahead() carries both across the loop's back edge — the step loop
// computes the right-hand side from the carried parts, updates the state and starts the next step's parts in the SAME basic
// block, so the two chains overlap with the quotients.  Same operations on the same values: nothing is rounded differently;
// the look-ahead of the last step is discarded.  States outside the unguarded sequences' window take rhs_lean's own path.
struct StepParts { double r, r2, rp, yr, s, c; };

template <class Shape>
__device__ __forceinline__ bool look_ahead(const FrameParams& p, const TrigPins& pins, bool ray_safe, const Ray& q, StepParts& a) {
    // rhs_lean's test.  Shape::kAheadBranchless: as one predicate chain instead of short-circuit branches — which form ptxas
    // schedules better is measured per metric (4K frames, whole-frame kernel: Interstellar 122.8 -> 119.2 ms, Ellis 83.3 -> 85.4)
    if (Shape::kAheadBranchless) {
        sincos_fast_pinned(pins, q.th, a.s, a.c);                   // (speculative: plain arithmetic on any operand)
        Shape::eval_shared(p, q.l, a.r, a.r2, a.rp, a.yr, pins.half);
        return ray_safe & (abs_hi(q.th) < pow2_hi(30)) & ((abs_hi(q.l) - pow2_hi(-100)) < (pow2_hi(100) - pow2_hi(-100))) &
               (abs_hi(q.pth) < pow2_hi(100)) & (abs_hi(a.s) >= pow2_hi(-60));
    }
    const bool pre = ray_safe && (abs_hi(q.th) < pow2_hi(30)) && ((abs_hi(q.l) - pow2_hi(-100)) < (pow2_hi(100) - pow2_hi(-100))) &&
                     (abs_hi(q.pth) < pow2_hi(100));
    sincos_fast_pinned(pins, q.th, a.s, a.c);
    Shape::eval_shared(p, q.l, a.r, a.r2, a.rp, a.yr, pins.half);
    return pre && abs_hi(a.s) >= pow2_hi(-60);
}

// Up to n Euler steps; stops after the step that takes |l| to the radius gate (`near`).  Returns the steps taken.
template <class Shape>
__device__ __forceinline__ uint32_t euler_steps_ahead(const FrameParams& p, const TrigPins& pins, bool ray_safe, Ray& q, uint32_t n, unsigned gate, bool& near) {
    uint32_t k = 0;
    StepParts a;
    bool ok = look_ahead<Shape>(p, pins, ray_safe, q, a);
    for (;;) {
        if (!ok) {                                                  // outside the window: the general step
            euler_step_lean<Shape, false, true>(p, q, ray_safe, nullptr, &pins);
            ++k;
            if (abs_hi(q.l) >= gate) { near = true; break; }
            if (k >= n) break;
            ok = look_ahead<Shape>(p, pins, ray_safe, q, a);
            continue;
        }
        bool at_gate;
        do {
            double dth, dph, dpl, dpth;
            rhs_shared_from(a.r, a.r2, a.rp, a.yr, a.s, a.c, q.pth, q.pph, q.pph2, dth, dph, dpl, dpth);
            q.l = q.l + q.pl * p.delta;                             // metrics.rs:295
            q.th = q.th + dth * p.delta;
            q.ph = q.ph + dph * p.delta;
            q.pl = q.pl + dpl * p.delta;                            // :296
            q.pth = q.pth + dpth * p.delta;
            ++k;
            ok = look_ahead<Shape>(p, pins, ray_safe, q, a);
            at_gate = abs_hi(q.l) >= gate;
        } while (ok & !at_gate & (k < n));
        if (at_gate) { near = true; break; }
        if (k >= n) break;
    }
    return k;
}

// CURVIS_INTEGRATOR_EULER_ADAPTIVE (extension; its oracle is oracle_step_adaptive, same operation order): the explicit
// Euler step above with the step size cut near a coordinate pole.  Monitor m = max(|delta dphi/dlambda|,
// |delta dtheta/dlambda| / |sin theta|) — the azimuth advance of a full step and the polar advance measured in units of
// the distance to the pole; m > step_tolerance: h = delta * (step_tolerance / m), else h = delta and the step is the
// reference's bit for bit.
template <class Shape, bool TRACK = false>
__device__ __forceinline__ void euler_step_adaptive(const FrameParams& p, Ray& q, bool ray_safe, RayDiag* diag = nullptr) {
    double dth, dph, dpl, dpth, s;
    rhs_lean<Shape>(p, ray_safe, q.l, q.th, q.pth, q.pph, q.pph2, dth, dph, dpl, dpth, s);
    const double step_phi = dph * p.delta;
    const double m = fmax(fabs(step_phi), fabs(dth * p.delta) / fabs(s));
    double h = p.delta;
    if (m > p.step_tolerance) h = p.delta * (p.step_tolerance / m);
    if (TRACK) {
        const double sp = dph * h;
        diag->min_abs_sin = fmin(diag->min_abs_sin, fabs(s));
        diag->stiffness = fmax(diag->stiffness, sp * sp);
    }
    q.l = q.l + q.pl * h;
    q.th = q.th + dth * h;
    q.ph = q.ph + dph * h;
    q.pl = q.pl + dpl * h;
    q.pth = q.pth + dpth * h;
}

// CURVIS_INTEGRATOR_RK4 (extension; the reference only has the Euler step above): classical
// Runge-Kutta on the same right-hand side.  Operation order (shared with the oracle's
// oracle_step_rk4 so the two agree bit for bit up to the transcendentals):
//   h2 = delta*0.5, d6 = delta/6;  y2 = y + h2*k1;  y3 = y + h2*k2;  y4 = y + delta*k3;
//   y += d6 * (((k1 + 2*k2) + 2*k3) + k4)
template <class Shape>
__device__ __forceinline__ void rk4_step_lean(const FrameParams& p, Ray& q, bool ray_safe) {
    const double h2 = p.delta * 0.5, d6 = p.delta / 6.0;
    double k1th, k1ph, k1pl, k1pth, k2th, k2ph, k2pl, k2pth, k3th, k3ph, k3pl, k3pth, k4th, k4ph, k4pl, k4pth, s;
    const double k1l = q.pl;
    rhs_lean<Shape>(p, ray_safe, q.l, q.th, q.pth, q.pph, q.pph2, k1th, k1ph, k1pl, k1pth, s);
    const double k2l = q.pl + h2 * k1pl;
    rhs_lean<Shape>(p, ray_safe, q.l + h2 * k1l, q.th + h2 * k1th, q.pth + h2 * k1pth, q.pph, q.pph2, k2th, k2ph, k2pl, k2pth, s);
    const double k3l = q.pl + h2 * k2pl;
    rhs_lean<Shape>(p, ray_safe, q.l + h2 * k2l, q.th + h2 * k2th, q.pth + h2 * k2pth, q.pph, q.pph2, k3th, k3ph, k3pl, k3pth, s);
    const double k4l = q.pl + p.delta * k3pl;
    rhs_lean<Shape>(p, ray_safe, q.l + p.delta * k3l, q.th + p.delta * k3th, q.pth + p.delta * k3pth, q.pph, q.pph2, k4th, k4ph, k4pl, k4pth, s);
    q.l = q.l + d6 * (((k1l + 2.0 * k2l) + 2.0 * k3l) + k4l);
    q.th = q.th + d6 * (((k1th + 2.0 * k2th) + 2.0 * k3th) + k4th);
    q.ph = q.ph + d6 * (((k1ph + 2.0 * k2ph) + 2.0 * k3ph) + k4ph);
    q.pl = q.pl + d6 * (((k1pl + 2.0 * k2pl) + 2.0 * k3pl) + k4pl);
    q.pth = q.pth + d6 * (((k1pth + 2.0 * k2pth) + 2.0 * k3pth) + k4pth);
}

// ---------------------------------------------------------------- direction -> texel
// Continuous equirectangular coordinates of a direction: theta_phi_of_image_from_vector3
// (images.rs:151-167) then the two expressions inside pixel_indexes_x_y_from_theta_phi_of_image
// (:115-121) before their truncating casts.
// `sin_img` (guard band of CURVIS_PRECISION_F64_FAST only): sin of the image polar angle, = d(direction)/d(phi_img).
__device__ __forceinline__ void image_coordinates(const Background& bg, double dx, double dy, double dz, double& fx, double& fy, double& sin_img) {
    double wx, wy, wz;
    mat3_mul(bg.inv_rot, dx, dy, dz, wx, wy, wz);             // images.rs:139-141
    const double rn = norm3(wx, wy, wz);
    double th = acos(wz / rn);                                // algebra.rs:130
    double ph = atan2(wy, wx);                                // :131
    normalize_theta_phi(th, ph);                              // :133
    normalize_theta_phi(th, ph);                              // images.rs:116
    fy = (th / CURVIS_PI) * (double)bg.height;                                   // :118
    fx = rem_euclid(0.5 - ph / (2.0 * CURVIS_PI), 1.0) * (double)bg.width;       // :119
    sin_img = sqrt(wx * wx + wy * wy) / rn;
}

__device__ __forceinline__ void image_coordinates(const Background& bg, double dx, double dy, double dz, double& fx, double& fy) {
    double unused;
    image_coordinates(bg, dx, dy, dz, fx, fy, unused);
}

// The truncating casts of images.rs:118-119 + the bounds the reference's get_pixel panics on.
__device__ __forceinline__ bool nearest_texel(const Background& bg, double fx, double fy, uint32_t& tx, uint32_t& ty) {
    ty = __double2uint_rz(fy);  // Rust `as u32`: truncate, saturate, NaN -> 0
    tx = __double2uint_rz(fx);
    bool clamped = false;
    if (tx >= bg.width) { tx = bg.width - 1; clamped = true; }
    if (ty >= bg.height) { ty = bg.height - 1; clamped = true; }
    return clamped;
}

// SphericalImage::get_pixel_from_vector3 (images.rs:171-174): texel index of a direction.
__device__ __forceinline__ bool texel_from_direction(const Background& bg, double dx, double dy, double dz,
                                                     uint32_t& tx, uint32_t& ty) {
    double fx, fy;
    image_coordinates(bg, dx, dy, dz, fx, fy);
    return nearest_texel(bg, fx, fy, tx, ty);
}

// ---------------------------------------------------------------- bilinear tap (extension)
// CURVIS_SAMPLING_BILINEAR — the reference only has the nearest lookup above.  Texel centres sit
// at integer + 0.5; wrap in x, clamp in y.  The integer parts are split off in fp64 (fp32 holds
// only ~10 fractional bits at x ~ 8192), the three lerps run in fp32 with fmaf on RGBA texels
// staged as float4 (one 128-bit load per tap).  tests/test_gpu_bilinear.py: bit-identical to the
// oracle's fp32 restatement when fed the same coordinates.
__device__ __forceinline__ float4 bilinear_tap(const Background& bg, double fx, double fy) {
    const double ux = fx - 0.5, uy = fy - 0.5;
    const double x0d = floor(ux), y0d = floor(uy);
    float wx = (float)(ux - x0d), wy = (float)(uy - y0d);
    long long x0 = (x0d == x0d) ? (long long)x0d : 0ll, y0 = (y0d == y0d) ? (long long)y0d : 0ll;   // NaN -> 0 like `as`
    if (!(wx == wx)) wx = 0.f;
    if (!(wy == wy)) wy = 0.f;
    const long long W = (long long)bg.width, H = (long long)bg.height;
    x0 = ((x0 % W) + W) % W;
    const long long x1 = (x0 + 1) % W;
    long long y1 = y0 + 1;
    y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > H - 1 ? H - 1 : y1);
    const float4 t00 = __ldg(bg.texels_f4 + y0 * W + x0), t10 = __ldg(bg.texels_f4 + y0 * W + x1);
    const float4 t01 = __ldg(bg.texels_f4 + y1 * W + x0), t11 = __ldg(bg.texels_f4 + y1 * W + x1);
    float4 o;
    const float tx_ = fmaf(wx, t10.x - t00.x, t00.x), bx = fmaf(wx, t11.x - t01.x, t01.x); o.x = fmaf(wy, bx - tx_, tx_);
    const float ty_ = fmaf(wx, t10.y - t00.y, t00.y), by = fmaf(wx, t11.y - t01.y, t01.y); o.y = fmaf(wy, by - ty_, ty_);
    const float tz_ = fmaf(wx, t10.z - t00.z, t00.z), bz = fmaf(wx, t11.z - t01.z, t01.z); o.z = fmaf(wy, bz - tz_, tz_);
    const float tw_ = fmaf(wx, t10.w - t00.w, t00.w), bw = fmaf(wx, t11.w - t01.w, t01.w); o.w = fmaf(wy, bw - tw_, tw_);
    return o;
}

__device__ __forceinline__ uint32_t quantize_channel(float v) {   // round to nearest even, clamp to u8; NaN -> 0
    v = rintf(v);
    return (v >= 255.f) ? 255u : ((v > 0.f) ? (uint32_t)v : 0u);
}

// ---------------------------------------------------------------- escaped photon -> lookup direction
// relativistic_vector_to_direction (metrics.rs:339-349) of the covariant momentum: CURVIS_FRAME_LOCAL and
// CURVIS_FRAME_WORLD_QUIRK scale the phi component by frame_field_22 like :347 does; CURVIS_FRAME_WORLD scales it by
// frame_field_33 (:122-124).  `s` returns sin(theta) of the final state.
template <class Shape, class Trig>
__device__ __forceinline__ void tangent_direction(const FrameParams& p, const Ray& q, double& dx, double& dy, double& dz, double& s) {
    s = Trig::sin(q.th);
    double r, r2, rp;
    Shape::eval(p, q.l, r, r2, rp);
    const double v2 = q.pth * (1.0 / r2);                                          // to_contravariant, :190-203
    const double v3 = q.pph * (1.0 / (r2 * (s * s)));
    dx = q.pl;                                                                     // :345  (g11 = frame_field_11 = 1)
    dy = v2 * r;                                                                   // :346
    dz = (p.frame == CURVIS_FRAME_WORLD) ? v3 * (r * s) : v3 * r;                  // :347 as written, or frame_field_33
}

// escaped_photon_to_world_direction (systems.rs:144-187): the tangent-frame direction rotated by
// rotation_from_two_vectors(x, vector3_from_theta_phi(theta, phi)) (algebra.rs:92-101, :118-126; nalgebra 0.33
// Rotation3::rotation_between = axis-angle about x^ x pos^ by acos(x^ . pos^), Rodrigues matrix).  Returns false where
// the reference panics ("v1 and v2 must not be parallel": the photon sits exactly on the +-x axis).  Out of line: it is
// evaluated once per ray and only with CURVIS_FRAME_WORLD*.
static __device__ __noinline__ bool rotate_tangent_to_world(double th, double ph, double& dx, double& dy, double& dz) {
    normalize_theta_phi(th, ph);                                                   // algebra.rs:120
    double st, ct, sp, cp;
    ::sincos(th, &st, &ct);
    ::sincos(ph, &sp, &cp);
    const double wx = st * cp, wy = st * sp, wz = ct;                              // :122-124
    // v1 = (1,0,0): cross = (0*wz - 0*wy, 0*wx - 1*wz, 1*wy - 0*wx)
    const double c0x = 0.0 * wz - 0.0 * wy, c0y = 0.0 * wx - 1.0 * wz, c0z = 1.0 * wy - 0.0 * wx;
    if (norm3(c0x, c0y, c0z) == 0.0) return false;                                 // algebra.rs:95-97
    const double n2 = norm3(wx, wy, wz);
    const double bx = wx / n2, by = wy / n2, bz = wz / n2;                         // try_normalize (n1 = 1: a^ = x)
    const double cx = 0.0 * bz - 0.0 * by, cy = 0.0 * bx - 1.0 * bz, cz = 1.0 * by - 0.0 * bx;
    const double cn = norm3(cx, cy, cz);
    const double dot = (1.0 * bx + 0.0 * by) + 0.0 * bz;
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (cn > 2.220446049250313e-16) {                                              // Unit::try_new(c, default_epsilon)
        const double ux = cx / cn, uy = cy / cn, uz = cz / cn;
        const double angle = acos(dot) * 1.0;
        if (angle != 0.0) {                                                        // Rotation3::from_axis_angle
            const double sqx = ux * ux, sqy = uy * uy, sqz = uz * uz;
            double sn, cs;
            ::sincos(angle, &sn, &cs);
            const double omc = 1.0 - cs;
            m[0] = sqx + (1.0 - sqx) * cs; m[1] = ux * uy * omc - uz * sn; m[2] = ux * uz * omc + uy * sn;
            m[3] = ux * uy * omc + uz * sn; m[4] = sqy + (1.0 - sqy) * cs; m[5] = uy * uz * omc - ux * sn;
            m[6] = ux * uz * omc - uy * sn; m[7] = uy * uz * omc + ux * sn; m[8] = sqz + (1.0 - sqz) * cs;
        }
    } else if (dot < 0.0) {
        return false;                                                              // rotation_between -> None -> unwrap panics
    }
    double ox, oy, oz;
    mat3_mul(m, dx, dy, dz, ox, oy, oz);                                           // systems.rs:183
    dx = ox; dy = oy; dz = oz;
    return true;
}

// ---------------------------------------------------------------- ray epilogue (all per-ray kernels)
// photon_escape_to_pixel + put_pixel (systems.rs:540-561, :324) for one finished ray, plus the
// optional outputs: fp32 RGBA (the unrounded tap) and the per-ray record.
struct RayTally { unsigned pos = 0, neg = 0, none = 0, clamped = 0; unsigned long long steps = 0; };

// GUARD (CURVIS_PRECISION_F64_FAST): `guard_eps` is the ray's relative state-error budget.  The texel is a truncation of
// the continuous image coordinates (fx, fy); when either lies closer to an integer than the budget propagated to the
// image (direction error <= eps * (4 + 2 |d_z| / sin theta): the phi component carries 1/sin^2 theta; d phi_img =
// d direction / sin theta_img), nothing is written and false is returned — the caller queues the ray for re-integration.
template <class Shape, class Trig, bool GUARD>
__device__ __forceinline__ bool finish_ray(const FrameParams& p, const Ray& q, int side, uint32_t steps,
                                           unsigned long long ray, RayTally& tally, const RayDiag& diag, double guard_eps) {
    uint32_t tx = 0, ty = 0;
    bool clamped = false;
    uint32_t rgba = 0;
    float4 tap = make_float4(0.f, 0.f, 0.f, 255.f);        // Rgba([0, 0, 0, 255])
    const bool want_pixel = p.out_rgb8 || p.out_rgba32f || p.n_peers;
    if (want_pixel && side != 0) {
        const Background& bg = p.bg[side > 0 ? 0 : 1];
        double dx, dy, dz, s;
        tangent_direction<Shape, Trig>(p, q, dx, dy, dz, s);                        // metrics.rs:339-349
        bool lookup = true;
        if (p.frame != CURVIS_FRAME_LOCAL) lookup = rotate_tangent_to_world(q.th, q.ph, dx, dy, dz);   // systems.rs:144-187
        if (lookup) {
            double fx, fy, sin_img;
            image_coordinates(bg, dx, dy, dz, fx, fy, sin_img);
            if (GUARD && guard_eps > 0.0) {
                const double e_dir = guard_eps * (4.0 + 2.0 * fabs(dz) / (fabs(s) * norm3(dx, dy, dz)));
                const double ey = e_dir * (double)bg.height * (1.0 / CURVIS_PI);
                const double ex = e_dir * (double)bg.width * (0.5 / CURVIS_PI) / fmax(sin_img, 1e-300);
                const double mx = fabs(fx - rint(fx)), my = fabs(fy - rint(fy));
                if (!(mx > ex && my > ey)) return false;                              // NaN coordinates fail the test too
            }
            clamped = nearest_texel(bg, fx, fy, tx, ty);
            if (p.sampling == CURVIS_SAMPLING_BILINEAR) {
                tap = bilinear_tap(bg, fx, fy);
                rgba = quantize_channel(tap.x) | (quantize_channel(tap.y) << 8) | (quantize_channel(tap.z) << 16);
            } else {
                rgba = __ldg(bg.texels + (size_t)ty * bg.width + tx);
                tap = make_float4((float)(rgba & 0xffu), (float)((rgba >> 8) & 0xffu), (float)((rgba >> 16) & 0xffu), (float)(rgba >> 24));
            }
        } else {
            clamped = true;   // the reference panics here; black pixel, counted in n_clamped
        }
    }
    if (side > 0) ++tally.pos; else if (side < 0) ++tally.neg; else ++tally.none;   // none: black, systems.rs:556-558
    tally.steps += steps;
    if (clamped) ++tally.clamped;
    if (want_pixel) {
        if (p.out_rgb8) {
            uint8_t* o = p.out_rgb8 + ray * 3ull;                                   // put_pixel on ImageRgb8 drops alpha (:324)
            o[0] = (uint8_t)(rgba & 0xffu);
            o[1] = (uint8_t)((rgba >> 8) & 0xffu);
            o[2] = (uint8_t)((rgba >> 16) & 0xffu);
        }
        if (p.out_rgba32f) p.out_rgba32f[ray] = tap;
        if (p.n_peers) {
            // fused all-gather: this pixel into the complete frame of every peer (frame_params.h)
            const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
            unsigned long long f; uint32_t px, k;
            split_ray_index(p, ray, tile_rays, f, px, k);
            uint32_t py;
            tile_pixel(p, k, px, px, py);
            const size_t off = (((size_t)f * p.height + py) * p.frame_width + px) * 3;
            for (uint32_t i = 0; i < p.n_peers; ++i) {
                uint8_t* o = p.out_peers[i] + off;
                o[0] = (uint8_t)(rgba & 0xffu);
                o[1] = (uint8_t)((rgba >> 8) & 0xffu);
                o[2] = (uint8_t)((rgba >> 16) & 0xffu);
            }
        }
    }
    if (p.records) {
        curvis_ray_record rec;
        rec.l = q.l; rec.theta = q.th; rec.phi = q.ph;
        rec.p_l = q.pl; rec.p_theta = q.pth; rec.p_phi = q.pph;
        rec.steps = steps; rec.side = side; rec.texel_x = tx; rec.texel_y = ty;
        rec.min_abs_sin_theta = diag.min_abs_sin; rec.stiffness = diag.stiffness;
        p.records[ray] = rec;
    }
    return true;
}

__device__ __forceinline__ void flush_tally(const FrameParams& p, RayTally t, unsigned lane) {
    for (int o = 16; o > 0; o >>= 1) {
        t.steps += __shfl_down_sync(0xffffffffu, t.steps, o);
        t.pos += __shfl_down_sync(0xffffffffu, t.pos, o);
        t.neg += __shfl_down_sync(0xffffffffu, t.neg, o);
        t.none += __shfl_down_sync(0xffffffffu, t.none, o);
        t.clamped += __shfl_down_sync(0xffffffffu, t.clamped, o);
    }
    if (lane == 0) {
        atomicAdd(&p.counters->total_steps, t.steps);
        if (t.pos) atomicAdd(&p.counters->n_positive, (unsigned long long)t.pos);
        if (t.neg) atomicAdd(&p.counters->n_negative, (unsigned long long)t.neg);
        if (t.none) atomicAdd(&p.counters->n_not_escaped, (unsigned long long)t.none);
        if (t.clamped) atomicAdd(&p.counters->n_clamped, (unsigned long long)t.clamped);
        unsigned warpid, smid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicAdd(&p.counters->slot_steps[warpid & 63u], t.steps);
        if (smid < 192u) atomicAdd(&p.counters->sm_steps[smid], t.steps);
    }
}

}  // namespace curvis
