// efficient_kernel.cu — per-pixel pass of the table-based renderer (reference
// src/systems.rs:400-433 and :489-523): pixel direction -> alpha -> interpolated escape angle and
// space -> axis-angle rotation of the camera-position direction -> texel.  One thread per pixel,
// fp64, reference operation order (compile with -fmad=false).
#include "geodesic_f64.cuh"
#include "efficient_params.h"
#include "launch.h"

namespace curvis {

__global__ void __launch_bounds__(256) efficient_pixels_kernel(const __grid_constant__ EfficientParams p) {
    const unsigned long long tile = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    unsigned acc_pos = 0, acc_neg = 0, acc_none = 0, acc_clamped = 0;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < tile;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t px = (uint32_t)(idx % p.width), py = p.row_begin + (uint32_t)(idx / p.width);
        double tx, ty, tz;
        outward_vector_on_world_space(p.cam, p.width, p.height, px, py, tx, ty, tz);          // systems.rs:408
        double bx, by, bz;
        mat3_mul(p.rot_bg, tx, ty, tz, bx, by, bz);                                          // :409
        const double cx = p.cam_pos_bg[0], cy = p.cam_pos_bg[1], cz = p.cam_pos_bg[2];
        const double ax = cy * bz - cz * by, ay = cz * bx - cx * bz, az = cx * by - cy * bx; // :412-414
        const double alpha = acos((tx * 1.0 + ty * 0.0) + tz * 0.0);                         // :428-431
        // interp_slice: index of the last table alpha strictly below alpha (0 if none), clamped
        uint32_t lo = 0, hi = p.n_points;                                                    // count of alphas < alpha
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.alphas[mid] < alpha) lo = mid + 1; else hi = mid;
        }
        uint32_t seg = lo ? lo - 1 : 0;
        if (seg > p.n_segments - 1) seg = p.n_segments - 1;
        const double angle = p.m_e[seg] * alpha + p.c_e[seg];                                // :489
        const double space = p.m_s[seg] * alpha + p.c_s[seg];                                // :491
        const double an = norm3(ax, ay, az);                                                 // Unit::new_normalize, :502
        const double ux = ax / an, uy = ay / an, uz = az / an;
        double m[9];
        if (angle != 0.0) {                                                                  // Rotation3::from_axis_angle
            const double sqx = ux * ux, sqy = uy * uy, sqz = uz * uz;
            double sn, cs;
            TrigFast::sincos(angle, sn, cs);
            const double omc = 1.0 - cs;
            m[0] = sqx + (1.0 - sqx) * cs;   m[1] = ux * uy * omc - uz * sn; m[2] = ux * uz * omc + uy * sn;
            m[3] = ux * uy * omc + uz * sn;  m[4] = sqy + (1.0 - sqy) * cs;  m[5] = uy * uz * omc - ux * sn;
            m[6] = ux * uz * omc - uy * sn;  m[7] = uy * uz * omc + ux * sn; m[8] = sqz + (1.0 - sqz) * cs;
        } else {
            m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1;
        }
        double fx, fy, fz;
        mat3_mul(m, cx, cy, cz, fx, fy, fz);                                                 // :503
        uint32_t rgba = 0;
        const int side = (space == 1.0) ? 1 : ((space == -1.0) ? -1 : 0);                    // :514-518 exact match
        if (side != 0) {
            const Background& bg = p.bg[side > 0 ? 0 : 1];
            uint32_t u, v;
            if (texel_from_direction(bg, fx, fy, fz, u, v)) ++acc_clamped;
            rgba = __ldg(bg.texels + (size_t)v * bg.width + u);
            if (side > 0) ++acc_pos; else ++acc_neg;
        } else {
            ++acc_none;
        }
        uint8_t* o = p.out_rgb8 + idx * 3ull;
        o[0] = (uint8_t)(rgba & 0xffu);
        o[1] = (uint8_t)((rgba >> 8) & 0xffu);
        o[2] = (uint8_t)((rgba >> 16) & 0xffu);
        if (p.dbg) { p.dbg[idx * 3] = alpha; p.dbg[idx * 3 + 1] = angle; p.dbg[idx * 3 + 2] = space; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc_pos += __shfl_down_sync(0xffffffffu, acc_pos, o);
        acc_neg += __shfl_down_sync(0xffffffffu, acc_neg, o);
        acc_none += __shfl_down_sync(0xffffffffu, acc_none, o);
        acc_clamped += __shfl_down_sync(0xffffffffu, acc_clamped, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        if (acc_pos) atomicAdd(&p.counters->n_positive, (unsigned long long)acc_pos);
        if (acc_neg) atomicAdd(&p.counters->n_negative, (unsigned long long)acc_neg);
        if (acc_none) atomicAdd(&p.counters->n_not_escaped, (unsigned long long)acc_none);
        if (acc_clamped) atomicAdd(&p.counters->n_clamped, (unsigned long long)acc_clamped);
    }
}

cudaError_t launch_efficient_pixels(const EfficientParams& p, int sm_count, cudaStream_t stream) {
    const unsigned long long tile = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    if (tile == 0) return cudaSuccess;
    unsigned long long want = (tile + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count * 8ull;
    efficient_pixels_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace curvis
