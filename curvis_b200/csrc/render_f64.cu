// render_f64.cu — the parity render kernel: one ray per lane, fp64, reference operation
// order (compile with -fmad=false, see geodesic_f64.cuh).
//
// Execution model.  The kernel is persistent: the grid is sized to exactly fill the machine
// (SM count x resident CTAs per SM) and every warp pulls rays from one global queue
// (DeviceCounters::next_ray).  A ray needs a data-dependent number of Euler steps (escape
// test after every step, reference src/systems.rs:126-135), so lanes of a warp finish at
// different times.  Every WINDOW steps the warp ballots the lanes that finished, runs their
// epilogues together (texel fetch + RGB8 store), and refills exactly those lanes from the
// queue with one warp-aggregated atomicAdd — finished lanes idle for at most WINDOW-1 steps
// (<1 % of a ~2000-step ray) instead of waiting for the slowest lane of the warp, and a ray
// that runs to max_iterations (NotEscaped, 40000 steps at defaults) costs one lane, not 32.
#include "geodesic_f64.cuh"
#include "launch.h"

namespace curvis {

constexpr int kBlock = 128;
constexpr int kWindow = 16;
constexpr unsigned kFull = 0xffffffffu;

template <class Shape, class Trig>
__global__ void __launch_bounds__(kBlock) render_rows_f64(const __grid_constant__ FrameParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;

    Ray q;
    bool live = false;      // this lane is integrating a ray
    bool pending = false;   // this lane finished a ray whose epilogue has not run yet
    bool drained = false;   // the queue is empty (warp-uniform)
    int side = 0;
    uint32_t steps = 0;
    unsigned long long ray = 0;

    unsigned long long acc_steps = 0;
    unsigned acc_pos = 0, acc_neg = 0, acc_none = 0, acc_clamped = 0;

    for (;;) {
        // ---- epilogue of the lanes that finished during the last window
        if (pending) {
            const uint32_t px = (uint32_t)(ray % p.width);
            uint32_t rgba = 0, tx = 0, ty = 0;
            if (side != 0) {
                const Background& bg = p.bg[side > 0 ? 0 : 1];
                if (escaped_texel<Shape, Trig>(p, q, bg, tx, ty)) ++acc_clamped;
                rgba = __ldg(bg.texels + (size_t)ty * bg.width + tx);
                if (side > 0) ++acc_pos; else ++acc_neg;
            } else {
                ++acc_none;  // systems.rs:556-558: black
            }
            uint8_t* o = p.out_rgb8 + ray * 3ull;  // put_pixel on ImageRgb8 drops alpha (:324)
            o[0] = (uint8_t)(rgba & 0xffu);
            o[1] = (uint8_t)((rgba >> 8) & 0xffu);
            o[2] = (uint8_t)((rgba >> 16) & 0xffu);
            if (p.records) {
                curvis_ray_record rec;
                rec.l = q.l; rec.theta = q.th; rec.phi = q.ph;
                rec.p_l = q.pl; rec.p_theta = q.pth; rec.p_phi = q.pph;
                rec.steps = steps; rec.side = side; rec.texel_x = tx; rec.texel_y = ty;
                p.records[ray] = rec;
            }
            (void)px;
            acc_steps += steps;
            pending = false;
        }

        // ---- refill idle lanes from the queue (one atomic per warp)
        const unsigned idle = __ballot_sync(kFull, !live);
        if (idle) {
            if (!drained) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(&p.counters->next_ray, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFull, base, leader);
                if (!live) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < tile_rays) {
                        ray = idx;
                        const uint32_t px = (uint32_t)(idx % p.width);
                        const uint32_t py = p.row_begin + (uint32_t)(idx / p.width);
                        new_photon_for_pixel<Shape, Trig>(p, px, py, q);
                        steps = 0;
                        side = 0;
                        if (p.max_iterations == 0) pending = true;  // loop of systems.rs:126 runs zero times
                        else live = true;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= tile_rays) drained = true;
            }
            if (__ballot_sync(kFull, live || pending) == 0u) break;
        }

        // ---- WINDOW Euler steps (escape_photon's loop body, systems.rs:126-135)
#pragma unroll 1
        for (int k = 0; k < kWindow; ++k) {
            if (live) {
                euler_step<Shape, Trig>(p, q);
                ++steps;
                if (q.l > p.max_radius) { side = 1; live = false; pending = true; }          // :129-131
                else if (q.l < -p.max_radius) { side = -1; live = false; pending = true; }   // :132-134
                else if (steps >= p.max_iterations) { side = 0; live = false; pending = true; }  // :137
            }
        }
    }

    // ---- per-warp reduction of the counters, one atomic per counter per warp
    for (int o = 16; o > 0; o >>= 1) {
        acc_steps += __shfl_down_sync(kFull, acc_steps, o);
        acc_pos += __shfl_down_sync(kFull, acc_pos, o);
        acc_neg += __shfl_down_sync(kFull, acc_neg, o);
        acc_none += __shfl_down_sync(kFull, acc_none, o);
        acc_clamped += __shfl_down_sync(kFull, acc_clamped, o);
    }
    if (lane == 0) {
        atomicAdd(&p.counters->total_steps, acc_steps);
        if (acc_pos) atomicAdd(&p.counters->n_positive, (unsigned long long)acc_pos);
        if (acc_neg) atomicAdd(&p.counters->n_negative, (unsigned long long)acc_neg);
        if (acc_none) atomicAdd(&p.counters->n_not_escaped, (unsigned long long)acc_none);
        if (acc_clamped) atomicAdd(&p.counters->n_clamped, (unsigned long long)acc_clamped);
    }
}

template <class Shape, class Trig>
static cudaError_t launch_one(const FrameParams& p, int sm_count, cudaStream_t stream) {
    static int blocks_per_sm = 0;  // per instantiation; same for every sm_100 device
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, render_rows_f64<Shape, Trig>, kBlock, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    unsigned long long want = (tile_rays + kBlock - 1) / kBlock;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    render_rows_f64<Shape, Trig><<<grid, kBlock, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_render_f64(const FrameParams& p, int metric_kind, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch_one<ShapeEllis, TrigCuda>(p, sm_count, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch_one<ShapeInterstellar, TrigCuda>(p, sm_count, stream);
    case CURVIS_METRIC_FLAT: return launch_one<ShapeFlat, TrigCuda>(p, sm_count, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace curvis
