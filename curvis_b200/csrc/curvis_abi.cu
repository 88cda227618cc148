// curvis_abi.cu — implementation of include/curvis_gpu.h: contexts, background residency,
// frame launches (single tile, device-resident tile, whole frame row-tiled over the
// context's devices).  The boundary replaces RelativisticSystem::render_image
// (reference src/systems.rs:307-330); there is no CPU compute path in this library.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "../../include/curvis_gpu.h"
#include "frame_params.h"
#include "host_error.h"
#include "launch.h"
#include "efficient.h"
#include "efficient_params.h"
#include "shape_table.h"

namespace curvis {

static thread_local std::string g_thread_error;
static std::atomic<uint64_t> g_kernel_launches{0};
int set_thread_error(int code, const char* msg) { g_thread_error = msg ? msg : ""; return code; }
const char* thread_error() { return g_thread_error.c_str(); }

struct DeviceState {
    int ordinal = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // resident scene
    uint32_t* bg_texels[2] = {nullptr, nullptr};
    float4* bg_texels_f4[2] = {nullptr, nullptr};   // staged on first CURVIS_SAMPLING_BILINEAR use
    uint32_t bg_w[2] = {0, 0}, bg_h[2] = {0, 0};
    // per-launch scratch
    DeviceCounters* d_counters = nullptr;
    DeviceCounters* h_counters = nullptr;  // pinned
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;  // pinned staging
    curvis_ray_record* d_records = nullptr; size_t d_records_cap = 0;
    CameraBlock* d_cameras = nullptr; size_t d_cameras_cap = 0;   // batched launches
    double2* d_shape_tab = nullptr;   // Interstellar shape-function table (shape_table.h), uploaded at context creation
    float4* d_shape_tab32 = nullptr;  // its fp32 edition
    double* d_atan_tab = nullptr;     // atan / ln tables of the operation-for-operation Interstellar step (shape_table.h)
    double* d_log_tab = nullptr;
    double2* d_inv_tab = nullptr;     // per-metric table of 1/r and r' (shape_table.h), rebuilt when (rho, m) change
    double inv_tab_rho = 0.0, inv_tab_m = 0.0;
    cudaEvent_t chunk_done[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // pipelined read-back
    // CURVIS_PRECISION_F64_FAST: indices of the rays inside the guard band, re-integrated by the parity kernel
    unsigned long long* d_redo = nullptr; size_t d_redo_cap = 0;
    // CURVIS_PRECISION_F64_FAST: indices of the rays the work queue hands out first (FrameParams::long_list)
    unsigned long long* d_long = nullptr; size_t d_long_cap = 0;
    // the per-launch scratch above (counters, cameras, redo list, events) is shared by every launch of the context: a
    // launch on a stream other than the previous one's first waits for the previous launch (launch_fence)
    cudaStream_t last_stream = nullptr; bool launched = false;
    // tile of the frame in flight
    uint32_t row_begin = 0, row_end = 0;
};

}  // namespace curvis

struct curvis_ctx {
    std::vector<curvis::DeviceState> devs;
    double bg_inv_rot[2][9];
    bool bg_set[2] = {false, false};
    curvis::LaunchTuning tuning;
    std::string err;
    std::vector<double> inv_tab_host;   // host copy of the last per-metric shape table built (shared by the devices)
    double inv_tab_host_rho = 0.0, inv_tab_host_m = 0.0;
    // caller-owned frame buffers page-locked by curvis_host_register: frames are DMA'd straight into them
    struct HostRegion { uint8_t* base; size_t bytes; };
    std::vector<HostRegion> host_regions;
    bool is_registered(const uint8_t* p, size_t bytes) const {
        for (const auto& r : host_regions)
            if (p >= r.base && bytes <= r.bytes && (size_t)(p - r.base) <= r.bytes - bytes) return true;
        return false;
    }
};

namespace curvis {

static int fail(curvis_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_thread_error = msg;
    return code;
}

static int cuda_fail(curvis_ctx* ctx, cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return fail(ctx, e == cudaErrorMemoryAllocation ? CURVIS_ERR_OUT_OF_MEMORY : CURVIS_ERR_CUDA, msg);
}

#define CURVIS_CUDA(ctx, expr)                                         \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return curvis::cuda_fail((ctx), _e, #expr); \
    } while (0)

static int validate_frame(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* cam, const curvis_sim* sim,
                          uint32_t row_begin, uint32_t row_end) {
    if (!ctx) return fail(nullptr, CURVIS_ERR_INVALID_ARGUMENT, "null context");
    if (!metric || !cam || !sim) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null metric/camera/sim");
    int rc = curvis_metric_validate(metric);
    if (rc != CURVIS_OK) return fail(ctx, rc, thread_error());
    if (cam->resolution_width == 0 || cam->resolution_height == 0)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "resolution_width and resolution_height must be greater than 0");
    if (row_begin > row_end || row_end > cam->resolution_height)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "row range outside the frame");
    if (sim->precision != CURVIS_PRECISION_F64 && sim->precision != CURVIS_PRECISION_F32 &&
        sim->precision != CURVIS_PRECISION_F64_FAST)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "unknown precision");
    if (sim->sampling != CURVIS_SAMPLING_NEAREST && sim->sampling != CURVIS_SAMPLING_BILINEAR)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "unknown sampling mode");
    if (sim->integrator != CURVIS_INTEGRATOR_EULER && sim->integrator != CURVIS_INTEGRATOR_RK4 &&
        sim->integrator != CURVIS_INTEGRATOR_EULER_ADAPTIVE)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "unknown integrator");
    if (sim->integrator != CURVIS_INTEGRATOR_EULER && sim->precision != CURVIS_PRECISION_F64)
        return fail(ctx, CURVIS_ERR_UNSUPPORTED, "the RK4 and adaptive-step extensions are implemented for CURVIS_PRECISION_F64 only");
    if (sim->integrator == CURVIS_INTEGRATOR_EULER_ADAPTIVE && !(sim->step_tolerance > 0.0))
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "CURVIS_INTEGRATOR_EULER_ADAPTIVE needs step_tolerance > 0");
    if (sim->frame != CURVIS_FRAME_LOCAL && sim->frame != CURVIS_FRAME_WORLD && sim->frame != CURVIS_FRAME_WORLD_QUIRK)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "unknown frame");
    if (sim->coordinates != CURVIS_COORDINATES_SPHERICAL && sim->coordinates != CURVIS_COORDINATES_CARTESIAN)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "unknown coordinates");
    if (sim->coordinates == CURVIS_COORDINATES_CARTESIAN &&
        (sim->precision == CURVIS_PRECISION_F32 || sim->integrator != CURVIS_INTEGRATOR_EULER))
        return fail(ctx, CURVIS_ERR_UNSUPPORTED, "CURVIS_COORDINATES_CARTESIAN is implemented for the fp64 precisions with the Euler integrator");
    if (!ctx->bg_set[0] || !ctx->bg_set[1])
        return fail(ctx, CURVIS_ERR_NO_BACKGROUND, "both backgrounds must be set before rendering");
    // escape_photon panics when the photon starts beyond the radius (systems.rs:122-124);
    // every ray starts at the camera, so the check is per frame.  NaN compares false, as in Rust.
    if (std::fabs(cam->position[1]) > sim->max_radius)
        return fail(ctx, CURVIS_ERR_CAMERA_OUTSIDE_RADIUS, "Photon already beyond the maximum radius. Cannot evaluate escape.");
    return CURVIS_OK;
}

static cudaError_t launch_render(const FrameParams& p, const curvis_metric* metric, const curvis_sim* sim,
                                 const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    if (sim->coordinates == CURVIS_COORDINATES_CARTESIAN)
        return sim->precision == CURVIS_PRECISION_F64_FAST ? launch_render_cart_fast(p, metric->kind, t, sm_count, stream)
                                                           : launch_render_cart(p, metric->kind, t, sm_count, stream);
    if (sim->precision == CURVIS_PRECISION_F32) return launch_render_f32(p, metric->kind, t, sm_count, stream);
    if (sim->precision == CURVIS_PRECISION_F64_FAST) {
        const bool prepass = render_f64_fast_has_prepass(p, metric->kind, t, sm_count);
        if (prepass) g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = launch_render_f64_fast(p, metric->kind, t, sm_count, stream);
        if (e != cudaSuccess || !p.redo_list) return e;
        // second launch: the parity kernel over the rays the fast kernel left in its guard band (list mode; the list's
        // length stays on the device)
        FrameParams r = p;
        r.ray_list = p.redo_list;
        r.ray_list_count = &p.counters->n_reintegrated;
        r.redo_list = nullptr;
        r.list_from_end = prepass ? 0u : 1u;
        r.window = 32;
        LaunchTuning rt = t;
        rt.blocks_per_sm = t.redo_blocks_per_sm;
        rt.kernel_variant = 4;
        g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
        return launch_render_f64(r, metric->kind, rt, sm_count, stream);
    }
    if (render_f64_has_prepass(p, t, sm_count)) g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_render_f64(p, metric->kind, t, sm_count, stream);
}

// Every launch of a context shares its per-device scratch (counters with the work-queue cursor, the camera block, the redo
// list, the timing events).  Launches on one stream are ordered by the stream; a launch on a DIFFERENT stream than the
// previous one waits for the previous launch to finish before it touches the scratch.
static cudaError_t launch_fence(DeviceState& d, cudaStream_t stream) {
    cudaError_t e = cudaSuccess;
    if (d.launched && d.last_stream != stream) e = cudaStreamWaitEvent(stream, d.ev_end, 0);
    d.last_stream = stream; d.launched = true;
    return e;
}

// CURVIS_PRECISION_F64_FAST + Interstellar: the per-metric table of 1/r(x) and r'(x) (shape_table.h) on device d, built on
// the host (~20 ms) the first time a (rho, m) pair is seen and uploaded stream-ordered.
static int ensure_inverse_table(curvis_ctx* ctx, DeviceState& d, const curvis_metric* metric, const curvis_sim* sim, cudaStream_t stream) {
    if (sim->precision != CURVIS_PRECISION_F64_FAST || metric->kind != CURVIS_METRIC_INTERSTELLAR) return CURVIS_OK;
    if (d.d_inv_tab && d.inv_tab_rho == metric->rho && d.inv_tab_m == metric->m) return CURVIS_OK;
    const size_t n = kInvTabIntervals * kShapeTabDoubles;
    if (ctx->inv_tab_host.size() != n || ctx->inv_tab_host_rho != metric->rho || ctx->inv_tab_host_m != metric->m) {
        ctx->inv_tab_host.resize(n);
        build_interstellar_inverse_table(metric->rho, metric->m, ctx->inv_tab_host.data());
        ctx->inv_tab_host_rho = metric->rho; ctx->inv_tab_host_m = metric->m;
    }
    if (!d.d_inv_tab) CURVIS_CUDA(ctx, cudaMalloc(&d.d_inv_tab, n * sizeof(double)));
    CURVIS_CUDA(ctx, cudaMemcpyAsync(d.d_inv_tab, ctx->inv_tab_host.data(), n * sizeof(double), cudaMemcpyHostToDevice, stream));
    // the table's last row holds its own device address (the kernel reads it from there: fast_f64.cuh, InverseShapeCache::reset)
    const unsigned long long self = (unsigned long long)(uintptr_t)d.d_inv_tab;
    CURVIS_CUDA(ctx, cudaMemcpyAsync((double*)d.d_inv_tab + kInvTabSelfRow * kShapeTabDoubles, &self, sizeof self, cudaMemcpyHostToDevice, stream));
    CURVIS_CUDA(ctx, cudaStreamSynchronize(stream));   // the host copy may be rebuilt by the next call
    d.inv_tab_rho = metric->rho; d.inv_tab_m = metric->m;
    return CURVIS_OK;
}

// The redo list of CURVIS_PRECISION_F64_FAST: one slot per ray of the launch (8 bytes each; a 4K frame: 66 MB), grown on demand.
// (Also the longest-first list of the same kernel: an eighth of the launch, at least 4096 slots; a list that overflows is ignored.)
static int ensure_redo(curvis_ctx* ctx, DeviceState& d, const curvis_sim* sim, size_t rays) {
    const bool strict_list = sim->precision == CURVIS_PRECISION_F64 && sim->coordinates == CURVIS_COORDINATES_SPHERICAL &&
                             sim->integrator == CURVIS_INTEGRATOR_EULER;       // the longest-first list of the default kernel
    if (!strict_list && (sim->precision != CURVIS_PRECISION_F64_FAST || ctx->tuning.fast_variant != 1 ||
                         sim->coordinates != CURVIS_COORDINATES_SPHERICAL)) return CURVIS_OK;
    if (ctx->tuning.longest_first) {
        const size_t want_long = std::max(size_t(4096), rays / 8);
        if (want_long > d.d_long_cap) {
            if (d.d_long) cudaFree(d.d_long);
            d.d_long = nullptr; d.d_long_cap = 0;
            CURVIS_CUDA(ctx, cudaMalloc(&d.d_long, want_long * sizeof(unsigned long long)));
            d.d_long_cap = want_long;
        }
    }
    if (!ctx->tuning.guard || strict_list) return CURVIS_OK;
    // one slot per ray up to 2^24 rays (128 MB: two 4K frames, half an 8K frame); beyond that an eighth of the launch (the
    // band takes ~1e-3 of the rays, 2.5 % with "guard" = 2); a full list makes the kernel re-integrate in line (correct, slow)
    const size_t want = rays <= (size_t(1) << 24) ? rays : std::max(size_t(1) << 24, rays / 8);
    if (want > d.d_redo_cap) {
        if (d.d_redo) cudaFree(d.d_redo);
        d.d_redo = nullptr; d.d_redo_cap = 0;
        CURVIS_CUDA(ctx, cudaMalloc(&d.d_redo, want * sizeof(unsigned long long)));
        d.d_redo_cap = want;
    }
    return CURVIS_OK;
}

static void fill_camera(const curvis_metric* metric, const curvis_camera* cam, CameraBlock& c) {
    std::memcpy(c.cam_pos, cam->position, sizeof c.cam_pos);
    std::memcpy(c.cam_to_world, cam->cam_to_world, sizeof c.cam_to_world);
    c.cam_r = host_shape_r(*metric, cam->position[1]);
    c.cam_sin_theta = host_sin(cam->position[2]);
    host_camera_basis(cam->position[2], cam->position[3], c.cam_n, c.cam_eth, c.cam_eph);
    c.focal_length = cam->focal_length; c.sensor_width = cam->sensor_width; c.sensor_height = cam->sensor_height;
}

static void fill_params(const curvis_ctx* ctx, const DeviceState& d, const curvis_metric* metric, const curvis_camera* cam,
                        const curvis_sim* sim, uint32_t row_begin, uint32_t row_end, uint8_t* d_out,
                        curvis_ray_record* d_records, FrameParams& p) {
    std::memset(&p, 0, sizeof p);
    p.rho = metric->rho; p.m = metric->m; p.a = metric->a;
    fill_camera(metric, cam, p.cam);
    p.cameras = nullptr; p.n_frames = 1;
    // auto: 32 steps between refill points.  CURVIS_PRECISION_F64_FAST has more per-window work (sin/cos re-derived
    // from theta) and shorter steps, so its window grows with the expected ray length (R + |l_camera|) / |delta| — a
    // photon moves ~delta per step — as length/16 clamped to [32, 128]: 128 at the default settings (2100 steps per
    // ray), 32 for a 150-step frame (tools/window_sweep.py: 4K Ellis 38.4 / 37.2 / 36.7 / 36.5 ms at 32 / 64 / 128 / 256)
    uint32_t auto_window = 32;
    if (sim->precision == CURVIS_PRECISION_F64_FAST) {
        const double expected_steps = (std::fabs(sim->max_radius) + std::fabs(cam->position[1])) / std::fabs(sim->delta);
        auto_window = expected_steps >= 2048.0 ? 128u : (expected_steps >= 512.0 ? (uint32_t)(expected_steps / 16.0) : 32u);   // NaN -> 32
    }
    p.window = (uint32_t)(ctx->tuning.window > 0 ? ctx->tuning.window : auto_window);
    p.width = cam->resolution_width; p.height = cam->resolution_height;
    p.max_iterations = sim->max_iterations; p.sampling = (uint32_t)sim->sampling;
    p.integrator = (uint32_t)sim->integrator;
    p.max_radius = sim->max_radius; p.delta = sim->delta;
    {
        uint64_t bits;
        std::memcpy(&bits, &sim->max_radius, 8);
        p.gate_hi = (sim->max_radius >= 0.0) ? (uint32_t)((bits >> 32) & 0x7fffffffu) : 0u;
    }
    p.row_begin = row_begin; p.row_end = row_end; p.row_stride = 1;
    p.favoured_slots = (uint32_t)ctx->tuning.favoured_slots;
    p.frame_width = cam->resolution_width; p.blocks_per_row = 1;
    p.inv_frame_width = 1.0 / (double)cam->resolution_width; p.inv_blocks_per_row = 1.0;
    p.inv_width = 1.0 / (double)cam->resolution_width;
    p.inv_tile_rays = 1.0 / ((double)(row_end - row_begin) * (double)cam->resolution_width);
    p.f_rho = (float)metric->rho; p.f_rho2 = (float)(metric->rho * metric->rho);
    p.f_m = (float)metric->m; p.f_a = (float)metric->a;
    p.f_xscale = (float)(2.0 / (3.14159265358979323846 * metric->m));
    p.f_delta = (float)sim->delta;
    p.d_rho2 = metric->rho * metric->rho;
    p.d_xscale = 2.0 / (3.14159265358979323846 * metric->m);
    p.shape_tab = d.d_shape_tab;
    p.shape_tab32 = d.d_shape_tab32;
    p.atan_tab = d.d_atan_tab; p.log_tab = d.d_log_tab;
    p.d_pim = 3.14159265358979323846 * metric->m;          // PI * self.m (metrics.rs:461), one rounding
    p.d_pim_rcp = 1.0 / p.d_pim;                           // correctly rounded (IEEE division on the host)
    p.inv_tab = d.d_inv_tab;
    p.d_xoff = -metric->a * p.d_xscale;
    p.fast_l_limit = metric->kind == CURVIS_METRIC_INTERSTELLAR ? interstellar_table_l_limit(metric->m, metric->a) : HUGE_VAL;
    // below this |l| no escape test is needed in the fp32 kernel (4 steps of slack at |p_l| <= ~1.3)
    p.f_near_radius = (float)(std::fabs(sim->max_radius) - 6.0 * std::fabs(sim->delta) - 1e-3 * std::fabs(sim->max_radius));
    for (int s = 0; s < 2; ++s) {
        p.bg[s].texels = d.bg_texels[s];
        p.bg[s].texels_f4 = d.bg_texels_f4[s];
        p.bg[s].width = d.bg_w[s]; p.bg[s].height = d.bg_h[s];
        std::memcpy(p.bg[s].inv_rot, ctx->bg_inv_rot[s], sizeof p.bg[s].inv_rot);
    }
    p.out_rgb8 = d_out; p.out_rgba32f = nullptr; p.records = d_records; p.counters = d.d_counters;
    p.frame = (uint32_t)sim->frame; p.coordinates = (uint32_t)sim->coordinates; p.step_tolerance = sim->step_tolerance;
    const bool guard = sim->precision == CURVIS_PRECISION_F64_FAST && ctx->tuning.guard && ctx->tuning.fast_variant == 1 &&
                       sim->coordinates == CURVIS_COORDINATES_SPHERICAL;
    p.redo_list = guard ? d.d_redo : nullptr;
    p.redo_capacity = guard ? d.d_redo_cap : 0;
    if (guard && ctx->tuning.redo_capacity_limit > 0 && (unsigned long long)ctx->tuning.redo_capacity_limit < p.redo_capacity)
        p.redo_capacity = (unsigned long long)ctx->tuning.redo_capacity_limit;
    p.guard_rel = ctx->tuning.guard_rel;
    p.guard_kicked = ctx->tuning.guard >= 2 ? 1u : 0u;
    const bool longest_first = ctx->tuning.longest_first && sim->coordinates == CURVIS_COORDINATES_SPHERICAL &&
                               ((sim->precision == CURVIS_PRECISION_F64_FAST && ctx->tuning.fast_variant == 1) ||
                                (sim->precision == CURVIS_PRECISION_F64 && sim->integrator == CURVIS_INTEGRATOR_EULER));
    p.long_list = longest_first ? d.d_long : nullptr;
    p.long_capacity = longest_first ? d.d_long_cap : 0;
}

// CURVIS_SAMPLING_BILINEAR reads the backgrounds as float4 (one 128-bit load per tap): staged once
// per background and device, on first use, by a conversion kernel (stream-ordered on `stream`).
static int ensure_float_backgrounds(curvis_ctx* ctx, DeviceState& d, const curvis_sim* sim, cudaStream_t stream) {
    if (sim->sampling != CURVIS_SAMPLING_BILINEAR) return CURVIS_OK;
    for (int s = 0; s < 2; ++s) {
        if (d.bg_texels_f4[s]) continue;
        const size_t n = (size_t)d.bg_w[s] * d.bg_h[s];
        CURVIS_CUDA(ctx, cudaMalloc(&d.bg_texels_f4[s], n * sizeof(float4)));
        CURVIS_CUDA(ctx, launch_texels_to_float4(d.bg_texels[s], d.bg_texels_f4[s], n, stream));
    }
    return CURVIS_OK;
}

// Enqueue one tile on `stream` of device d: zero counters, kernel bracketed by events.
static int enqueue_tile(curvis_ctx* ctx, DeviceState& d, const curvis_metric* metric, const curvis_camera* cam,
                        const curvis_sim* sim, uint32_t row_begin, uint32_t row_end, uint8_t* d_out,
                        curvis_ray_record* d_records, cudaStream_t stream, float4* d_out_f32 = nullptr) {
    CURVIS_CUDA(ctx, launch_fence(d, stream));
    int frc = ensure_float_backgrounds(ctx, d, sim, stream);
    if (frc != CURVIS_OK) return frc;
    frc = ensure_redo(ctx, d, sim, (size_t)(row_end - row_begin) * cam->resolution_width);
    if (frc != CURVIS_OK) return frc;
    frc = ensure_inverse_table(ctx, d, metric, sim, stream);
    if (frc != CURVIS_OK) return frc;
    FrameParams p;
    fill_params(ctx, d, metric, cam, sim, row_begin, row_end, d_out, d_records, p);
    p.out_rgba32f = d_out_f32;
    CURVIS_CUDA(ctx, cudaMemsetAsync(d.d_counters, 0, sizeof(DeviceCounters), stream));
    CURVIS_CUDA(ctx, cudaEventRecord(d.ev_begin, stream));
    if (row_end > row_begin) CURVIS_CUDA(ctx, launch_render(p, metric, sim, ctx->tuning, d.sm_count, stream));
    CURVIS_CUDA(ctx, cudaEventRecord(d.ev_end, stream));
    return CURVIS_OK;
}

static void add_counters(const DeviceCounters& c, uint64_t n_rays, curvis_stats* s) {
    s->total_steps += c.total_steps; s->n_rays += n_rays;
    s->n_positive += c.n_positive; s->n_negative += c.n_negative; s->n_not_escaped += c.n_not_escaped;
    s->n_clamped += c.n_clamped; s->n_reintegrated += c.n_reintegrated; s->n_kicked += c.n_kicked;
}

static int ensure_capacity(curvis_ctx* ctx, DeviceState& d, size_t out_bytes, size_t n_records, bool staging) {
    if (out_bytes > d.d_out_cap) {
        if (d.d_out) cudaFree(d.d_out);
        d.d_out = nullptr; d.d_out_cap = 0;
        CURVIS_CUDA(ctx, cudaMalloc(&d.d_out, out_bytes));
        d.d_out_cap = out_bytes;
    }
    if (staging && out_bytes > d.h_out_cap) {
        if (d.h_out) cudaFreeHost(d.h_out);
        d.h_out = nullptr; d.h_out_cap = 0;
        CURVIS_CUDA(ctx, cudaMallocHost(&d.h_out, out_bytes));
        d.h_out_cap = out_bytes;
    }
    if (n_records > d.d_records_cap) {
        if (d.d_records) cudaFree(d.d_records);
        d.d_records = nullptr; d.d_records_cap = 0;
        CURVIS_CUDA(ctx, cudaMalloc(&d.d_records, n_records * sizeof(curvis_ray_record)));
        d.d_records_cap = n_records;
    }
    return CURVIS_OK;
}

// Frame read-back: D2H into the pinned staging buffer in up to 8 chunks, each followed by an
// event, so the host's copy of chunk k into the caller's (pageable) frame overlaps the DMA of
// chunk k+1.  enqueue_readback only enqueues; finish_readback blocks chunk by chunk.
// Read-back into a buffer registered with curvis_host_register: one DMA, no staging copy.
static int enqueue_readback_direct(curvis_ctx* ctx, DeviceState& d, uint8_t* dst, size_t bytes) {
    CURVIS_CUDA(ctx, cudaMemcpyAsync(dst, d.d_out, bytes, cudaMemcpyDeviceToHost, d.stream));
    return CURVIS_OK;
}

static int enqueue_readback(curvis_ctx* ctx, DeviceState& d, size_t bytes) {
    const size_t chunks = bytes >= (8u << 20) ? 8 : 1;
    const size_t step = ((bytes + chunks - 1) / chunks + 255) & ~size_t(255);
    for (size_t k = 0, off = 0; k < chunks && off < bytes; ++k, off += step) {
        const size_t len = off + step <= bytes ? step : bytes - off;
        if (!d.chunk_done[k]) CURVIS_CUDA(ctx, cudaEventCreateWithFlags(&d.chunk_done[k], cudaEventDisableTiming));
        CURVIS_CUDA(ctx, cudaMemcpyAsync(d.h_out + off, d.d_out + off, len, cudaMemcpyDeviceToHost, d.stream));
        CURVIS_CUDA(ctx, cudaEventRecord(d.chunk_done[k], d.stream));
    }
    return CURVIS_OK;
}

static int finish_readback(curvis_ctx* ctx, DeviceState& d, uint8_t* dst, size_t bytes) {
    const size_t chunks = bytes >= (8u << 20) ? 8 : 1;
    const size_t step = ((bytes + chunks - 1) / chunks + 255) & ~size_t(255);
    for (size_t k = 0, off = 0; k < chunks && off < bytes; ++k, off += step) {
        const size_t len = off + step <= bytes ? step : bytes - off;
        CURVIS_CUDA(ctx, cudaEventSynchronize(d.chunk_done[k]));
        std::memcpy(dst + off, d.h_out + off, len);
    }
    return CURVIS_OK;
}

static void release_device(DeviceState& d) {
    if (d.ordinal < 0) return;
    cudaSetDevice(d.ordinal);
    for (int s = 0; s < 2; ++s) if (d.bg_texels[s]) cudaFree(d.bg_texels[s]);
    for (int s = 0; s < 2; ++s) if (d.bg_texels_f4[s]) cudaFree(d.bg_texels_f4[s]);
    if (d.d_counters) cudaFree(d.d_counters);
    if (d.h_counters) cudaFreeHost(d.h_counters);
    if (d.d_out) cudaFree(d.d_out);
    if (d.h_out) cudaFreeHost(d.h_out);
    if (d.d_records) cudaFree(d.d_records);
    if (d.d_cameras) cudaFree(d.d_cameras);
    if (d.d_shape_tab) cudaFree(d.d_shape_tab);
    if (d.d_atan_tab) cudaFree(d.d_atan_tab);
    if (d.d_log_tab) cudaFree(d.d_log_tab);
    if (d.d_shape_tab32) cudaFree(d.d_shape_tab32);
    if (d.d_inv_tab) cudaFree(d.d_inv_tab);
    if (d.d_redo) cudaFree(d.d_redo);
    if (d.d_long) cudaFree(d.d_long);
    for (auto& ev : d.chunk_done) if (ev) cudaEventDestroy(ev);
    if (d.ev_begin) cudaEventDestroy(d.ev_begin);
    if (d.ev_end) cudaEventDestroy(d.ev_end);
    if (d.stream) cudaStreamDestroy(d.stream);
    d = DeviceState();
}

}  // namespace curvis

using namespace curvis;

extern "C" int curvis_abi_version(void) { return CURVIS_ABI_VERSION; }

extern "C" uint64_t curvis_kernel_launch_count(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

extern "C" const char* curvis_last_error(const curvis_ctx* ctx) { return ctx ? ctx->err.c_str() : thread_error(); }

extern "C" int curvis_ctx_device_count(const curvis_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

extern "C" int curvis_ctx_create(const int* devices, int n_devices, curvis_ctx** out) {
    if (!out) return fail(nullptr, CURVIS_ERR_INVALID_ARGUMENT, "curvis_ctx_create: null out");
    *out = nullptr;
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0) {
        cudaGetLastError();
        return fail(nullptr, CURVIS_ERR_NO_DEVICE,
                    "no CUDA device visible (libcurvis_b200 has no CPU fallback; it needs an sm_100 GPU)");
    }
    std::vector<int> ords;
    if (devices && n_devices > 0) ords.assign(devices, devices + n_devices);
    else for (int i = 0; i < visible; ++i) ords.push_back(i);
    curvis_ctx* ctx = new (std::nothrow) curvis_ctx();
    if (!ctx) return fail(nullptr, CURVIS_ERR_OUT_OF_MEMORY, "out of host memory");
    for (int s = 0; s < 2; ++s) {
        const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        std::memcpy(ctx->bg_inv_rot[s], ident, sizeof ident);
    }
    ctx->devs.resize(ords.size());
    std::vector<double> shape_tab(kShapeTabIntervals * kShapeTabDoubles);
    build_interstellar_shape_table(shape_tab.data());
    std::vector<float> shape_tab32(kShapeTab32Intervals * kShapeTab32Floats);
    build_interstellar_shape_table_f32(shape_tab32.data());
    std::vector<double> atan_tab(kAtanTabIntervals * kFnTabDoubles), log_tab(kLogTabIntervals * kFnTabDoubles);
    build_atan_table(atan_tab.data());
    build_log_table(log_tab.data());
    for (size_t i = 0; i < ords.size(); ++i) {
        DeviceState& d = ctx->devs[i];
        const int ord = ords[i];
        int rc = CURVIS_OK;
        cudaDeviceProp prop;
        if (ord < 0 || ord >= visible) rc = fail(nullptr, CURVIS_ERR_INVALID_ARGUMENT, "device ordinal out of range");
        else if ((e = cudaSetDevice(ord)) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaSetDevice");
        else if ((e = cudaGetDeviceProperties(&prop, ord)) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaGetDeviceProperties");
        else if (prop.major != 10) rc = fail(nullptr, CURVIS_ERR_NO_DEVICE, "device is not sm_100 (this library ships sm_100a code only)");
        else {
            d.ordinal = ord;
            d.sm_count = prop.multiProcessorCount;
            if ((e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaStreamCreate");
            else if ((e = cudaEventCreate(&d.ev_begin)) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaEventCreate");
            else if ((e = cudaEventCreate(&d.ev_end)) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaEventCreate");
            else if ((e = cudaMalloc(&d.d_counters, sizeof(DeviceCounters))) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaMalloc(counters)");
            else if ((e = cudaMallocHost(&d.h_counters, sizeof(DeviceCounters))) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaMallocHost(counters)");
            else if ((e = cudaMalloc(&d.d_shape_tab, shape_tab.size() * sizeof(double))) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaMalloc(shape table)");
            else if ((e = cudaMemcpy(d.d_shape_tab, shape_tab.data(), shape_tab.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
                rc = cuda_fail(nullptr, e, "cudaMemcpy(shape table)");
            else if ((e = cudaMalloc(&d.d_atan_tab, atan_tab.size() * sizeof(double))) != cudaSuccess ||
                     (e = cudaMalloc(&d.d_log_tab, log_tab.size() * sizeof(double))) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaMalloc(atan / ln tables)");
            else if ((e = cudaMemcpy(d.d_atan_tab, atan_tab.data(), atan_tab.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
                     (e = cudaMemcpy(d.d_log_tab, log_tab.data(), log_tab.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
                rc = cuda_fail(nullptr, e, "cudaMemcpy(atan / ln tables)");
            else if ((e = cudaMalloc(&d.d_shape_tab32, shape_tab32.size() * sizeof(float))) != cudaSuccess) rc = cuda_fail(nullptr, e, "cudaMalloc(shape table f32)");
            else if ((e = cudaMemcpy(d.d_shape_tab32, shape_tab32.data(), shape_tab32.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
                rc = cuda_fail(nullptr, e, "cudaMemcpy(shape table f32)");
        }
        if (rc != CURVIS_OK) {
            curvis_ctx_destroy(ctx);
            return rc;
        }
    }
    *out = ctx;
    return CURVIS_OK;
}

extern "C" void curvis_ctx_destroy(curvis_ctx* ctx) {
    if (!ctx) return;
    for (const auto& r : ctx->host_regions) cudaHostUnregister(r.base);
    for (auto& d : ctx->devs) release_device(d);
    delete ctx;
}

extern "C" int curvis_set_background(curvis_ctx* ctx, int side, const uint8_t* rgba8,
                                     uint32_t width, uint32_t height, const double inv_rot[9]) {
    if (!ctx) return fail(nullptr, CURVIS_ERR_INVALID_ARGUMENT, "null context");
    if (side == 0 || !rgba8 || width == 0 || height == 0)
        return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_set_background: side must be +-1, image non-empty");
    const int s = side > 0 ? 0 : 1;
    const size_t bytes = (size_t)width * height * 4;
    for (auto& d : ctx->devs) {
        CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
        if (d.bg_texels[s]) { cudaFree(d.bg_texels[s]); d.bg_texels[s] = nullptr; }
        if (d.bg_texels_f4[s]) { cudaFree(d.bg_texels_f4[s]); d.bg_texels_f4[s] = nullptr; }
        CURVIS_CUDA(ctx, cudaMalloc(&d.bg_texels[s], bytes));
        CURVIS_CUDA(ctx, cudaMemcpyAsync(d.bg_texels[s], rgba8, bytes, cudaMemcpyHostToDevice, d.stream));
        d.bg_w[s] = width; d.bg_h[s] = height;
    }
    for (auto& d : ctx->devs) {
        CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
        CURVIS_CUDA(ctx, cudaStreamSynchronize(d.stream));
    }
    const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(ctx->bg_inv_rot[s], inv_rot ? inv_rot : ident, sizeof ident);
    ctx->bg_set[s] = true;
    return CURVIS_OK;
}

extern "C" int curvis_render_rows_device(curvis_ctx* ctx, const curvis_metric* metric,
                                         const curvis_camera* camera, const curvis_sim* sim,
                                         uint32_t row_begin, uint32_t row_end,
                                         void* d_out_rgb8_rows, void* d_records,
                                         void* stream, curvis_stats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = validate_frame(ctx, metric, camera, sim, row_begin, row_end);
    if (rc != CURVIS_OK) return rc;
    if (!d_out_rgb8_rows && row_end > row_begin) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null device output");
    DeviceState& d = ctx->devs[0];
    cudaStream_t st = (cudaStream_t)stream;
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    rc = enqueue_tile(ctx, d, metric, camera, sim, row_begin, row_end, (uint8_t*)d_out_rgb8_rows,
                      (curvis_ray_record*)d_records, st);
    if (rc != CURVIS_OK) return rc;
    if (stats) {
        CURVIS_CUDA(ctx, cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, st));
        CURVIS_CUDA(ctx, cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof *stats);
        add_counters(*d.h_counters, (uint64_t)(row_end - row_begin) * camera->resolution_width, stats);
        float ms = 0.f;
        CURVIS_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end));
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return CURVIS_OK;
}

static int render_frames_impl(curvis_ctx* ctx, const curvis_metric* metric,
                              const curvis_camera* cameras, uint32_t n_frames, const curvis_sim* sim,
                              uint32_t row_begin, uint32_t row_end,
                              void* d_out_rgb8_tiles, void* const* d_peer_frames, uint32_t n_peers, uint32_t row_stride, void* stream,
                              curvis_stats* stats, uint32_t block_width = 0) {
    // block_width != 0 (curvis_render_frames_peers_blocks): row_begin / row_end / row_stride count BLOCKS of block_width pixels
    // of a frame row, numbered row-major over the frame (frame_params.h)
    const auto t0 = std::chrono::steady_clock::now();
    if (!ctx) return fail(nullptr, CURVIS_ERR_INVALID_ARGUMENT, "null context");
    if (!cameras || n_frames == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "no cameras");
    uint32_t blocks_per_row = 1;
    if (block_width) {
        const uint32_t W0 = cameras[0].resolution_width;
        if (W0 == 0 || W0 % block_width) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "block_width must divide the frame width");
        blocks_per_row = W0 / block_width;
        if (row_begin > row_end || (uint64_t)row_end > (uint64_t)cameras[0].resolution_height * blocks_per_row)
            return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "block range outside the frame");
    }
    const uint32_t tile_width = block_width ? block_width : cameras[0].resolution_width;
    for (uint32_t f = 0; f < n_frames; ++f) {
        int rc = block_width ? validate_frame(ctx, metric, &cameras[f], sim, 0, 0) : validate_frame(ctx, metric, &cameras[f], sim, row_begin, row_end);
        if (rc != CURVIS_OK) return rc;
        if (cameras[f].resolution_width != cameras[0].resolution_width || cameras[f].resolution_height != cameras[0].resolution_height)
            return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "all frames of a batch must share one resolution");
    }
    if (!d_out_rgb8_tiles && n_peers == 0 && row_end > row_begin) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null device output");
    if (n_peers > CURVIS_MAX_PEERS) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "more than CURVIS_MAX_PEERS peer buffers");
    for (uint32_t i = 0; i < n_peers; ++i)
        if (!d_peer_frames || !d_peer_frames[i]) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null peer frame buffer");
    DeviceState& d = ctx->devs[0];
    cudaStream_t st = (cudaStream_t)stream;
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    CURVIS_CUDA(ctx, launch_fence(d, st));
    if (n_frames > d.d_cameras_cap) {
        if (d.d_cameras) cudaFree(d.d_cameras);
        d.d_cameras = nullptr; d.d_cameras_cap = 0;
        CURVIS_CUDA(ctx, cudaMalloc(&d.d_cameras, (size_t)n_frames * sizeof(CameraBlock)));
        d.d_cameras_cap = n_frames;
    }
    {
        int frc = ensure_float_backgrounds(ctx, d, sim, st);
        if (frc != CURVIS_OK) return frc;
    }
    {
        const uint32_t rows_launch = (row_end - row_begin + row_stride - 1) / row_stride;
        int rrc = ensure_redo(ctx, d, sim, (size_t)rows_launch * tile_width * n_frames);
        if (rrc != CURVIS_OK) return rrc;
        rrc = ensure_inverse_table(ctx, d, metric, sim, st);
        if (rrc != CURVIS_OK) return rrc;
    }
    std::vector<CameraBlock> blocks(n_frames);
    for (uint32_t f = 0; f < n_frames; ++f) fill_camera(metric, &cameras[f], blocks[f]);
    // pageable source: the driver stages it before returning, so `blocks` may die at scope exit
    CURVIS_CUDA(ctx, cudaMemcpyAsync(d.d_cameras, blocks.data(), (size_t)n_frames * sizeof(CameraBlock), cudaMemcpyHostToDevice, st));
    FrameParams p;
    fill_params(ctx, d, metric, &cameras[0], sim, row_begin, row_end, (uint8_t*)d_out_rgb8_tiles, nullptr, p);
    p.cameras = d.d_cameras; p.n_frames = n_frames;
    p.n_peers = n_peers;
    for (uint32_t i = 0; i < n_peers; ++i) p.out_peers[i] = (uint8_t*)d_peer_frames[i];
    // strided rows: the kernels index a tile of n_rows rows; p.row_end - p.row_begin is that count
    const uint32_t n_rows = (row_end - row_begin + row_stride - 1) / row_stride;
    p.row_stride = row_stride;
    p.row_end = row_begin + n_rows;
    p.width = tile_width; p.inv_width = 1.0 / (double)tile_width;
    p.blocks_per_row = blocks_per_row; p.inv_blocks_per_row = 1.0 / (double)blocks_per_row;
    p.inv_tile_rays = 1.0 / ((double)n_rows * (double)tile_width);
    CURVIS_CUDA(ctx, cudaMemsetAsync(d.d_counters, 0, sizeof(DeviceCounters), st));
    CURVIS_CUDA(ctx, cudaEventRecord(d.ev_begin, st));
    if (row_end > row_begin) CURVIS_CUDA(ctx, launch_render(p, metric, sim, ctx->tuning, d.sm_count, st));
    CURVIS_CUDA(ctx, cudaEventRecord(d.ev_end, st));
    if (stats) {
        CURVIS_CUDA(ctx, cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, st));
        CURVIS_CUDA(ctx, cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof *stats);
        add_counters(*d.h_counters, (uint64_t)n_rows * tile_width * n_frames, stats);
        float ms = 0.f;
        CURVIS_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end));
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return CURVIS_OK;
}

extern "C" int curvis_render_frames_device(curvis_ctx* ctx, const curvis_metric* metric,
                                           const curvis_camera* cameras, uint32_t n_frames, const curvis_sim* sim,
                                           uint32_t row_begin, uint32_t row_end,
                                           void* d_out_rgb8_tiles, void* stream, curvis_stats* stats) {
    return render_frames_impl(ctx, metric, cameras, n_frames, sim, row_begin, row_end, d_out_rgb8_tiles, nullptr, 0, 1, stream, stats);
}

extern "C" int curvis_render_frames_peers(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* cameras, uint32_t n_frames,
                                          const curvis_sim* sim, uint32_t row_begin, uint32_t row_end, uint32_t row_stride,
                                          void* const* d_frames, uint32_t n_peers, void* stream, curvis_stats* stats) {
    if (n_peers == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_render_frames_peers: no peer buffers");
    if (row_stride == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_render_frames_peers: row_stride must be at least 1");
    return render_frames_impl(ctx, metric, cameras, n_frames, sim, row_begin, row_end, nullptr, d_frames, n_peers, row_stride, stream, stats);
}

extern "C" int curvis_render_frames_peers_blocks(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* cameras, uint32_t n_frames,
                                                 const curvis_sim* sim, uint32_t block_begin, uint32_t block_end, uint32_t block_stride,
                                                 uint32_t block_width, void* const* d_frames, uint32_t n_peers, void* stream, curvis_stats* stats) {
    if (n_peers == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_render_frames_peers_blocks: no peer buffers");
    if (block_stride == 0 || block_width == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_render_frames_peers_blocks: block_stride and block_width must be at least 1");
    return render_frames_impl(ctx, metric, cameras, n_frames, sim, block_begin, block_end, nullptr, d_frames, n_peers, block_stride, stream, stats, block_width);
}

extern "C" int curvis_peer_buffer_create(curvis_ctx* ctx, size_t bytes, void** d_ptr, uint8_t ipc_handle[CURVIS_IPC_HANDLE_BYTES]) {
    if (!ctx || !d_ptr || !ipc_handle || bytes == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_peer_buffer_create: null argument or zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == CURVIS_IPC_HANDLE_BYTES, "IPC handle size");
    CURVIS_CUDA(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    void* ptr = nullptr;
    CURVIS_CUDA(ctx, cudaMalloc(&ptr, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); return cuda_fail(ctx, e, "cudaIpcGetMemHandle"); }
    std::memcpy(ipc_handle, &h, sizeof h);
    *d_ptr = ptr;
    return CURVIS_OK;
}

extern "C" int curvis_peer_buffer_open(curvis_ctx* ctx, const uint8_t ipc_handle[CURVIS_IPC_HANDLE_BYTES], void** d_ptr) {
    if (!ctx || !d_ptr || !ipc_handle) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_peer_buffer_open: null argument");
    CURVIS_CUDA(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof h);
    CURVIS_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CURVIS_OK;
}

extern "C" int curvis_peer_buffer_close(curvis_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_peer_buffer_close: null argument");
    CURVIS_CUDA(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    CURVIS_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return CURVIS_OK;
}

extern "C" int curvis_peer_buffer_destroy(curvis_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_peer_buffer_destroy: null argument");
    CURVIS_CUDA(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    CURVIS_CUDA(ctx, cudaFree(d_ptr));
    return CURVIS_OK;
}

extern "C" int curvis_render_rows(curvis_ctx* ctx, const curvis_metric* metric,
                                  const curvis_camera* camera, const curvis_sim* sim,
                                  uint32_t row_begin, uint32_t row_end,
                                  uint8_t* out_rgb8_rows, curvis_ray_record* records,
                                  curvis_stats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = validate_frame(ctx, metric, camera, sim, row_begin, row_end);
    if (rc != CURVIS_OK) return rc;
    const size_t n_rays = (size_t)(row_end - row_begin) * camera->resolution_width;
    if (n_rays && !out_rgb8_rows) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null output buffer");
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    const bool direct = n_rays && ctx->is_registered(out_rgb8_rows, n_rays * 3);
    rc = ensure_capacity(ctx, d, n_rays * 3 + 1, records ? n_rays : 0, !direct);
    if (rc != CURVIS_OK) return rc;
    rc = enqueue_tile(ctx, d, metric, camera, sim, row_begin, row_end, d.d_out, records ? d.d_records : nullptr, d.stream);
    if (rc != CURVIS_OK) return rc;
    if (n_rays) { rc = direct ? enqueue_readback_direct(ctx, d, out_rgb8_rows, n_rays * 3) : enqueue_readback(ctx, d, n_rays * 3); if (rc != CURVIS_OK) return rc; }
    if (records && n_rays)
        CURVIS_CUDA(ctx, cudaMemcpyAsync(records, d.d_records, n_rays * sizeof(curvis_ray_record), cudaMemcpyDeviceToHost, d.stream));
    CURVIS_CUDA(ctx, cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, d.stream));
    if (n_rays && !direct) { rc = finish_readback(ctx, d, out_rgb8_rows, n_rays * 3); if (rc != CURVIS_OK) return rc; }
    CURVIS_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        add_counters(*d.h_counters, n_rays, stats);
        float ms = 0.f;
        CURVIS_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end));
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return CURVIS_OK;
}

// One device's share of a frame whose rows are INTERLEAVED over the context's devices: rows row_begin, row_begin + stride, ...
// (n_rows of them).  d_frame != nullptr: the kernel stores every pixel at its place in that complete frame (a mapped host
// frame: zero copy); else the rows go, packed, to d_out_packed.
static int enqueue_rows_strided(curvis_ctx* ctx, DeviceState& d, const curvis_metric* metric, const curvis_camera* cam, const curvis_sim* sim,
                                uint32_t row_begin, uint32_t stride, uint32_t n_rows, uint8_t* d_out_packed, uint8_t* d_frame, cudaStream_t stream) {
    CURVIS_CUDA(ctx, launch_fence(d, stream));
    int rc = ensure_float_backgrounds(ctx, d, sim, stream);
    if (rc != CURVIS_OK) return rc;
    if ((rc = ensure_redo(ctx, d, sim, (size_t)n_rows * cam->resolution_width)) != CURVIS_OK) return rc;
    if ((rc = ensure_inverse_table(ctx, d, metric, sim, stream)) != CURVIS_OK) return rc;
    FrameParams p;
    fill_params(ctx, d, metric, cam, sim, row_begin, row_begin + n_rows, d_frame ? nullptr : d_out_packed, nullptr, p);
    p.row_stride = stride;
    if (d_frame) { p.n_peers = 1; p.out_peers[0] = d_frame; }
    CURVIS_CUDA(ctx, cudaMemsetAsync(d.d_counters, 0, sizeof(DeviceCounters), stream));
    CURVIS_CUDA(ctx, cudaEventRecord(d.ev_begin, stream));
    if (n_rows) CURVIS_CUDA(ctx, launch_render(p, metric, sim, ctx->tuning, d.sm_count, stream));
    CURVIS_CUDA(ctx, cudaEventRecord(d.ev_end, stream));
    return CURVIS_OK;
}

extern "C" int curvis_render_image(curvis_ctx* ctx, const curvis_metric* metric,
                                   const curvis_camera* camera, const curvis_sim* sim,
                                   uint8_t* out_rgb8, curvis_stats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = validate_frame(ctx, metric, camera, sim, 0, camera ? camera->resolution_height : 0);
    if (rc != CURVIS_OK) return rc;
    if (!out_rgb8) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null output buffer");
    const uint32_t W = camera->resolution_width, H = camera->resolution_height;
    const size_t n = ctx->devs.size();
    const bool direct = ctx->is_registered(out_rgb8, (size_t)W * H * 3);
    if (n > 1) {
        // Several devices: rows INTERLEAVED (device g renders rows g, g + n, ...), so every device gets the same mix of
        // short and long rays (a contiguous tile of central rows carries ~3 % more steps than the mean).
        // Every pixel is independent (systems.rs:316-326), so no exchange is needed: into a registered frame the kernels store
        // their pixels in place; otherwise each device's packed rows are scattered by one strided copy.
        for (size_t g = 0; g < n; ++g) {
            DeviceState& d = ctx->devs[g];
            d.row_begin = (uint32_t)g;
            d.row_end = (g < H) ? (uint32_t)((H - g + n - 1) / n) : 0u;      // here: the NUMBER of rows of this device
            const size_t bytes = (size_t)d.row_end * W * 3;
            CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
            const bool in_place = direct && ctx->tuning.zero_copy;
            uint8_t* d_frame = nullptr;
            if (in_place) CURVIS_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_frame, out_rgb8, 0));
            else if ((rc = ensure_capacity(ctx, d, bytes + 1, 0, false)) != CURVIS_OK) return rc;
            rc = enqueue_rows_strided(ctx, d, metric, camera, sim, d.row_begin < H ? d.row_begin : H, (uint32_t)n, d.row_end, d.d_out, d_frame, d.stream);
            if (rc != CURVIS_OK) return rc;
            CURVIS_CUDA(ctx, cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, d.stream));
            if (bytes && !in_place)
                CURVIS_CUDA(ctx, cudaMemcpy2DAsync(out_rgb8 + g * (size_t)W * 3, n * (size_t)W * 3, d.d_out, (size_t)W * 3, (size_t)W * 3, d.row_end,
                                                   cudaMemcpyDeviceToHost, d.stream));
        }
        if (stats) std::memset(stats, 0, sizeof *stats);
        for (size_t g = 0; g < n; ++g) {
            DeviceState& d = ctx->devs[g];
            CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
            CURVIS_CUDA(ctx, cudaStreamSynchronize(d.stream));
            if (stats) {
                add_counters(*d.h_counters, (uint64_t)d.row_end * W, stats);
                float ms = 0.f;
                CURVIS_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end));
                if (ms > stats->kernel_ms) stats->kernel_ms = ms;
            }
        }
        if (stats) stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return CURVIS_OK;
    }
    // One device: the whole frame, straight into a registered caller frame or through the pinned staging buffer.
    for (size_t g = 0; g < n; ++g) {
        DeviceState& d = ctx->devs[g];
        d.row_begin = (uint32_t)((uint64_t)H * g / n);
        d.row_end = (uint32_t)((uint64_t)H * (g + 1) / n);
        const size_t bytes = (size_t)(d.row_end - d.row_begin) * W * 3;
        CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
        rc = ensure_capacity(ctx, d, bytes + 1, 0, !direct);
        if (rc != CURVIS_OK) return rc;
        uint8_t* tile_out = out_rgb8 + (size_t)d.row_begin * W * 3;
        // zero_copy: the kernel stores its pixels straight into the registered (mapped) host frame
        uint8_t* kernel_out = d.d_out;
        if (direct && ctx->tuning.zero_copy) CURVIS_CUDA(ctx, cudaHostGetDevicePointer((void**)&kernel_out, tile_out, 0));
        rc = enqueue_tile(ctx, d, metric, camera, sim, d.row_begin, d.row_end, kernel_out, nullptr, d.stream);
        if (rc != CURVIS_OK) return rc;
        CURVIS_CUDA(ctx, cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, d.stream));
        if (bytes && kernel_out == d.d_out) { rc = direct ? enqueue_readback_direct(ctx, d, tile_out, bytes) : enqueue_readback(ctx, d, bytes); if (rc != CURVIS_OK) return rc; }
    }
    if (stats) std::memset(stats, 0, sizeof *stats);
    for (size_t g = 0; g < n; ++g) {
        DeviceState& d = ctx->devs[g];
        const size_t bytes = (size_t)(d.row_end - d.row_begin) * W * 3;
        CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
        if (bytes && !direct) { rc = finish_readback(ctx, d, out_rgb8 + (size_t)d.row_begin * W * 3, bytes); if (rc != CURVIS_OK) return rc; }
        CURVIS_CUDA(ctx, cudaStreamSynchronize(d.stream));
        if (stats) {
            add_counters(*d.h_counters, (uint64_t)(d.row_end - d.row_begin) * W, stats);
            float ms = 0.f;
            CURVIS_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end));
            if (ms > stats->kernel_ms) stats->kernel_ms = ms;
        }
    }
    if (stats) stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return CURVIS_OK;
}

extern "C" int curvis_host_register(curvis_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr || bytes == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_host_register: null context / buffer or zero size");
    for (const auto& r : ctx->host_regions)
        if (r.base == (uint8_t*)ptr) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_host_register: buffer already registered");
    CURVIS_CUDA(ctx, cudaSetDevice(ctx->devs[0].ordinal));
    CURVIS_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    ctx->host_regions.push_back({(uint8_t*)ptr, bytes});
    return CURVIS_OK;
}

extern "C" int curvis_host_unregister(curvis_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_host_unregister: null argument");
    for (size_t i = 0; i < ctx->host_regions.size(); ++i) {
        if (ctx->host_regions[i].base != (uint8_t*)ptr) continue;
        ctx->host_regions.erase(ctx->host_regions.begin() + (long)i);
        CURVIS_CUDA(ctx, cudaHostUnregister(ptr));
        return CURVIS_OK;
    }
    return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "curvis_host_unregister: buffer was not registered with this context");
}

extern "C" int curvis_measure_fma_peak(curvis_ctx* ctx, double* fp64_tflops, double* fp32_tflops) {
    if (!ctx || !fp64_tflops || !fp32_tflops) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    CURVIS_CUDA(ctx, measure_fma_peak(d.sm_count, d.stream, fp64_tflops, fp32_tflops));
    return CURVIS_OK;
}

extern "C" int curvis_ctx_set_option(curvis_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    const std::string k(key);
    if (k == "kernel_variant" && value >= 0 && value <= 5) ctx->tuning.kernel_variant = (int)value;
    else if (k == "blocks_per_sm" && value >= 0 && value <= 32) ctx->tuning.blocks_per_sm = (int)value;
    else if (k == "window" && value >= 0 && value <= 4096) ctx->tuning.window = (int)value;
    else if (k == "fast_variant" && value >= 0 && value <= 1) ctx->tuning.fast_variant = (int)value;
    else if (k == "zero_copy" && value >= 0 && value <= 1) ctx->tuning.zero_copy = (int)value;
    else if (k == "guard" && value >= 0 && value <= 2) ctx->tuning.guard = (int)value;
    else if (k == "fast_regs" && (value == 0 || value == 96 || value == 128)) ctx->tuning.fast_regs = (int)value;
    else if (k == "redo_blocks_per_sm" && value >= 0 && value <= 32) ctx->tuning.redo_blocks_per_sm = (int)value;
    else if (k == "redo_ahead" && value >= 0 && value <= 1) ctx->tuning.redo_ahead = (int)value;
    else if (k == "redo_capacity_limit" && value >= 0) ctx->tuning.redo_capacity_limit = value;
    else if (k == "longest_first" && value >= 0 && value <= 2) ctx->tuning.longest_first = (int)value;
    else if (k == "favoured_slots" && value >= 0 && value <= 64) ctx->tuning.favoured_slots = (int)value;
    else if (k == "guard_rel_e15" && value >= 1 && value <= 1000000000000ll) ctx->tuning.guard_rel = (double)value * 1e-15;
    else return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "unknown option or value out of range: " + k);
    return CURVIS_OK;
}

extern "C" int curvis_debug_last_step_shares(curvis_ctx* ctx, uint64_t slot_steps[64], uint64_t sm_steps[192]) {
    if (!ctx || !slot_steps || !sm_steps) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    const DeviceCounters& c = *ctx->devs[0].h_counters;   // as read back by the last launch that asked for stats
    for (int i = 0; i < 64; ++i) slot_steps[i] = c.slot_steps[i];
    for (int i = 0; i < 192; ++i) sm_steps[i] = c.sm_steps[i];
    return CURVIS_OK;
}

extern "C" int curvis_debug_eval(curvis_ctx* ctx, int op, const double* a, const double* b, double* out, size_t n) {
    if (!ctx || !a || !out) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    double *da = nullptr, *db = nullptr, *dout = nullptr;
    const size_t bytes = n * sizeof(double);
    int rc = CURVIS_OK;
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&da, bytes ? bytes : 8)) == cudaSuccess && (e = cudaMalloc(&dout, bytes ? bytes : 8)) == cudaSuccess &&
        (!b || (e = cudaMalloc(&db, bytes ? bytes : 8)) == cudaSuccess) &&
        (e = cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, d.stream)) == cudaSuccess &&
        (!b || (e = cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, d.stream)) == cudaSuccess) &&
        (e = (op == 13 || op == 14)   ? launch_debug_shape(d.d_shape_tab, op - 13, da, dout, n, d.stream)
             : (op == 15 || op == 16) ? launch_debug_shape32(d.d_shape_tab32, op - 15, da, dout, n, d.stream)
             : (op == 17 || op == 18) ? launch_debug_atan_log(d.d_atan_tab, d.d_log_tab, op - 17, da, dout, n, d.stream)
                                      : launch_debug_eval(op, da, db, dout, n, d.stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, d.stream)) == cudaSuccess)
        e = cudaStreamSynchronize(d.stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "curvis_debug_eval");
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return rc;
}

extern "C" int curvis_debug_inverse_shape(curvis_ctx* ctx, const curvis_metric* metric, const double* x, double* y, double* g, size_t n) {
    if (!ctx || !metric || !x || !y || !g) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    if (metric->kind != CURVIS_METRIC_INTERSTELLAR) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "the inverse shape table belongs to the Interstellar metric");
    int rc = curvis_metric_validate(metric);
    if (rc != CURVIS_OK) return fail(ctx, rc, thread_error());
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    curvis_sim sim; std::memset(&sim, 0, sizeof sim); sim.precision = CURVIS_PRECISION_F64_FAST;
    rc = ensure_inverse_table(ctx, d, metric, &sim, d.stream);
    if (rc != CURVIS_OK) return rc;
    double *dx = nullptr, *dy = nullptr, *dg = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(double);
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&dx, bytes)) == cudaSuccess && (e = cudaMalloc(&dy, bytes)) == cudaSuccess && (e = cudaMalloc(&dg, bytes)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, d.stream)) == cudaSuccess &&
        (e = launch_debug_inverse_shape(d.d_inv_tab, dx, dy, dg, n, d.stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, d.stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(g, dg, n * sizeof(double), cudaMemcpyDeviceToHost, d.stream)) == cudaSuccess)
        e = cudaStreamSynchronize(d.stream);
    cudaFree(dx); cudaFree(dy); cudaFree(dg);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "curvis_debug_inverse_shape");
    return CURVIS_OK;
}

extern "C" int curvis_debug_rhs_check(curvis_ctx* ctx, const curvis_metric* metric, uint64_t n_samples, uint64_t seed, uint64_t mismatches[4]) {
    if (!ctx || !metric || !mismatches) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    int rc = curvis_metric_validate(metric);
    if (rc != CURVIS_OK) return fail(ctx, rc, thread_error());
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    FrameParams p;
    std::memset(&p, 0, sizeof p);
    p.rho = metric->rho; p.m = metric->m; p.a = metric->a;
    unsigned long long* d_bad = nullptr;
    CURVIS_CUDA(ctx, cudaMalloc(&d_bad, 4 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_bad, 0, 4 * sizeof(unsigned long long), d.stream);
    if (e == cudaSuccess) e = launch_debug_rhs_check(p, metric->kind, seed, n_samples, d_bad, d.stream);
    unsigned long long h[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_bad, sizeof h, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(d_bad);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "curvis_debug_rhs_check");
    for (int k = 0; k < 4; ++k) mismatches[k] = h[k];
    return CURVIS_OK;
}

extern "C" int curvis_render_image_efficient(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* camera,
                                             const curvis_sim* sim, const curvis_sampling_settings* sampling,
                                             uint8_t* out_rgb8, double* dbg, curvis_stats* stats, curvis_efficient_info* info) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = validate_frame(ctx, metric, camera, sim, 0, camera ? camera->resolution_height : 0);
    if (rc != CURVIS_OK) return rc;
    if (!sampling || !out_rgb8) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null sampling settings / output buffer");
    if (sampling->alphas_num < 3) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "alphas_num must be at least 3");
    // F64: the table equals the CPU's bit for bit; F64_FAST: its photons are integrated by the regrouped kernel, whose
    // dependency chain per step is ~5x shorter — the table passes are latency-bound — at ~1e-13 relative in the table
    if (sim->precision != CURVIS_PRECISION_F64 && sim->precision != CURVIS_PRECISION_F64_FAST)
        return fail(ctx, CURVIS_ERR_UNSUPPORTED, "the table-based renderer is fp64 only (CURVIS_PRECISION_F64 or _F64_FAST)");
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    const uint32_t W = camera->resolution_width, H = camera->resolution_height;

    // ---- step 3 of the reference (systems.rs:437-486): the escape-angle table
    curvis_camera eq = *camera;                       // compute_escape_angle starts at (0, l, pi/2, 0), systems.rs:224-227
    eq.position[0] = 0.0; eq.position[2] = 3.14159265358979323846264338327950288 / 2.0; eq.position[3] = 0.0;
    double* d_dirs = nullptr; size_t d_dirs_cap = 0;
    uint64_t launches = 0;
    int cuda_rc = CURVIS_OK;
    auto integrate = [&](const double* dirs, size_t n, curvis_ray_record* out) -> int {
        if (n > d_dirs_cap) {
            if (d_dirs) cudaFree(d_dirs);
            d_dirs = nullptr; d_dirs_cap = 0;
            cudaError_t e = cudaMalloc(&d_dirs, n * 3 * sizeof(double));
            if (e != cudaSuccess) return cuda_rc = cuda_fail(ctx, e, "cudaMalloc(ray directions)");
            d_dirs_cap = n;
        }
        int r = ensure_capacity(ctx, d, 1, n, false);
        if (r != CURVIS_OK) return cuda_rc = r;
        if ((r = ensure_redo(ctx, d, sim, n)) != CURVIS_OK) return cuda_rc = r;
        if ((r = ensure_inverse_table(ctx, d, metric, sim, d.stream)) != CURVIS_OK) return cuda_rc = r;
        if (launch_fence(d, d.stream) != cudaSuccess) return cuda_rc = fail(ctx, CURVIS_ERR_CUDA, "launch fence");
        cudaError_t e = cudaMemcpyAsync(d_dirs, dirs, n * 3 * sizeof(double), cudaMemcpyHostToDevice, d.stream);
        if (e != cudaSuccess) return cuda_rc = cuda_fail(ctx, e, "cudaMemcpyAsync(ray directions)");
        FrameParams p;
        eq.resolution_width = (uint32_t)n; eq.resolution_height = 1;
        fill_params(ctx, d, metric, &eq, sim, 0, 1, nullptr, d.d_records, p);
        p.ray_dirs = d_dirs;
        // The table's photons carry no integer decision for a guard band to protect (the sampler reads escape angles, and its table
        // is accurate to ~5e-4 rad): with CURVIS_PRECISION_F64_FAST they take the raw regrouped kernel.  They are equatorial, so
        // the band would flag every one of them (theta = pi/2 sits exactly on a texel edge) and the pass would cost the strict
        // launch on top of the fast one.
        p.redo_list = nullptr; p.redo_capacity = 0;
        if ((e = cudaMemsetAsync(d.d_counters, 0, sizeof(DeviceCounters), d.stream)) != cudaSuccess ||
            (e = launch_render(p, metric, sim, ctx->tuning, d.sm_count, d.stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(out, d.d_records, n * sizeof(curvis_ray_record), cudaMemcpyDeviceToHost, d.stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(d.stream)) != cudaSuccess)
            return cuda_rc = cuda_fail(ctx, e, "escape-angle table launch");
        ++launches;
        return CURVIS_OK;
    };
    EscapeTable table;
    std::string err;
    rc = build_escape_table(*metric, camera->position[1], sampling->alphas_num, sampling->max_iterations_sampling,
                            sampling->threshold_1, sampling->threshold_2, integrate, table, err);
    if (d_dirs) cudaFree(d_dirs);
    if (rc != CURVIS_OK) return cuda_rc != CURVIS_OK ? cuda_rc : fail(ctx, rc, err);
    const double table_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    // ---- steps 1, 2, 4, 5 (systems.rs:388-433, :489-523): one thread per pixel
    EfficientParams ep;
    std::memset(&ep, 0, sizeof ep);
    fill_camera(metric, camera, ep.cam);
    ep.width = W; ep.height = H; ep.row_begin = 0; ep.row_end = H;
    host_vector3_from_theta_phi(camera->position[2], camera->position[3], ep.cam_pos_bg);
    const double ex[3] = {1.0, 0.0, 0.0};
    if (!host_rotation_from_two_vectors(ex, ep.cam_pos_bg, ep.rot_bg))              // systems.rs:409 panics here
        return fail(ctx, CURVIS_ERR_PARALLEL_VECTORS, "v1 and v2 must not be parallel");
    const size_t n_pts = table.alphas.size(), n_seg = table.m_e.size();
    const size_t table_doubles = n_pts + 4 * n_seg;
    std::vector<double> packed;
    packed.reserve(table_doubles);
    packed.insert(packed.end(), table.alphas.begin(), table.alphas.end());
    packed.insert(packed.end(), table.m_e.begin(), table.m_e.end());
    packed.insert(packed.end(), table.c_e.begin(), table.c_e.end());
    packed.insert(packed.end(), table.m_s.begin(), table.m_s.end());
    packed.insert(packed.end(), table.c_s.begin(), table.c_s.end());
    double* d_table = nullptr;
    double* d_dbg = nullptr;
    const size_t px = (size_t)W * H;
    auto cleanup = [&]() { if (d_table) cudaFree(d_table); if (d_dbg) cudaFree(d_dbg); };
    cudaError_t e = cudaMalloc(&d_table, (table_doubles ? table_doubles : 1) * sizeof(double));
    if (e == cudaSuccess && dbg) e = cudaMalloc(&d_dbg, px * 3 * sizeof(double));
    if (e != cudaSuccess) { cleanup(); return cuda_fail(ctx, e, "cudaMalloc(table)"); }
    const bool direct = ctx->is_registered(out_rgb8, px * 3);     // registered caller frame: one DMA, no staging copy
    rc = ensure_capacity(ctx, d, px * 3 + 1, 0, !direct);
    if (rc != CURVIS_OK) { cleanup(); return rc; }
    ep.alphas = d_table; ep.m_e = d_table + n_pts; ep.c_e = ep.m_e + n_seg; ep.m_s = ep.c_e + n_seg; ep.c_s = ep.m_s + n_seg;
    ep.n_points = (uint32_t)n_pts; ep.n_segments = (uint32_t)n_seg;
    for (int s = 0; s < 2; ++s) {
        ep.bg[s].texels = d.bg_texels[s]; ep.bg[s].width = d.bg_w[s]; ep.bg[s].height = d.bg_h[s];
        std::memcpy(ep.bg[s].inv_rot, ctx->bg_inv_rot[s], sizeof ep.bg[s].inv_rot);
    }
    ep.out_rgb8 = d.d_out; ep.dbg = d_dbg; ep.counters = d.d_counters;
    if ((e = cudaMemcpyAsync(d_table, packed.data(), table_doubles * sizeof(double), cudaMemcpyHostToDevice, d.stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(d.d_counters, 0, sizeof(DeviceCounters), d.stream)) != cudaSuccess ||
        (e = cudaEventRecord(d.ev_begin, d.stream)) != cudaSuccess ||
        (e = launch_efficient_pixels(ep, d.sm_count, d.stream)) != cudaSuccess ||
        (e = cudaEventRecord(d.ev_end, d.stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(direct ? out_rgb8 : d.h_out, d.d_out, px * 3, cudaMemcpyDeviceToHost, d.stream)) != cudaSuccess ||
        (dbg && (e = cudaMemcpyAsync(dbg, d_dbg, px * 3 * sizeof(double), cudaMemcpyDeviceToHost, d.stream)) != cudaSuccess) ||
        (e = cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, d.stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(d.stream)) != cudaSuccess) {
        cleanup();
        return cuda_fail(ctx, e, "per-pixel pass");
    }
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    if (!direct) std::memcpy(out_rgb8, d.h_out, px * 3);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end);
    cleanup();
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        add_counters(*d.h_counters, px, stats);
        stats->total_steps = table.steps;
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (info) {
        info->table_points = (uint32_t)n_pts; info->table_passes = table.passes;
        info->table_evaluations = table.evaluations; info->table_steps = table.steps;
        info->table_ms = table_ms; info->pixels_ms = ms;
    }
    return CURVIS_OK;
}

extern "C" int curvis_render_rows_rgba32f(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* camera,
                                          const curvis_sim* sim, uint32_t row_begin, uint32_t row_end,
                                          float* out_rgba32f_rows, curvis_stats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = validate_frame(ctx, metric, camera, sim, row_begin, row_end);
    if (rc != CURVIS_OK) return rc;
    const size_t n_rays = (size_t)(row_end - row_begin) * camera->resolution_width;
    if (n_rays && !out_rgba32f_rows) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null output buffer");
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    float4* d_f32 = nullptr;
    CURVIS_CUDA(ctx, cudaMalloc(&d_f32, (n_rays ? n_rays : 1) * sizeof(float4)));
    rc = enqueue_tile(ctx, d, metric, camera, sim, row_begin, row_end, nullptr, nullptr, d.stream, d_f32);
    cudaError_t e = cudaSuccess;
    if (rc == CURVIS_OK && n_rays) e = cudaMemcpyAsync(out_rgba32f_rows, d_f32, n_rays * sizeof(float4), cudaMemcpyDeviceToHost, d.stream);
    if (rc == CURVIS_OK && e == cudaSuccess) e = cudaMemcpyAsync(d.h_counters, d.d_counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, d.stream);
    if (rc == CURVIS_OK && e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(d_f32);
    if (rc != CURVIS_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "curvis_render_rows_rgba32f");
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        add_counters(*d.h_counters, n_rays, stats);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end);
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return CURVIS_OK;
}

extern "C" int curvis_debug_bilinear(curvis_ctx* ctx, int side, const double* fx, const double* fy, float* out_rgba32f, size_t n) {
    if (!ctx || !fx || !fy || !out_rgba32f || side == 0) return fail(ctx, CURVIS_ERR_INVALID_ARGUMENT, "null argument");
    const int s = side > 0 ? 0 : 1;
    if (!ctx->bg_set[s]) return fail(ctx, CURVIS_ERR_NO_BACKGROUND, "background not set");
    DeviceState& d = ctx->devs[0];
    CURVIS_CUDA(ctx, cudaSetDevice(d.ordinal));
    curvis_sim bil; std::memset(&bil, 0, sizeof bil); bil.sampling = CURVIS_SAMPLING_BILINEAR;
    int rc = ensure_float_backgrounds(ctx, d, &bil, d.stream);
    if (rc != CURVIS_OK) return rc;
    double *dx = nullptr, *dy = nullptr; float4* dout = nullptr;
    cudaError_t e = cudaSuccess;
    const size_t nb = (n ? n : 1);
    Background bg;
    bg.texels = d.bg_texels[s]; bg.texels_f4 = d.bg_texels_f4[s]; bg.width = d.bg_w[s]; bg.height = d.bg_h[s];
    std::memcpy(bg.inv_rot, ctx->bg_inv_rot[s], sizeof bg.inv_rot);
    if ((e = cudaMalloc(&dx, nb * 8)) == cudaSuccess && (e = cudaMalloc(&dy, nb * 8)) == cudaSuccess &&
        (e = cudaMalloc(&dout, nb * sizeof(float4))) == cudaSuccess &&
        (e = cudaMemcpyAsync(dx, fx, n * 8, cudaMemcpyHostToDevice, d.stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dy, fy, n * 8, cudaMemcpyHostToDevice, d.stream)) == cudaSuccess &&
        (e = launch_debug_bilinear(bg, dx, dy, dout, n, d.stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(out_rgba32f, dout, n * sizeof(float4), cudaMemcpyDeviceToHost, d.stream)) == cudaSuccess)
        e = cudaStreamSynchronize(d.stream);
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "curvis_debug_bilinear");
    return CURVIS_OK;
}
