// frame_params.h — the one kernel argument of the render kernels: everything
// RelativisticSystem::render_image reads from `self` and its three arguments
// (reference src/systems.rs:68-73, :307-312), flattened to plain data so it travels in the
// kernel parameter space (constant bank, uniform loads) with no per-launch H2D copy.
#pragma once
#include <stdint.h>
#include <vector_types.h>
#include "../../include/curvis_gpu.h"

namespace curvis {

// Device-side counters of one launch (one 128-byte line; zeroed before each launch).
struct DeviceCounters {
    unsigned long long next_ray;      // work queue: next ray index of the tile to hand out
    unsigned long long total_steps;   // curvis_stats.total_steps
    unsigned long long n_positive;
    unsigned long long n_negative;
    unsigned long long n_not_escaped;
    unsigned long long n_clamped;
    unsigned long long n_reintegrated; // CURVIS_PRECISION_F64_FAST: rays pushed onto the re-integration list (= its length)
    unsigned long long redo_next;      // work queue of the re-integration launch
    unsigned long long n_kicked;       // CURVIS_PRECISION_F64_FAST: rays with stiffness >= 1
    unsigned long long n_long;         // CURVIS_PRECISION_F64_FAST: length of the longest-first list (render_f64_fast.cu: collect_long_rays)
    unsigned long long long_next;      // work queue of the longest-first list (claimed by the favoured warp slots first)
    unsigned long long _pad[5];
    // diagnostics (curvis_debug_last_step_shares): Euler steps executed per hardware warp slot (%warpid) and per SM (%smid) — every
    // warp of a persistent launch lives as long as the kernel, so these are the shares of the issue slots the scheduler handed out
    unsigned long long slot_steps[64];
    unsigned long long sm_steps[192];
};

struct Background {
    const uint32_t* texels;  // RGBA8 packed little-endian (R in the low byte), row-major
    const float4* texels_f4; // the same texels as float4 (0..255), staged on first bilinear use
    uint32_t width, height;
    double inv_rot[9];       // image orientation inverse, row-major (images.rs:132-142)
};

// Per-frame camera block (cameras.rs:30-43 after Camera::new, plus two per-frame uniforms).
struct CameraBlock {
    double cam_pos[4];
    double cam_to_world[9];
    // r(l_camera) and sin(theta_camera): per-frame uniforms of new_photon (metrics.rs:329-330),
    // evaluated once on the host with the platform libm like the reference does per ray
    double cam_r, cam_sin_theta;
    double focal_length, sensor_width, sensor_height;
    // CURVIS_COORDINATES_CARTESIAN: unit position vector of the camera and the unit vectors of increasing theta / phi
    // there (host libm, once per frame)
    double cam_n[3], cam_eth[3], cam_eph[3];
};

struct FrameParams {
    // metric parameters (metrics.rs:399-401, :431-435)
    double rho, m, a;
    // camera of the frame (n_frames == 1), in the constant bank
    CameraBlock cam;
    // batched launch (video): n_frames cameras in device memory, one tile of each frame per launch;
    // ray index = frame * tile_rays + pixel-in-tile, output tiles frame-major
    const CameraBlock* cameras;
    uint32_t n_frames;
    uint32_t _pad1;
    // explicit-ray launch (the escape-angle table of render_image_efficient): when non-null, ray i
    // starts at cam.cam_pos with tangent-space direction ray_dirs[3i..3i+2] instead of a camera
    // pixel; width = number of rays, one row; only `records` is written
    const double* ray_dirs;
    uint32_t width, height;
    // render_image arguments (systems.rs:309-311)
    uint32_t max_iterations;
    uint32_t sampling;
    double max_radius, delta;
    // 1 / width and 1 / (rays of one frame's tile), for splitting a ray index without 64-bit integer division (geodesic_f64.cuh)
    double inv_width, inv_tile_rays;
    // escape-test gate: the high word of max_radius when it is >= 0, else 0 (|l| > R needs |l|'s high word >= it)
    uint32_t gate_hi, _pad0;
    // tile of the frame this launch renders: rows [row_begin, row_end)
    uint32_t row_begin, row_end;
    // steps between two refill points of a warp (render_f64.cu), tuning knob
    uint32_t window;
    uint32_t integrator;   // curvis_integrator
    // fp32 copies of the uniforms for CURVIS_PRECISION_F32 (render_f32.cu), rounded once on the host
    float f_rho, f_rho2, f_m, f_a, f_xscale, f_delta, f_near_radius, _pad2;
    // fp64 uniforms of CURVIS_PRECISION_F64_FAST (render_f64_fast.cu): rho^2 and 2/(pi*m)
    double d_rho2, d_xscale;
    // Interstellar shape-function table of CURVIS_PRECISION_F64_FAST (shape_table.h), resident per device
    const double2* shape_tab;
    const float4* shape_tab32;   // the fp32 edition for CURVIS_PRECISION_F32
    // atan and ln tables of the operation-for-operation Interstellar step (shape_table.h), and pi m with its correctly rounded
    // reciprocal (host): x = 2 (|l| - a) / (pi m) is then one correctly rounded quotient in three fp64 instructions
    const double* atan_tab;
    const double* log_tab;
    double d_pim, d_pim_rcp;
    // per-metric table of 1/r^2 and r'/r^3 as functions of z = |l| - a (shape_table.h: build_interstellar_inverse_table), and
    // the |l| beyond which z leaves the table (+inf for the other metrics)
    const double2* inv_tab;
    double d_xoff, fast_l_limit;
    // scene (systems.rs:70-71): [0] = background_positive, [1] = background_negative
    Background bg[2];
    // Fused render + all-gather (curvis_render_frames_peers): n_peers > 0 = every finished ray stores its RGB8
    // into the COMPLETE frames of all peers (own device + peer devices mapped over NVLink), frame f at byte
    // offset f*W*H*3, so the row tiles are "gathered" by the render kernel itself and no collective moves pixels
    uint8_t* out_peers[CURVIS_MAX_PEERS];
    // row_stride > 1 (peers launches only): the launch renders rows row_begin + k*row_stride, k in [0, row_end - row_begin)
    // — interleaved row ownership, which gives every rank statistically the same work
    uint32_t n_peers, row_stride;
    // Block tiles (curvis_render_frames_peers_blocks): the "rows" above are BLOCKS of `width` consecutive pixels of a frame row,
    // numbered row-major over the frame (blocks_per_row = frame_width / width of them per row): interleaving blocks instead of
    // rows splits EVERY row over the ranks — the 10^4-step rays of a frame sit in two or three rows.  Whole rows: width =
    // frame_width, blocks_per_row = 1.  frame_width and height are what the camera and the output addresses see.
    uint32_t favoured_slots;   // longest-first refill: hardware warp slots (%warpid <) that claim the list first (render_f64_fast.cu)
    uint32_t list_from_end;    // list mode: walk ray_list from its last entry (render_f64.cu)
    uint32_t frame_width, blocks_per_row;
    double inv_frame_width, inv_blocks_per_row;
    // curvis_sim extensions (all 0 in parity mode): curvis_frame, curvis_coordinates, adaptive-step tolerance
    uint32_t frame, coordinates;
    double step_tolerance;
    // CURVIS_PRECISION_F64_FAST guard band (render_f64_fast.cu: guard_ok): a finished ray whose escape step or texel lies
    // closer to a decision boundary than the band is not written; its index goes to redo_list (capacity redo_capacity,
    // length counters->n_reintegrated) and the parity kernel re-integrates the list in a second launch (ray_list mode)
    unsigned long long* redo_list;
    unsigned long long redo_capacity;
    double guard_rel;        // relative state error budget of a ray with stiffness < 1
    uint32_t guard_kicked;   // 1: rays with stiffness >= 1 are re-integrated too ("guard" = 2)
    uint32_t _pad3;
    // longest-first refill (render_f64_fast.cu): a pre-pass kernel lists the rays predicted to graze a coordinate pole (the
    // 10^4-step stragglers); the work queue hands those out first.  Length in counters->n_long; a list that overflowed is ignored
    unsigned long long* long_list;
    unsigned long long long_capacity;
    // list mode (the second launch): ray i of the launch is ray ray_list[i] of the tile; the launch holds
    // *ray_list_count rays (device-resident count: the host never learns it before launching)
    const unsigned long long* ray_list;
    const unsigned long long* ray_list_count;
    // outputs: RGB8 rows of the tile (packed, row-major), optional per-ray records, counters
    uint8_t* out_rgb8;
    float4* out_rgba32f;     // optional: the unrounded colour of every ray (RGBA, 0..255 scale)
    curvis_ray_record* records;
    DeviceCounters* counters;
};

}  // namespace curvis
