// launch.h — host-callable launchers of the render kernels (one per precision TU).
#pragma once
#include <cuda_runtime.h>
#include "frame_params.h"
#include "launch_host.h"

namespace curvis {

// Tuning knobs of a context (curvis_ctx_set_option).
struct LaunchTuning {
    int kernel_variant = 5;  // 0 plain operators + CUDA sincos (round-1 v0); 1 unguarded IEEE sequences + CUDA sincos;
                             // 2 + in-kernel sincos; 3 lean loop: integer-pipe guards, gated escape test;
                             // 4 the lean loop with the step's six reciprocals built from two seeds;
                             // 5 (default) 4 with the next step's shape function and sincos carried across the loop's back edge
                             //   (geodesic_f64.cuh: euler_steps_ahead — 117 instructions per Ellis step instead of 131, and a
                             //   shorter dependent chain); every variant performs the same operations on the same values
    int blocks_per_sm = 0;   // 0 = occupancy maximum
    int window = 0;          // Euler steps between two refill points of a warp; 0 = auto (32; 32..128 in F64_FAST)
    int zero_copy = 1;       // curvis_render_image into a registered host frame: 1 (default) = the kernel stores its pixels
                             // straight into it (measured: kernel time unchanged, 0.04 ms exposed); 0 = device frame + one DMA (0.5 ms)
    int guard = 1;           // CURVIS_PRECISION_F64_FAST: 1 (default) = guard band + re-integration for the rays with stiffness < 1;
                             // 2 = kicked rays (stiffness >= 1) re-integrated too (every ray equals CURVIS_PRECISION_F64's);
                             // 0 = the raw regrouped kernel (A/B, tools/guard_study.py)
    double guard_rel = 1e-9; // relative state-error budget of a ray with stiffness < 1 (render_f64_fast.cu: guard_eps)
    long long redo_capacity_limit = 0;   // test knob: cap the re-integration list (0 = automatic) to exercise the in-line fallback
    int redo_blocks_per_sm = 2;   // CTAs per SM of the re-integration launch (a few per cent of the frame's rays: fewer, fuller warps)
    int redo_ahead = 1;      // the re-integration launch runs the step loop in latency form (geodesic_f64.cuh: euler_steps_ahead: the next
                             // step's shape function and sincos overlap this step's quotients); 0 = the throughput form of the frame kernel (A/B)
    int fast_regs = 0;       // CURVIS_PRECISION_F64_FAST register budget: 96 (5 CTAs per SM) or 128 (4 CTAs); 0 (default) = 96.  4K frames,
                             // 96 against 128: Ellis 36.6 / 37.1 ms, Interstellar 55.6 / 58.2 ms (profiles/r02_time_fast.json)
    int longest_first = 2;   // CURVIS_PRECISION_F64_FAST: rays predicted to be stragglers (near-critical, pole-grazing) are listed by a pre-pass
                             // kernel and claimed first, by the warp slots the schedulers favour: 1 = always, 0 = never (index order),
                             // 2 (default) = Ellis: always; the other metrics: in launches of at most 64 rays per lane of the grid, where
                             // the straggler's latency is comparable to the kernel's (render_f64_fast.cu).  CURVIS_PRECISION_F64 (kernel_variant 5)
                             // follows the same rule for its tiles (render_f64.cu: lean_longest_first)
    int favoured_slots = 8;  // longest-first refill: the listed rays are claimed first by the warps in hardware slots %warpid < this (the first two
                             // resident CTAs of an SM: 1.65x / 1.51x the mean share of their scheduler); 0 = no warp is favoured (the list is taken
                             // when the index walk is exhausted), 64 = every warp (round 2's first form)
    int fast_variant = 1;    // CURVIS_PRECISION_F64_FAST: 0 sin/cos from theta every step; 1 (default) (sin, cos) carried and
                             // rotated by the step's small dtheta, re-derived from theta once per window
};

// fp64 parity kernel (render_f64.cu, compiled with -fmad=false).
cudaError_t launch_render_f64(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream);

// fp32 fast mode, CURVIS_PRECISION_F32 (render_f32.cu) — extension, not the parity path.
cudaError_t launch_render_f32(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream);

// fp64 with a regrouped right-hand side, CURVIS_PRECISION_F64_FAST (render_f64_fast.cu) — extension.
cudaError_t launch_render_f64_fast(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream);
bool render_f64_fast_has_prepass(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count);   // one more kernel in front of it (longest-first list)

// The longest-first pre-pass (render_f64_fast.cu: collect_long_rays writes FrameParams::long_list) and the rule that decides whether a
// launch runs it, shared with the operation-for-operation kernel; kLongRaySin2 is the predicate's limit on sin^2 theta_min.
constexpr double kLongRaySin2 = 0.03 * 0.03;
bool longest_first_prepass_wanted(const FrameParams& p, int mode, int sm_count, bool whole_frames);
cudaError_t launch_collect_long_rays(const FrameParams& p, int sm_count, cudaStream_t stream);
bool render_f64_has_prepass(const FrameParams& p, const LaunchTuning& t, int sm_count);   // CURVIS_PRECISION_F64 launches it too (whole Euler frames / tiles)

// fp64, chart-free angular state, CURVIS_COORDINATES_CARTESIAN (render_f64_cart.cu) — extension ("pole-safe").
cudaError_t launch_render_cart(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream);
// the same scheme regrouped for the fp64 pipe (CURVIS_COORDINATES_CARTESIAN + CURVIS_PRECISION_F64_FAST)
cudaError_t launch_render_cart_fast(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream);

// RGBA8 -> float4 staging for CURVIS_SAMPLING_BILINEAR, and the tap evaluated at explicit
// continuous coordinates (test hook curvis_debug_bilinear).  render_f64.cu.
cudaError_t launch_texels_to_float4(const uint32_t* texels, float4* out, size_t n, cudaStream_t stream);
cudaError_t launch_debug_bilinear(const Background& bg, const double* fx, const double* fy, float4* out, size_t n, cudaStream_t stream);

// Elementwise evaluation of one device math primitive (test hook, curvis_debug_eval).
cudaError_t launch_debug_eval(int op, const double* a, const double* b, double* out, size_t n, cudaStream_t stream);



// Right-hand side of kernel_variant 4 against the plain operators on n pseudo-random states (test hook, curvis_debug_rhs_check).
cudaError_t launch_debug_rhs_check(const FrameParams& p, int metric_kind, unsigned long long seed, unsigned long long n,
                                   unsigned long long* d_mismatches, cudaStream_t stream);

// F(x) (which = 0) / G(x) (which = 1) of the Interstellar shape-function table as the fast kernel
// evaluates them (test hook, curvis_debug_eval ops 13 / 14).  render_f64_fast.cu.
cudaError_t launch_debug_shape(const double2* tab, int which, const double* x, double* out, size_t n, cudaStream_t stream);
// Y(x) = 1/r and G(x) of the per-metric table as fast_variant 1 evaluates them (x given; test hook curvis_debug_inverse_shape).
cudaError_t launch_debug_inverse_shape(const double2* tab, const double* x, double* y, double* g, size_t n, cudaStream_t stream);
// atan x (which = 0) / ln(1 + x^2) (which = 1) as the operation-for-operation Interstellar step evaluates them (ops 17 / 18).  render_f64.cu.
cudaError_t launch_debug_atan_log(const double* atan_tab, const double* log_tab, int which, const double* x, double* out, size_t n, cudaStream_t stream);
// The same for the fp32 table of CURVIS_PRECISION_F32 (ops 15 / 16; x is rounded to float first).  render_f32.cu.
cudaError_t launch_debug_shape32(const float4* tab, int which, const double* x, double* out, size_t n, cudaStream_t stream);

// FMA-only micro-kernels used as the measured compute-roofline denominator (peak_kernels.cu).
cudaError_t measure_fma_peak(int sm_count, cudaStream_t stream, double* fp64_tflops, double* fp32_tflops);

}  // namespace curvis
