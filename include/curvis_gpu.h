/*
 * curvis_gpu.h — C ABI of libcurvis_b200.so, the B200 (sm_100a) replacement for the
 * per-pixel null-geodesic renderer of fragarriss/CurVis.
 *
 * The reference has no FFI seam today (it is one single-threaded Rust crate).  The
 * narrowest seam on the hot path is
 *
 *     RelativisticSystem::render_image(&self, max_iterations: u32, max_radius: f64,
 *                                      delta: f64) -> image::DynamicImage
 *                                                          (src/systems.rs:307-330)
 *
 * with  self = { metric, background_positive, background_negative, camera }
 *                                                          (src/systems.rs:68-73).
 *
 * Everything render_image reads crosses this boundary as plain-old-data (doubles,
 * fixed-width integers, raw pointers + sizes); every struct below can be mirrored by a
 * Rust `#[repr(C)]` struct field for field (INTEGRATION.md shows the binding).
 * No torch / C++ types appear in any signature.  No entry point aborts or unwinds:
 * each returns a curvis_status and leaves a message for curvis_last_error().
 *
 * Threading: one context is used by one host thread at a time; distinct contexts may
 * be used concurrently.  A context keeps no caller pointer after a call returns.
 */
#ifndef CURVIS_GPU_H
#define CURVIS_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: curvis_sim.integrator;  3: CURVIS_PRECISION_F64_FAST, curvis_host_register/unregister, curvis_peer_buffer_*,
 *    curvis_render_frames_peers, curvis_debug_shape_table_host (additions only: struct layouts are those of version 2)
 * 4: curvis_sim grows (frame, coordinates, step_tolerance; 40 -> 56 bytes), curvis_ray_record grows (min_abs_sin_theta,
 *    stiffness; 64 -> 80 bytes), curvis_stats.n_big_theta becomes n_reintegrated, CURVIS_PRECISION_F64_FAST renders the
 *    same integers as CURVIS_PRECISION_F64 (guard band + re-integration), CURVIS_INTEGRATOR_EULER_ADAPTIVE,
 *    CURVIS_FRAME_WORLD, CURVIS_COORDINATES_CARTESIAN, curvis_render_video */
#define CURVIS_ABI_VERSION 4

/* ---- status codes ------------------------------------------------------------------
 * The reference panics (src/systems.rs:122-124, src/algebra.rs:19-21, src/cameras.rs:
 * 94-105, src/metrics.rs:407-409 / 446-456, image index out of bounds) or bubbles a
 * Result<(), String> (src/rendering.rs:85).  Across the C ABI each becomes a code. */
typedef enum curvis_status {
    CURVIS_OK = 0,
    CURVIS_ERR_INVALID_ARGUMENT = 1,      /* null pointer, zero resolution, bad enum, bad row range */
    CURVIS_ERR_CAMERA_OUTSIDE_RADIUS = 2, /* |l_camera| > max_radius   (systems.rs:122-124 panic)   */
    CURVIS_ERR_PARALLEL_VECTORS = 3,      /* forward x up == 0         (algebra.rs:19-21 panic)     */
    CURVIS_ERR_INVALID_METRIC = 4,        /* rho/m/a <= 0              (metrics.rs:407-409,446-456) */
    CURVIS_ERR_NO_BACKGROUND = 5,         /* render before both backgrounds were set                */
    CURVIS_ERR_CUDA = 6,                  /* a CUDA runtime call failed; see curvis_last_error      */
    CURVIS_ERR_OUT_OF_MEMORY = 7,
    CURVIS_ERR_NO_DEVICE = 8,             /* no sm_100 device visible: there is NO CPU fallback     */
    CURVIS_ERR_UNSUPPORTED = 9            /* e.g. a sim option this build does not implement        */
} curvis_status;

/* ---- metric: closed enum + parameters ----------------------------------------------
 * Replaces the generic `M: DiagonalSphericalMetric` (src/metrics.rs:40-44).  Kernels are
 * template-instantiated per kind; arbitrary user metrics cannot cross the ABI. */
typedef enum curvis_metric_kind {
    CURVIS_METRIC_ELLIS = 0,        /* EllisMetric{rho}             src/metrics.rs:399-421 */
    CURVIS_METRIC_INTERSTELLAR = 1, /* InterstellarMetric{m,a,rho}  src/metrics.rs:431-487 */
    CURVIS_METRIC_FLAT = 2          /* FlatSphericalMetric{}        src/metrics.rs:492-505 */
} curvis_metric_kind;

typedef struct curvis_metric {
    int32_t kind;  /* curvis_metric_kind */
    int32_t _pad;
    double rho;    /* Ellis, Interstellar */
    double m;      /* Interstellar        */
    double a;      /* Interstellar        */
} curvis_metric;

/* ---- camera --------------------------------------------------------------------------
 * The fields Camera (src/cameras.rs:30-43) holds after Camera::new ran.  Fill it with
 * curvis_camera_init() (which restates Camera::new + Orientation::new) or, from a Rust
 * host, copy the values nalgebra already computed. */
typedef struct curvis_camera {
    double position[4];      /* (t, l, theta, phi), contravariant   cameras.rs:31           */
    double cam_to_world[9];  /* row-major 3x3                        cameras.rs:42           */
    double focal_length;     /*                                      cameras.rs:35           */
    double sensor_width;     /*                                      cameras.rs:36, :107-110 */
    double sensor_height;    /*                                      cameras.rs:37           */
    uint32_t resolution_width;
    uint32_t resolution_height;
} curvis_camera;

/* ---- simulation settings -------------------------------------------------------------
 * The three arguments of render_image (systems.rs:307-312) plus opt-in extensions.  All
 * extension fields are 0 in parity mode. */
typedef enum curvis_precision {
    CURVIS_PRECISION_F64 = 0, /* reference arithmetic: IEEE double, reference operation order */
    CURVIS_PRECISION_F32 = 1, /* extension: fp32 state (not bit-comparable)    */
    CURVIS_PRECISION_F64_FAST = 2 /* fp64 with the right-hand side regrouped around ONE reciprocal per step and fused
                                     multiply-adds (every operation <= 1 ulp, rounding points differ from the reference's,
                                     state agrees to ~1e-13 relative) as a PREDICTOR: a ray whose escape step or texel lies
                                     within a guard band of a decision boundary is re-integrated with the
                                     CURVIS_PRECISION_F64 arithmetic in a second launch over the compacted list, so RGB8 / side
                                     / step count / texel equal CURVIS_PRECISION_F64's on every ray with stiffness < 1 (97.5 %
                                     of the default frame; curvis_stats.n_reintegrated).  Rays with stiffness >= 1 ("kicked":
                                     their end state amplifies a last-bit change up to 1e10-fold) are counted in n_kicked and,
                                     with the context option "guard" = 2, re-integrated as well (then every ray equals
                                     CURVIS_PRECISION_F64's; +20 % time); with the default "guard" = 1 they are kept as
                                     integrated (measured: 1 differing pixel in 21 M).  DESIGN.md section 4 */
} curvis_precision;

typedef enum curvis_sampling {
    CURVIS_SAMPLING_NEAREST = 0, /* images.rs:115-121: truncating nearest texel, u8 copy */
    CURVIS_SAMPLING_BILINEAR = 1 /* extension: fp32 2x2 tap on float4 texels (wrap in x, clamp in y), rounded to u8 */
} curvis_sampling;

typedef enum curvis_integrator {
    CURVIS_INTEGRATOR_EULER = 0, /* update_relativistic_object, metrics.rs:283-297: explicit Euler (the reference's only stepper) */
    CURVIS_INTEGRATOR_RK4 = 1,   /* extension: classical 4th-order Runge-Kutta on the same right-hand side; one
                                    "iteration" = one RK4 step, escape test after every step */
    CURVIS_INTEGRATOR_EULER_ADAPTIVE = 2 /* extension: the same explicit Euler step with a per-step size
                                    h = delta * min(1, step_tolerance / kappa), kappa = delta^2 p_phi^2 / (r^2 sin^2 theta)^2 the
                                    stiffness of the theta equation at the current state (the quantity that makes fixed-step
                                    rays near a coordinate pole "chaotic"); far from the poles h = delta and the
                                    step is the reference's, bit for bit.  One "iteration" = one step of whatever size. */
} curvis_integrator;

typedef enum curvis_frame {
    CURVIS_FRAME_LOCAL = 0,  /* render_image as written: the escaped photon's LOCAL tangent-frame direction indexes the
                                background, phi component scaled by frame_field_22 (systems.rs:540-561, metrics.rs:347)       */
    CURVIS_FRAME_WORLD = 1,  /* extension: escaped_photon_to_world_direction (systems.rs:144-187) applied to the escaped
                                photon — tangent-frame direction with the phi component scaled by frame_field_33 (the fix of
                                metrics.rs:347), rotated by rotation_from_two_vectors(x, vector3_from_theta_phi(theta, phi))   */
    CURVIS_FRAME_WORLD_QUIRK = 2 /* the same rotation on the direction exactly as metrics.rs:339-349 returns it (frame_field_22
                                twice) — what compute_escape_angle evaluates for the table of render_image_efficient           */
} curvis_frame;

typedef enum curvis_coordinates {
    CURVIS_COORDINATES_SPHERICAL = 0, /* the reference's state (l, theta, phi, p_l, p_theta, p_phi)                        */
    CURVIS_COORDINATES_CARTESIAN = 1  /* extension ("pole-safe"): the angular part of the state is the unit position vector n
                                         and the conserved angular-momentum vector; no 1/sin(theta) anywhere (DESIGN.md
                                         section 7).  CURVIS_PRECISION_F64: the oracle's operation order;
                                         CURVIS_PRECISION_F64_FAST: the same scheme regrouped (17 fp64 instructions per
                                         Ellis step, the fastest integrator of the library; no guard band)              */
} curvis_coordinates;

typedef struct curvis_sim {
    uint32_t max_iterations; /* systems.rs:309 */
    int32_t frame;           /* curvis_frame */
    double max_radius;       /* systems.rs:310 */
    double delta;            /* systems.rs:311 */
    int32_t precision;       /* curvis_precision  */
    int32_t sampling;        /* curvis_sampling   */
    int32_t integrator;      /* curvis_integrator */
    int32_t coordinates;     /* curvis_coordinates */
    double step_tolerance;   /* CURVIS_INTEGRATOR_EULER_ADAPTIVE: largest stiffness a full step may see (e.g. 1e-3); else 0 */
    double _reserved;        /* 0 */
} curvis_sim;

/* ---- per-frame counters (the reference has none) ----------------- */
typedef struct curvis_stats {
    uint64_t total_steps;   /* sum over rays of Euler steps executed (early exit counted) */
    uint64_t n_rays;
    uint64_t n_positive;    /* PhotonEscape::PositiveSpace  systems.rs:129-131 */
    uint64_t n_negative;    /* PhotonEscape::NegativeSpace  systems.rs:132-134 */
    uint64_t n_not_escaped; /* PhotonEscape::NotEscaped     systems.rs:137     */
    uint64_t n_clamped;     /* texel index the reference would have panicked on (images.rs:107-111) */
    uint64_t n_reintegrated;/* CURVIS_PRECISION_F64_FAST: rays inside the guard band, re-integrated with the F64 arithmetic */
    uint64_t n_kicked;      /* CURVIS_PRECISION_F64_FAST: rays with stiffness >= 1 (some step advanced phi by a radian or more: explicit
                               Euler is no longer integrating their theta motion) */
    double kernel_ms;       /* device time of the render kernel(s), CUDA events, max over devices */
    double total_ms;        /* host wall time of the call, copies included */
} curvis_stats;

/* ---- per-ray record (debug / parity; optional) ---------------------------------------
 * Final photon state as escape_photon left it (systems.rs:115-139) and the texel chosen by
 * images.rs:115-121.  side: +1 PositiveSpace, -1 NegativeSpace, 0 NotEscaped. */
typedef struct curvis_ray_record {
    double l, theta, phi, p_l, p_theta, p_phi;
    uint32_t steps;
    int32_t side;
    uint32_t texel_x, texel_y;
    /* trajectory diagnostics (the regular / chaotic classifier reads them):
     *   min_abs_sin_theta  min |sin theta| over the states 0..steps of the ray (every state the right-hand side or the
     *                      final direction was evaluated at)
     *   stiffness          max over the steps of kappa = delta^2 p_phi^2 / (r^2 sin^2 theta)^2 = |d(delta dp_theta)/d theta| *
     *                      |d(delta dtheta)/d p_theta| up to a factor <= 3: explicit Euler is a faithful map of the theta
     *                      motion while kappa << 1 and amplifies perturbations once kappa >~ 1 (a "kicked" ray) */
    double min_abs_sin_theta;
    double stiffness;
} curvis_ray_record;

typedef struct curvis_ctx curvis_ctx;

/* ---- lifetime ------------------------------------------------------------------------ */

/* Opens `n_devices` CUDA devices (ordinals in `devices`; NULL/0 = every visible device).
 * A frame is row-tiled over the context's devices.  Fails with CURVIS_ERR_NO_DEVICE when no
 * CUDA device is present — the library never computes on the CPU. */
int curvis_ctx_create(const int* devices, int n_devices, curvis_ctx** out);
void curvis_ctx_destroy(curvis_ctx* ctx);

/* Message of the last failing call on `ctx` (or of the last failing ctx-less call when
 * ctx == NULL).  Valid until the next call on the same context/thread. */
const char* curvis_last_error(const curvis_ctx* ctx);
int curvis_abi_version(void);
int curvis_ctx_device_count(const curvis_ctx* ctx);

/* ---- host-side setup helpers (pure CPU, no device needed) ----------------------------- */

/* Orientation::new (src/algebra.rs:16-38) + rotation_matrix_from_forward_up_pairs
 * (:64-74): rotation taking (x, z) to (forward, orthogonalised up).  rot / inv_rot are
 * row-major 3x3; any output may be NULL. */
int curvis_orientation(const double forward[3], const double up[3],
                       double rot[9], double inv_rot[9], double up_orthogonal[3]);

/* Camera::new (src/cameras.rs:79-122). */
int curvis_camera_init(curvis_camera* cam, const double position[4],
                       const double forward[3], const double up[3],
                       double focal_length, double sensor_diagonal,
                       uint32_t resolution_width, uint32_t resolution_height);

/* EllisMetric::new / InterstellarMetric::new / FlatSphericalMetric::new parameter checks
 * (src/metrics.rs:404-414, :441-459, :496-498). */
int curvis_metric_validate(const curvis_metric* metric);

/* ---- scene --------------------------------------------------------------------------- */

/* SphericalImage::new (src/images.rs:71-89): side > 0 = background_positive, side < 0 =
 * background_negative (systems.rs:70-71).  `rgba8` is what DynamicImage::get_pixel would
 * return for every texel (images.rs:107-111), row-major, 4 bytes per texel; it is copied
 * to every device of the context.  `inv_rot` = the image orientation's inverse rotation
 * (images.rs:132-142), row-major; NULL = identity (images are always loaded with
 * forward/up = None, rendering.rs:36-39). */
int curvis_set_background(curvis_ctx* ctx, int side, const uint8_t* rgba8,
                          uint32_t width, uint32_t height, const double inv_rot[9]);

/* ---- fused render + all-gather over NVLink peer memory (one rank per GPU) -------------------
 * curvis_render_frames_device leaves this rank's row tiles on its own device and the host program
 * all-gathers them (NCCL).  The fused form needs no collective for the pixels: every rank owns one
 * device buffer holding the COMPLETE frames of a step (n_frames * W*H*3 bytes, frame-major), exports
 * it to its peers (CUDA IPC), and the render kernel of every rank stores each finished ray's RGB8
 * straight into all of them — its own and, over NVLink, its peers'.  After the launch a rank only
 * has to learn that its peers' kernels have ended (any barrier on the stream, e.g. a 4-byte
 * all-reduce) before it reads its buffer.
 *   curvis_peer_buffer_create  cudaMalloc on the context's first device + the 64-byte IPC handle to send to the peers
 *   curvis_peer_buffer_open    maps a peer's buffer from its handle (another process on the same node)
 *   curvis_peer_buffer_close   unmaps an opened buffer;  curvis_peer_buffer_destroy frees a created one
 *   curvis_render_frames_peers rows row_begin, row_begin + row_stride, ... (< row_end) of n_frames frames in ONE
 *                              launch, stored into d_frames[0..n_peers) (device pointers valid on this device:
 *                              own buffer and opened peers; n_peers <= CURVIS_MAX_PEERS).  row_stride = 1: a
 *                              contiguous tile; rank g of N with (row_begin, row_stride) = (g, N): interleaved rows,
 *                              which gives every rank the same mix of short and long rays (the pixels land in place
 *                              either way).  Asynchronous on `stream` unless stats.
 *   curvis_render_frames_peers_blocks  the same launch over BLOCKS of block_width consecutive pixels of a frame row
 *                              (block_width divides the width; blocks numbered row-major over the frame, W / block_width per
 *                              row): blocks block_begin, block_begin + block_stride, ... (< block_end <= H * W / block_width).
 *                              Rank g of N with (block_begin, block_stride) = (g, N) owns an N-th of EVERY row: the 10^4-step
 *                              rays of a frame sit in two or three rows (photons grazing the coordinate poles), and with whole
 *                              rows the ranks that hold those rows finish last (one 4K frame over 8 GPUs: 6.0 against 5.1 ms).
 *                              block_width = W is curvis_render_frames_peers. */
#define CURVIS_MAX_PEERS 8
#define CURVIS_IPC_HANDLE_BYTES 64
int curvis_peer_buffer_create(curvis_ctx* ctx, size_t bytes, void** d_ptr, uint8_t ipc_handle[CURVIS_IPC_HANDLE_BYTES]);
int curvis_peer_buffer_open(curvis_ctx* ctx, const uint8_t ipc_handle[CURVIS_IPC_HANDLE_BYTES], void** d_ptr);
int curvis_peer_buffer_close(curvis_ctx* ctx, void* d_ptr);
int curvis_peer_buffer_destroy(curvis_ctx* ctx, void* d_ptr);
int curvis_render_frames_peers(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* cameras, uint32_t n_frames,
                               const curvis_sim* sim, uint32_t row_begin, uint32_t row_end, uint32_t row_stride,
                               void* const* d_frames, uint32_t n_peers, void* stream, curvis_stats* stats);
int curvis_render_frames_peers_blocks(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* cameras, uint32_t n_frames,
                                      const curvis_sim* sim, uint32_t block_begin, uint32_t block_end, uint32_t block_stride,
                                      uint32_t block_width, void* const* d_frames, uint32_t n_peers, void* stream, curvis_stats* stats);

/* Optional: page-locks a caller-owned host buffer that will be passed as `out_rgb8` to
 * curvis_render_image / curvis_render_rows again and again (the frame buffer of a video loop; the
 * reference allocates a fresh DynamicImage per frame, systems.rs:314, a binding would keep one).
 * A frame whose destination lies inside a registered buffer is written there by the render kernel
 * itself (mapped memory; a tile of curvis_render_rows is DMA'd into it) — no device frame, no staging
 * copy on the host: 2.5 ms -> 0.04 ms of read-back per 4K frame.  The buffer must stay allocated
 * until curvis_host_unregister (or curvis_ctx_destroy, which unregisters what is left). */
int curvis_host_register(curvis_ctx* ctx, void* ptr, size_t bytes);
int curvis_host_unregister(curvis_ctx* ctx, void* ptr);

/* ---- the hot path -------------------------------------------------------------------- */

/* RelativisticSystem::render_image (src/systems.rs:307-330).  Renders the whole frame into the HOST buffer `out_rgb8`
 * (W*H*3 bytes, row-major: out[(y*W + x)*3 + c], the layout of DynamicImage::ImageRgb8).  With several devices in the
 * context the rows are interleaved over them (device g renders rows g, g + n, ...: equal work for every device; pixels are
 * independent, so there is no exchange — each device stores or copies its rows to their places in the frame).
 * `stats` may be NULL.
 *
 * One launch in flight per context: every launch of a context shares its per-device scratch (work-queue cursor, counters,
 * batched cameras, re-integration list, timing events).  Launches passed the SAME stream are ordered by the stream; a launch
 * on a different stream than the context's previous one first waits (cudaStreamWaitEvent) for that launch to finish.  Use
 * one context per concurrent pipeline. */
int curvis_render_image(curvis_ctx* ctx, const curvis_metric* metric,
                        const curvis_camera* camera, const curvis_sim* sim,
                        uint8_t* out_rgb8, curvis_stats* stats);

/* Rows [row_begin, row_end) of the same frame on the context's FIRST device, into host
 * buffers holding only those rows.  `records` (nullable) receives one curvis_ray_record per
 * pixel of the tile, row-major.  This is the unit a multi-process host (one rank per GPU)
 * calls before its own all-gather. */
int curvis_render_rows(curvis_ctx* ctx, const curvis_metric* metric,
                       const curvis_camera* camera, const curvis_sim* sim,
                       uint32_t row_begin, uint32_t row_end,
                       uint8_t* out_rgb8_rows, curvis_ray_record* records,
                       curvis_stats* stats);

/* Same tile, device-resident: `d_out_rgb8_rows` is a DEVICE pointer on the context's first
 * device ((row_end-row_begin)*W*3 bytes), `d_records` a nullable device pointer, `stream` a
 * cudaStream_t passed as void* (NULL = the legacy default stream).  The launch is
 * asynchronous on `stream`; when `stats` is non-NULL the call synchronises the stream and
 * fills it.  Used to keep the frame in HBM for an NCCL all-gather or a following kernel. */
int curvis_render_rows_device(curvis_ctx* ctx, const curvis_metric* metric,
                              const curvis_camera* camera, const curvis_sim* sim,
                              uint32_t row_begin, uint32_t row_end,
                              void* d_out_rgb8_rows, void* d_records,
                              void* stream, curvis_stats* stats);

/* The same tile as unrounded colours: 4 floats (R, G, B, A on the 0..255 scale) per pixel into a
 * HOST buffer.  With CURVIS_SAMPLING_NEAREST these are the texel's bytes as floats; with
 * CURVIS_SAMPLING_BILINEAR (extension) the fp32 2x2 tap before quantisation. */
int curvis_render_rows_rgba32f(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* camera,
                               const curvis_sim* sim, uint32_t row_begin, uint32_t row_end,
                               float* out_rgba32f_rows, curvis_stats* stats);

/* Batched form for video (VideoRenderingSystem::render, src/rendering.rs:291-316: one
 * render per frame after update_camera): rows [row_begin,row_end) of `n_frames` frames —
 * one camera per frame, same resolution, same metric/sim/backgrounds — in ONE launch over a
 * single work queue.  Output: n_frames tiles, frame-major ((row_end-row_begin)*W*3 bytes each)
 * at the device pointer `d_out_rgb8_tiles`.  Asynchronous on `stream` unless `stats` != NULL. */
int curvis_render_frames_device(curvis_ctx* ctx, const curvis_metric* metric,
                                const curvis_camera* cameras, uint32_t n_frames, const curvis_sim* sim,
                                uint32_t row_begin, uint32_t row_end,
                                void* d_out_rgb8_tiles, void* stream, curvis_stats* stats);

/* ---- the table-based renderer (what `curvis image` / `curvis video` run) -------------------
 * RelativisticSystem::render_image_efficient (src/systems.rs:333-527): an adaptive sampler
 * (src/sampling.rs) tabulates alpha -> (escape angle, escape space) from a few hundred
 * equatorial photon integrations, then every pixel interpolates the table and rotates the
 * camera-position direction.  Arguments mirror the reference's (systems.rs:333-342); the
 * photon integrations and the per-pixel pass run on the context's first device. */
typedef struct curvis_sampling_settings {
    uint32_t alphas_num;              /* systems.rs:338  (settings key sampling_initial_nums)          */
    uint32_t max_iterations_sampling; /* systems.rs:339  (main.rs:46-47 passes sampling_initial_nums)  */
    double threshold_1;               /* systems.rs:340  sampling_convergence_threshold_1               */
    double threshold_2;               /* systems.rs:341  sampling_convergence_threshold_2               */
} curvis_sampling_settings;

typedef struct curvis_efficient_info {
    uint32_t table_points;       /* points the sampler kept                              */
    uint32_t table_passes;       /* device launches of the sampler (look-ahead merges passes) */
    uint64_t table_evaluations;  /* photons integrated                                   */
    uint64_t table_steps;        /* Euler steps of those photons                         */
    double table_ms;             /* host wall time of the sampler (launches included)    */
    double pixels_ms;            /* device time of the per-pixel pass                    */
} curvis_efficient_info;

/* `dbg` (nullable, host, 3 doubles per pixel): alpha, interpolated escape angle, escape space.
 * stats->total_steps counts the table's Euler steps; n_positive/negative/not_escaped classify
 * PIXELS (not_escaped = black, including the reference's seam artefact, README.md:108).
 * sim->precision: CURVIS_PRECISION_F64 — the table equals the CPU's bit for bit; CURVIS_PRECISION_F64_FAST — the
 * table's photons are integrated by the regrouped fp64 kernel (the passes are latency-bound and its dependency
 * chain is shorter: 4K Ellis frame 6.0 -> 3.1 ms; same sampler decisions and frames on every tested scene);
 * CURVIS_PRECISION_F32 -> CURVIS_ERR_UNSUPPORTED.  A frame registered with curvis_host_register is DMA'd in place. */
int curvis_render_image_efficient(curvis_ctx* ctx, const curvis_metric* metric, const curvis_camera* camera,
                                  const curvis_sim* sim, const curvis_sampling_settings* sampling,
                                  uint8_t* out_rgb8, double* dbg, curvis_stats* stats, curvis_efficient_info* info);

/* ---- measurement helpers ------------------------------------------------------------- */

/* Number of render-kernel launches this process has issued so far (all contexts).  bench.py
 * reads it on both sides of the timed region to report `gpu_launches`. */
uint64_t curvis_kernel_launch_count(void);

/* Runs an FMA-only micro-kernel on the context's first device and returns the achieved
 * fp64 and fp32 FMA rates in TFLOP/s (2 flop per FMA).  bench.py uses it as the measured
 * compute-roofline denominator (MEASURED_PEAKS.json has no fp64/fp32 ALU entry). */
int curvis_measure_fma_peak(curvis_ctx* ctx, double* fp64_tflops, double* fp32_tflops);

/* ---- tuning and test hooks ------------------------------------------------------------ */

/* Tuning knobs of a context; results never depend on them (tests/test_gpu_parity.py).
 *   "kernel_variant": 0 plain `/`, sqrt and CUDA sincos; 1 unguarded IEEE sequences + CUDA
 *                     sincos; 2 unguarded IEEE sequences + in-kernel sincos; 3 the same
 *                     arithmetic in the lean loop (integer-pipe guards, gated escape test); 4 the lean loop
 *                     with the step's six reciprocals built from two seeds and one correction step each; 5 (default) 4 with
 *                     the next step's shape function and sincos carried across the loop's back edge (fewer instructions, a
 *                     shorter dependent chain; the form the re-integration launch of CURVIS_PRECISION_F64_FAST runs too)
 *   "redo_ahead":     1 (default) the re-integration launch runs that latency form; 0 the form of variant 4 (A/B)
 *   "guard":          CURVIS_PRECISION_F64_FAST: 1 (default) guard band + re-integration of the rays with stiffness < 1;
 *                     2 kicked rays (stiffness >= 1) re-integrated too; 0 the raw regrouped kernel (A/B)
 *   "guard_rel_e15":  the guard's relative budget in units of 1e-15 (default 1000000 = 1e-9)
 *   "fast_regs":      96 / 128: register budget of the fast kernel (5 / 4 resident CTAs per SM); 0 (default) = 96
 *   "redo_blocks_per_sm": resident CTAs per SM of the re-integration launch (default 2)
 *   "redo_capacity_limit": test knob — caps the re-integration list (0 = automatic: one slot per ray up to 2^24 rays); a ray
 *                     that finds the list full is re-integrated in line by the fast kernel (same result, slower)
 *   "favoured_slots": the warps in hardware slots %warpid < this claim the longest-first list first (default 8: a scheduler gives
 *                     its five warps 1.65 / 1.51 / 1.08 / 0.55 / 0.21 of the mean share in slot order); 0 none, 64 all
 *   "longest_first":  CURVIS_PRECISION_F64_FAST: a pre-pass kernel lists the rays predicted to be the 10^4-step stragglers (near-
 *                     critical photons grazing a coordinate pole) and the work queue hands them out first: 1 always, 0 never
 *                     (index order), 2 (default) in launches of at most 64 rays per lane of the grid — one frame split over
 *                     several GPUs — where one straggler's latency is comparable to the kernel's
 *   "blocks_per_sm":  resident CTAs per SM of the persistent grid (0 = occupancy maximum)
 *   "window":         Euler steps between two refill points of a warp (0 = default: 32; for
 *                     CURVIS_PRECISION_F64_FAST 32..128, growing with the expected ray length)
 *   "zero_copy":      curvis_render_image into a buffer registered with curvis_host_register: 1 (default)
 *                     = the kernel stores its pixels straight into the mapped host frame (no device
 *                     frame, no copy; 3 bytes per ray over PCIe do not slow the kernel); 0 = device
 *                     frame + one DMA
 *   "fast_variant":   CURVIS_PRECISION_F64_FAST only: 0 sin/cos from theta every step; 1 (default)
 *                     (sin theta, cos theta) carried along and rotated by the step's small dtheta,
 *                     re-derived from theta once per window (results agree to ~1e-13, same frames) */
int curvis_ctx_set_option(curvis_ctx* ctx, const char* key, int64_t value);

/* Op-level test hook: out[i] = op(a[i], b[i]) evaluated on the device (host pointers).
 *   0 rcp_rn_unguarded(a)  1 div_rn_unguarded(a,b)  2 sqrt_rn_unguarded(a)
 *   3 / 4 sin / cos of the in-kernel sincos fast path   5 / 6 the same with its large-argument fallback
 *   7 a/b   8 sqrt(a)   9 1/a   (the compiler's IEEE operators, for reference)
 *   10 the <= 1 ulp reciprocal of CURVIS_PRECISION_F64_FAST   11 / 12 its sin^2(a) / sin(a)cos(a)
 *   13 / 14 its Interstellar shape functions a atan a - ln(1 + a^2)/2 and (2/pi) atan a (table; 0 for a <= 0)
 *   15 / 16 the same two functions as CURVIS_PRECISION_F32 evaluates them (fp32 table; a is rounded to float)
 *   17 / 18 atan a and ln(1 + a^2) as the CURVIS_PRECISION_F64 Interstellar step evaluates them (tables inside a in [2^-10, 2^16),
 *           the CUDA library outside)   */
int curvis_debug_eval(curvis_ctx* ctx, int op, const double* a, const double* b, double* out, size_t n);

/* Test hooks of the per-metric Interstellar table CURVIS_PRECISION_F64_FAST reads (csrc/shape_table.h): with r = rho + m (x atan x
 * - ln(1 + x^2)/2) at x = 2 z/(pi m), y[i] = 1/r^2 and g[i] = (2/pi) atan x / r^3 for z[i] = |l| - a; every z below 2^-44 (zero
 * and negative included: the plateau |l| <= a) reads the constant row y = 1/rho^2, g = 0.  _host: the table built and evaluated
 * on the host with the kernel's arithmetic, no GPU needed (returns 1 when every z was below the table's end 2^14, else 0
 * and NaN entries); the other evaluates on the context's first device. */
int curvis_debug_inverse_table_host(double rho, double m, const double* z, double* y, double* g, size_t n);
int curvis_debug_inverse_shape(curvis_ctx* ctx, const curvis_metric* metric, const double* z, double* y, double* g, size_t n);

/* Diagnostics of the last launch that returned stats (first device): Euler steps executed per hardware warp slot (%warpid, 64
 * entries) and per SM (%smid, 192 entries).  Every warp of a persistent launch lives as long as the kernel, so the entries are
 * the shares of issue slots the warp schedulers handed out (tools/scheduler_shares.py). */
int curvis_debug_last_step_shares(curvis_ctx* ctx, uint64_t slot_steps[64], uint64_t sm_steps[192]);

/* Test hook of kernel_variant 4 (the default CURVIS_PRECISION_F64 step: the six reciprocals of metrics.rs:257-262 from two
 * MUFU seeds and one correction step each): evaluates the right-hand side of n_samples pseudo-random photon states both
 * ways — shared reciprocals vs. the plain IEEE operators — and returns in mismatches[0..3] how many of the outputs
 * dtheta, dphi, dp_l, dp_theta differ in any bit. */
int curvis_debug_rhs_check(curvis_ctx* ctx, const curvis_metric* metric, uint64_t n_samples, uint64_t seed, uint64_t mismatches[4]);

/* Test hook, host only (no GPU needed): the piecewise-polynomial table of the Interstellar shape
 * functions that CURVIS_PRECISION_F64_FAST uploads to every device (csrc/shape_table.h), evaluated on
 * the host with the kernel's arithmetic: f[i] = x atan x - ln(1 + x^2)/2, g[i] = (2/pi) atan x.  Returns 1
 * when every x[i] lay inside the table's range [2^-10, 2^16), else 0 (those entries are NaN). */
int curvis_debug_shape_table_host(const double* x, double* f, double* g, size_t n);

/* Test hook, host only: the atan (which = 0, x in [2^-10, 2^16)) and ln (which = 1, argument in [1, 2^33)) tables of the
 * CURVIS_PRECISION_F64 Interstellar step (csrc/shape_table.h), evaluated on the host with the kernel's arithmetic.  Returns 1
 * when every argument lay inside the table's range, else 0 (those entries are NaN).  On the device: curvis_debug_eval ops 17 / 18. */
int curvis_debug_fn_table_host(int which, const double* x, double* out, size_t n);

/* Test hook of CURVIS_SAMPLING_BILINEAR: the fp32 tap of background `side` at explicit continuous
 * texel coordinates (fx in [0,W), fy in [0,H]; texel centres at integer + 0.5; wrap in x, clamp in
 * y) — 4 floats per point into `out_rgba32f`.  Host pointers. */
int curvis_debug_bilinear(curvis_ctx* ctx, int side, const double* fx, const double* fy, float* out_rgba32f, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* CURVIS_GPU_H */
