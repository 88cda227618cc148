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
#include "fast_f64.cuh"
#include "launch.h"

namespace curvis {

constexpr int kBlock = 128;
constexpr unsigned kFull = 0xffffffffu;

// TUNED = true: euler_step_tuned (unguarded IEEE sequences behind one merged range check);
// TUNED = false: euler_step (plain operators) — kept as the A/B baseline and safety net.
template <class Shape, class Trig, bool TUNED>
__global__ void __launch_bounds__(kBlock) render_rows_f64(const __grid_constant__ FrameParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    const unsigned long long launch_rays = tile_rays * (p.n_frames ? p.n_frames : 1u);

    Ray q;
    bool live = false;      // this lane is integrating a ray
    bool pending = false;   // this lane finished a ray whose epilogue has not run yet
    bool drained = false;   // the queue is empty (warp-uniform)
    bool ray_safe = false;  // per-ray part of the tuned step's operand check
    const bool frame_safe = Shape::params_safe(p);
    int side = 0;
    uint32_t steps = 0;
    unsigned long long ray = 0;

    RayTally tally;

    for (;;) {
        // ---- epilogue of the lanes that finished during the last window
        if (pending) {
            const RayDiag nodiag = {__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll)};
            finish_ray<Shape, Trig, false>(p, q, side, steps, ray, tally, nodiag, 0.0);
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
                    if (idx < launch_rays) {
                        ray = idx;
                        new_photon_for_ray(p, idx, tile_rays, q);
                        ray_safe = frame_safe && ray_operands_safe(q);
                        steps = 0;
                        side = 0;
                        if (p.max_iterations == 0) pending = true;  // loop of systems.rs:126 runs zero times
                        else live = true;
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFull, live || pending) == 0u) break;
        }

        // ---- WINDOW Euler steps (escape_photon's loop body, systems.rs:126-135)
#pragma unroll 1
        for (uint32_t k = 0; k < p.window; ++k) {
            if (live) {
                if (TUNED) euler_step_tuned<Shape, Trig>(p, q, ray_safe);
                else euler_step<Shape, Trig>(p, q);
                ++steps;
                if (q.l > p.max_radius) { side = 1; live = false; pending = true; }          // :129-131
                else if (q.l < -p.max_radius) { side = -1; live = false; pending = true; }   // :132-134
                else if (steps >= p.max_iterations) { side = 0; live = false; pending = true; }  // :137
            }
        }
    }

    flush_tally(p, tally, lane);   // per-warp reduction of the counters, one atomic per counter per warp
}

// ---------------------------------------------------------------- default kernel (variant 3)
// Same execution model as render_rows_f64, leaner loop state: a down-counter instead of
// steps/max compare, the escape side decided in the epilogue from the final l, and the
// per-step escape test gated by an integer compare of |l|'s high word against the radius's
// (the fp64 compares only run within 2^-20 of the radius, or for NaN).
//   INTEG  0 Euler (the reference), 1 RK4, 2 Euler with the pole-adaptive step (extensions)
//   TRACK  the trajectory diagnostics of curvis_ray_record (launches that write records)
//   SHARED the step's six reciprocals from two seeds + one correction each (kernel_variant 4, default) or one full
//          correctly-rounded sequence per quotient (kernel_variant 3); same roundings, geodesic_f64.cuh: rhs_lean
// List mode (p.ray_list != nullptr): the launch re-integrates the rays CURVIS_PRECISION_F64_FAST left in its guard
// band — ray i of the launch is ray ray_list[i] of the tile, *ray_list_count of them.
// AHEAD: the step loop in latency form (geodesic_f64.cuh: euler_steps_ahead).  Built for list mode, where a few hundred rays leave the
// launch bound by one ray's dependent chain; its merged loop is also shorter (117 against 131 instructions per Ellis step), so whole
// frames run it too (kernel_variant 5, the default).  Same arithmetic.
// LONGFIRST: the longest-first refill of render_f64_fast.cu — a pre-pass has listed the rays predicted to take 10^4 steps, and the
// warps in the hardware slots the schedulers favour claim them first (this kernel's five warps per scheduler get 2.02 / 1.53 /
// 0.92 / 0.39 / 0.14 of the mean share in %warpid order: a 20,000-step ray needs 5 ms in the first resident CTA of its SM and
// 70 ms in the fifth).
template <class Shape, int INTEG, bool TRACK, bool SHARED, bool AHEAD = false, bool LONGFIRST = false>
__global__ void __launch_bounds__(kBlock) render_rows_f64_lean(const __grid_constant__ FrameParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width;
    // (list mode: the count includes rays a full list turned away — those were re-integrated in line by the fast kernel)
    const unsigned long long launch_rays = p.ray_list ? min(*p.ray_list_count, p.redo_capacity) : tile_rays * (p.n_frames ? p.n_frames : 1u);
    unsigned long long* const queue = p.ray_list ? &p.counters->redo_next : &p.counters->next_ray;
    const double R = p.max_radius;
    // |l| > R needs abs_hi(l) >= hi(R) when R >= 0; for negative or NaN R the gate is open (computed on the host)
    const unsigned gate = p.gate_hi;
    const bool frame_safe = Shape::params_safe(p);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    Ray q;
    int state = 0;            // 0 idle, 1 integrating, 2 finished (epilogue pending)
    bool drained = false;     // the queue is empty (warp-uniform)
    bool ray_safe = false;
    uint32_t remaining = 0;   // steps left before NotEscaped
    unsigned long long ray = 0;
    RayDiag diag = {qnan, qnan};
    TrigPins pins;            // three constants of the step's sincos, pinned in registers (trig_f64.cuh)
    pins.load(AHEAD && Shape::kAheadPinHalf);

    RayTally tally;
    unsigned long long n_long = 0;
    bool favoured = false;
    if (LONGFIRST) {
        n_long = p.counters->n_long;
        if (n_long > p.long_capacity) n_long = 0;      // a list that overflowed is ignored
        unsigned warpid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
        favoured = warpid < p.favoured_slots;
    }
    bool long_drained = (n_long == 0);

    for (;;) {
        if (state == 2) {
            const int side = (q.l > R) ? 1 : ((q.l < -R) ? -1 : 0);   // systems.rs:129-134 on the final state
            if (TRACK) diag.min_abs_sin = fmin(diag.min_abs_sin, fabs(TrigFast::sin(q.th)));   // the final direction reads it too
            finish_ray<Shape, TrigFast, false>(p, q, side, p.max_iterations - remaining, ray, tally, diag, 0.0);
            state = 0;
        }

        unsigned idle = __ballot_sync(kFull, state == 0);
        if (idle) {
            if (LONGFIRST) {
                // favoured slots: the list first, then the index walk (skipping the listed rays); the others: the walk, then what is left
                while (idle && !(drained && long_drained)) {
                    const bool from_list = !long_drained && (favoured || drained);
                    unsigned long long* const qu = from_list ? &p.counters->long_next : queue;
                    const unsigned long long qu_end = from_list ? n_long : launch_rays;
                    const int leader = __ffs(idle) - 1;
                    unsigned long long base = 0;
                    if ((int)lane == leader) base = atomicAdd(qu, (unsigned long long)__popc(idle));
                    base = __shfl_sync(kFull, base, leader);
                    if (state == 0) {
                        const unsigned long long ticket = base + (unsigned long long)__popc(idle & lt_mask);
                        if (ticket < qu_end) {
                            unsigned long long idx = ticket;
                            bool take = true;
                            if (from_list) idx = p.long_list[ticket];
                            else if (n_long) take = !ray_predicted_long(p, idx, tile_rays, kLongRaySin2);
                            if (take) {
                                ray = idx;
                                new_photon_for_ray(p, ray, tile_rays, q);
                                ray_safe = frame_safe && ray_operands_safe(q);
                                remaining = p.max_iterations;
                                state = (remaining == 0) ? 2 : 1;
                            }
                        }
                    }
                    if (base + (unsigned long long)__popc(idle) >= qu_end) {
                        if (from_list) long_drained = true;
                        else drained = true;
                    }
                    idle = __ballot_sync(kFull, state == 0);
                }
            } else if (!drained) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(queue, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFull, base, leader);
                if (state == 0) {
                    const unsigned long long idx = base + (unsigned long long)__popc(idle & lt_mask);
                    if (idx < launch_rays) {
                        // (list mode: the fast kernel appended the rays as they finished.  Where it claimed its predicted stragglers first,
                        // the longest rays are near the list's head; where it walked the indices they are its last entries, and the list is
                        // walked from its end — with guard = 2, when the list outnumbers the launch's lanes, they must start first)
                        ray = p.ray_list ? p.ray_list[p.list_from_end ? launch_rays - 1ull - idx : idx] : idx;
                        new_photon_for_ray(p, ray, tile_rays, q);
                        ray_safe = frame_safe && ray_operands_safe(q);
                        remaining = p.max_iterations;
                        if (TRACK) { diag.min_abs_sin = __longlong_as_double(0x7ff0000000000000ll); diag.stiffness = 0.0; }
                        state = (remaining == 0) ? 2 : 1;   // the loop of systems.rs:126 may run zero times
                    }
                }
                if (base + (unsigned long long)__popc(idle) >= launch_rays) drained = true;
            }
            if (__ballot_sync(kFull, state != 0) == 0u) break;
        }

        // ---- up to `window` steps (escape_photon's loop body, systems.rs:126-135).  Every lane runs its own loop; the only
        // per-step test is one integer compare of |l|'s high word against the radius's (host-computed: p.gate_hi) — the fp64
        // escape test, five DSETP, runs after the loop, for the step that reached the gate.  Lanes reconverge at the end.
        if (state == 1) {
            const uint32_t n = min(p.window, remaining);
            uint32_t k = 0;
            bool near = false;
            if (AHEAD) k = euler_steps_ahead<Shape>(p, pins, ray_safe, q, n, gate, near);
            else {
#pragma unroll 1
                do {
                    if (INTEG == 1) rk4_step_lean<Shape>(p, q, ray_safe);
                    else if (INTEG == 2) euler_step_adaptive<Shape, TRACK>(p, q, ray_safe, &diag);
                    else euler_step_lean<Shape, TRACK, SHARED>(p, q, ray_safe, &diag, &pins);
                    ++k;
                    if (abs_hi(q.l) >= gate) { near = true; break; }     // |l| >= R (1 - 2^-20), or NaN
                } while (k < n);
            }
            remaining -= k;
            bool done = (remaining == 0);                                       // systems.rs:137
            if (near) {
                done = done || (q.l > R) || (q.l < -R);                         // :129-134
                // A NaN l never compares true and never recovers (l += NaN): the reference would
                // spin through all remaining iterations and return NotEscaped.  Same result, same
                // step count, without the spinning.
                if (q.l != q.l) { remaining = 0; done = true; }
            }
            if (done) state = 2;
        }
        __syncwarp();
    }

    flush_tally(p, tally, lane);
}

template <class Kernel>
static cudaError_t launch_persistent(Kernel kernel, int& blocks_per_sm_auto, const FrameParams& p, int sm_count,
                                     int blocks_per_sm_override, cudaStream_t stream) {
    if (blocks_per_sm_auto == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_auto, kernel, kBlock, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm_auto < 1) blocks_per_sm_auto = 1;
    }
    int blocks_per_sm = blocks_per_sm_auto;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < blocks_per_sm) blocks_per_sm = blocks_per_sm_override;
    const unsigned long long tile_rays = (unsigned long long)(p.row_end - p.row_begin) * p.width * (p.n_frames ? p.n_frames : 1u);
    unsigned long long want = (tile_rays + kBlock - 1) / kBlock;
    unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)blocks_per_sm;
    if (p.ray_list) want = cap;   // list mode: the count lives on the device; surplus CTAs leave after one atomic
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    kernel<<<grid, kBlock, 0, stream>>>(p);
    return cudaGetLastError();
}

template <class Shape, int INTEG, bool TRACK, bool SHARED, bool AHEAD = false, bool LONGFIRST = false>
static cudaError_t launch_lean_one(const FrameParams& p, int sm_count, int blocks_per_sm_override, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;   // per instantiation
    return launch_persistent(render_rows_f64_lean<Shape, INTEG, TRACK, SHARED, AHEAD, LONGFIRST>, blocks_per_sm_auto, p, sm_count, blocks_per_sm_override, stream);
}

// The default kernel's launches that run the longest-first pre-pass: plain Euler tiles (no records, not the re-integration list) of
// at least 2^15 rays and at most 64 rays per lane — one rank's share of a frame split over several GPUs (rows 7, 15, ... of a 4K
// frame: 11.2 .. 12.3 ms in index order, 10.97 +- 0.02 with the list).  A whole frame on one GPU does not need it: its stragglers end
// before the kernel does even in the slowest slot, and the instantiation with the list handling is 1-2 % slower (4K Ellis 85.5
// against 83.9 ms).
static bool lean_longest_first(const FrameParams& p, const LaunchTuning& t, int sm_count) {
    return t.kernel_variant >= 5 && p.integrator == CURVIS_INTEGRATOR_EULER && !p.records && !p.ray_list && !p.ray_dirs &&
           longest_first_prepass_wanted(p, t.longest_first, sm_count, false);
}

bool render_f64_has_prepass(const FrameParams& p, const LaunchTuning& t, int sm_count) { return lean_longest_first(p, t, sm_count); }

template <class Shape>
static cudaError_t launch_lean(const FrameParams& p, int sm_count, int blocks_per_sm_override, bool shared, bool ahead, cudaStream_t stream) {
    const bool track = p.records != nullptr;
    if (ahead && p.integrator == CURVIS_INTEGRATOR_EULER && !track)
        return launch_lean_one<Shape, 0, false, true, true>(p, sm_count, blocks_per_sm_override, stream);   // latency form (re-integration list; whole frames)
    if (p.integrator == CURVIS_INTEGRATOR_RK4)
        return launch_lean_one<Shape, 1, false, true>(p, sm_count, blocks_per_sm_override, stream);
    if (p.integrator == CURVIS_INTEGRATOR_EULER_ADAPTIVE)
        return track ? launch_lean_one<Shape, 2, true, true>(p, sm_count, blocks_per_sm_override, stream)
                     : launch_lean_one<Shape, 2, false, true>(p, sm_count, blocks_per_sm_override, stream);
    if (track) return launch_lean_one<Shape, 0, true, true>(p, sm_count, blocks_per_sm_override, stream);
    return shared ? launch_lean_one<Shape, 0, false, true>(p, sm_count, blocks_per_sm_override, stream)
                  : launch_lean_one<Shape, 0, false, false>(p, sm_count, blocks_per_sm_override, stream);
}

template <class Shape, class Trig, bool TUNED>
static cudaError_t launch_one(const FrameParams& p, int sm_count, int blocks_per_sm_override, cudaStream_t stream) {
    static int blocks_per_sm_auto = 0;  // per instantiation; same for every sm_100 device
    return launch_persistent(render_rows_f64<Shape, Trig, TUNED>, blocks_per_sm_auto, p, sm_count, blocks_per_sm_override, stream);
}

template <class Shape>
static cudaError_t launch_variant(const FrameParams& p, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    if (p.integrator != CURVIS_INTEGRATOR_EULER || p.records || p.ray_list)
        return launch_lean<Shape>(p, sm_count, t.blocks_per_sm, true, p.ray_list && t.redo_ahead != 0, stream);   // extensions, diagnostics, list mode: lean kernel only
    switch (t.kernel_variant) {
    case 0: return launch_one<Shape, TrigCuda, false>(p, sm_count, t.blocks_per_sm, stream);   // round-1 v0
    case 1: return launch_one<Shape, TrigCuda, true>(p, sm_count, t.blocks_per_sm, stream);
    case 2: return launch_one<Shape, TrigFast, true>(p, sm_count, t.blocks_per_sm, stream);
    case 3: return launch_lean<Shape>(p, sm_count, t.blocks_per_sm, false, false, stream);    // one full division sequence per quotient
    case 5:                                                                                   // 4 with the step loop in latency form
        if (lean_longest_first(p, t, sm_count)) {
            cudaError_t e = launch_collect_long_rays(p, sm_count, stream);
            if (e != cudaSuccess) return e;
            return launch_lean_one<Shape, 0, false, true, true, true>(p, sm_count, t.blocks_per_sm, stream);
        }
        return launch_lean<Shape>(p, sm_count, t.blocks_per_sm, true, true, stream);
    default: return launch_lean<Shape>(p, sm_count, t.blocks_per_sm, true, false, stream);     // 4: shared reciprocals, one step per trip of the loop
    }
}

cudaError_t launch_render_f64(const FrameParams& p, int metric_kind, const LaunchTuning& t, int sm_count, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: return launch_variant<ShapeEllis>(p, t, sm_count, stream);
    case CURVIS_METRIC_INTERSTELLAR: return launch_variant<ShapeInterstellar>(p, t, sm_count, stream);
    case CURVIS_METRIC_FLAT: return launch_variant<ShapeFlat>(p, t, sm_count, stream);
    default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------- op-level test hook
__global__ void debug_eval_kernel(int op, const double* a, const double* b, double* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = a[i], y = b ? b[i] : 0.0;
    double r = 0.0, s, c;
    switch (op) {
    case 0: r = rcp_rn_unguarded(x); break;
    case 1: r = div_rn_unguarded(x, y); break;
    case 2: r = sqrt_rn_unguarded(x); break;
    case 3: sincos_fast(x, s, c); r = s; break;
    case 4: sincos_fast(x, s, c); r = c; break;
    case 5: TrigFast::sincos(x, s, c); r = s; break;
    case 6: TrigFast::sincos(x, s, c); r = c; break;
    case 7: r = x / y; break;
    case 8: r = sqrt(x); break;
    case 9: r = 1.0 / x; break;
    case 10: r = rcp_1ulp(x); break;
    case 11: case 12: { TrigRegs tr; tr.load(); sin2_sincos(tr, x, s, c); r = (op == 11) ? s : c; break; }
    default: break;
    }
    out[i] = r;
}

// Test hook (curvis_debug_rhs_check): the right-hand side of kernel_variant 4 (shared reciprocals, one correction step per
// quotient) against the plain operators on pseudo-random photon states; counts the outputs that differ in any bit.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long& x) {
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double unit_double(unsigned long long& x) { return (double)(splitmix64(x) >> 11) * (1.0 / 9007199254740992.0); }

template <class Shape>
__global__ void debug_rhs_check_kernel(const FrameParams p, unsigned long long seed, unsigned long long n, unsigned long long* mismatches) {
    unsigned long long bad[4] = {0, 0, 0, 0};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long x = seed + i * 0x632be59bd9b4e019ull;
        const int mode = (int)(i & 3);
        double l = (unit_double(x) * 2.0 - 1.0) * (mode == 0 ? 100.0 : mode == 1 ? 5.0 : mode == 2 ? 0.3 : 30.0);
        double th = unit_double(x) * CURVIS_PI;
        if (mode == 2) th = unit_double(x) * 0.05 + 1e-4;
        if (mode == 3) th = CURVIS_PI - unit_double(x) * 0.01;
        if (Shape::kind == CURVIS_METRIC_FLAT) l = fabs(l) + 1e-3;
        const double pth = (unit_double(x) * 2.0 - 1.0) * 6.0;
        const double pph = (unit_double(x) * 2.0 - 1.0) * 6.0 * sin(th);
        const double pph2 = pph * pph;
        double a[4], b[4], s;
        rhs_lean<Shape, true>(p, true, l, th, pth, pph, pph2, a[0], a[1], a[2], a[3], s);
        rhs_lean<Shape, true>(p, false, l, th, pth, pph, pph2, b[0], b[1], b[2], b[3], s);   // ray_safe = false: plain operators
        for (int k = 0; k < 4; ++k) bad[k] += (__double_as_longlong(a[k]) != __double_as_longlong(b[k])) ? 1ull : 0ull;
    }
    for (int k = 0; k < 4; ++k)
        if (bad[k]) atomicAdd(&mismatches[k], bad[k]);
}

cudaError_t launch_debug_rhs_check(const FrameParams& p, int metric_kind, unsigned long long seed, unsigned long long n,
                                   unsigned long long* d_mismatches, cudaStream_t stream) {
    switch (metric_kind) {
    case CURVIS_METRIC_ELLIS: debug_rhs_check_kernel<ShapeEllis><<<148 * 8, 256, 0, stream>>>(p, seed, n, d_mismatches); break;
    case CURVIS_METRIC_INTERSTELLAR: debug_rhs_check_kernel<ShapeInterstellar><<<148 * 8, 256, 0, stream>>>(p, seed, n, d_mismatches); break;
    case CURVIS_METRIC_FLAT: debug_rhs_check_kernel<ShapeFlat><<<148 * 8, 256, 0, stream>>>(p, seed, n, d_mismatches); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void texels_to_float4_kernel(const uint32_t* texels, float4* out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t t = texels[i];
        out[i] = make_float4((float)(t & 0xffu), (float)((t >> 8) & 0xffu), (float)((t >> 16) & 0xffu), (float)(t >> 24));
    }
}

cudaError_t launch_texels_to_float4(const uint32_t* texels, float4* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    texels_to_float4_kernel<<<148 * 8, 256, 0, stream>>>(texels, out, n);
    return cudaGetLastError();
}

__global__ void debug_bilinear_kernel(const Background bg, const double* fx, const double* fy, float4* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bilinear_tap(bg, fx[i], fy[i]);
}

cudaError_t launch_debug_bilinear(const Background& bg, const double* fx, const double* fy, float4* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    debug_bilinear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(bg, fx, fy, out, n);
    return cudaGetLastError();
}

__global__ void debug_atan_log_kernel(const double* atan_tab, const double* log_tab, int which, const double* x, double* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FrameParams p;
    p.atan_tab = atan_tab; p.log_tab = log_tab;
    double at, lg;
    ShapeInterstellar::atan_log(p, x[i], at, lg);
    out[i] = which ? lg : at;
}

cudaError_t launch_debug_atan_log(const double* atan_tab, const double* log_tab, int which, const double* x, double* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    debug_atan_log_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(atan_tab, log_tab, which, x, out, n);
    return cudaGetLastError();
}

cudaError_t launch_debug_eval(int op, const double* a, const double* b, double* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    debug_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(op, a, b, out, n);
    return cudaGetLastError();
}

}  // namespace curvis
