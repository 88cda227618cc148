#!/usr/bin/env python
"""bench.py — ray-steps/s of the per-pixel null-geodesic renderer (reference
RelativisticSystem::render_image, src/systems.rs:307-330) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "ray-steps/sec and frames/sec at 3840x2160 Ellis"): the
Ellis rho=1 wormhole at 3840x2160 with the reference's default camera and simulation settings
(settings/defaults/*.toml: l=5, forward -x, diag 43, focal 15; escape_radius 100, max
iterations 40000, step 0.05), two synthetic decodable 8192x4096 RGBA8 backgrounds.

One "step" = one pass of the hot path over one batch: at N GPUs the batch is N frames of that
4K scene (frame f of a synthetic camera path), EACH frame row-tiled over the N ranks (rank g
renders rows [g*H/N, (g+1)*H/N)) and its tiles all-gathered over NCCL so every rank holds every
complete frame.  Per-GPU work is one frame's worth of rays whatever N is -> "scaling": "weak".
At N=1 there is no collective.

Printed (rank 0, ONE JSON line): see the task contract; `value` = ray-steps/s with the scene
resident in HBM (CUDA events, max over ranks); `e2e` = the same through the host-buffer C-ABI
call curvis_render_image / curvis_render_rows (per-frame parameters in, RGB8 frame out to
host); `roofline` against the live-measured fp64 FMA peak; `cpu_baseline` = the oracle port on
one host core (the reference is single-threaded, README.md:110) on a bounded row sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
BG_W, BG_H = 8192, 4096
BLOCK_WIDTH = 64          # --tiles blocks: pixels of a row per block (curvis_render_frames_peers_blocks)
FLOP_PER_STEP = {"ellis": 33, "interstellar": 43}   # SURVEY.md 8a canonical count after CSE
METRIC_NAME = "ray_steps_per_sec"
UNIT = "ray-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="peers", choices=["peers", "nccl"],
                    help="N > 1: peers (default) = the render kernel stores every pixel into the complete frames of all ranks over "
                         "NVLink (fused render + all-gather, curvis_render_frames_peers) and a 4-byte all-reduce is the step barrier; "
                         "nccl = tiles rendered locally, then one coalesced NCCL all-gather per step")
    ap.add_argument("--tiles", default="rows", choices=["blocks", "rows"],
                    help="N > 1, fused gather: what is interleaved over the ranks — whole rows (default) or blocks of 64 pixels of a "
                         "row (every rank owns an N-th of every row; measured 2 %% slower at N = 8: profiles/r02_bench_n8_block_tiles.json)")
    ap.add_argument("--precision", default="f64_fast", choices=["f64_fast", "f64"],
                    help="f64_fast (default): fp64, right-hand side regrouped around one reciprocal per step "
                         "(CURVIS_PRECISION_F64_FAST); f64: one rounding per reference operation (CURVIS_PRECISION_F64)")
    return ap.parse_args()


def workload_config(args, n):
    return {
        "workload": f"Ellis rho=1 wormhole, {args.width}x{args.height}, default camera (l=5, fwd -x, diag 43, focal 15), "
                    f"escape_radius 100 / max_iterations 40000 / step 0.05 (forward Euler, early exit), nearest u8 lookup in "
                    f"two {BG_W}x{BG_H} RGBA8 backgrounds",
        "frames_per_step": n,
        "precision": {"f64_fast": "CURVIS_PRECISION_F64_FAST: fp64, same Euler scheme, right-hand side regrouped around one "
                                  "reciprocal per step, momenta pre-scaled by the step, (sin, cos) of theta carried and rotated by the "
                                  "step's dtheta (each operation <= 1 ulp); rays whose step count or texel lies inside the guard band of a "
                                  "decision boundary are re-integrated with the CURVIS_PRECISION_F64 arithmetic inside the timed region, so "
                                  "the frame's integers equal CURVIS_PRECISION_F64's (`parity_check` compares with the CPU oracle)",
                      "f64": "CURVIS_PRECISION_F64: fp64, one rounding per reference operation"}[args.precision],
        "parallelism": "single GPU" if n == 1 else (
            f"{n} frames/step (camera path), " + (f"{BLOCK_WIDTH}-pixel blocks of the rows of each frame interleaved over {n} ranks (rank g: blocks g, g+{n}, ... in row-major order — an {n}-th of every row)"
                                                   if args.tiles == "blocks" else f"rows of each frame interleaved over {n} ranks (rank g: rows g, g+{n}, ...)") +
            ": one batched launch per rank whose epilogue stores every pixel into "
            f"the complete frames of all {n} ranks over NVLink peer memory (fused render + all-gather), then a 4-byte NCCL all-reduce as the step barrier"
            if args.gather == "peers" else
            f"{n} frames/step (camera path), each row-tiled over {n} ranks: one batched launch per rank + the {n} NCCL all-gathers of row tiles "
            f"(one NCCL group) between two renders"),
        "l2": "256 MiB device buffer rewritten between steps inside the timed region (L2 flush); the kernel is ALU-bound",
    }


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int):
                names[getattr(nv, n)] = n[len("nvmlClocksThrottleReason"):]
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if bit and (mask & bit) and name not in ("None", "All", "GpuIdle", "ApplicationsClocksSetting"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------- reference arm
def oracle_sample(args, rows_per_step: int, threads: int):
    """Times the oracle port on a bounded sample: `rows_per_step` rows of the frame, spread
    evenly over its height."""
    from curvis_b200 import scenes
    from oracle import oracle as O

    bp = scenes.decodable_background(BG_W, BG_H)
    bn = scenes.decodable_background(BG_W, BG_H, negative=True)
    cam = O.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP,
                   scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, args.width, args.height)
    g = O.metric("ellis", rho=1.0)
    s = O.sim(scenes.DEFAULT_MAX_ITERATIONS, scenes.DEFAULT_ESCAPE_RADIUS, scenes.DEFAULT_STEP)
    stride = max(1, args.height // rows_per_step)
    first = stride // 2

    def one():
        t0 = time.perf_counter()
        _, _, st = O.render_rows(g, cam, s, bp, bn, row_begin=first, row_end=args.height, row_stride=stride,
                                 threads=threads, with_records=False)
        return time.perf_counter() - t0, st

    return one, f"rows {first}::{stride} of the {args.width}x{args.height} frame ({len(range(first, args.height, stride))} rows)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # The reference (pure Rust, no cargo in the image) cannot be built here: the arm is the
    # oracle port, on ONE thread because the reference's render_image is single-threaded
    # (README.md:110, src/systems.rs:316-326 is a plain nested loop).
    one, sample = oracle_sample(args, rows_per_step=4, threads=1)
    for _ in range(args.warmup):
        one()
    t_total, steps_total = 0.0, 0
    for _ in range(args.steps):
        dt, st = one()
        t_total += dt
        steps_total += st["total_steps"]
    value = steps_total / t_total
    ncpu = os.cpu_count() or 1
    one_mt, _ = oracle_sample(args, rows_per_step=4 * min(ncpu, 64), threads=ncpu)
    dt, st = one_mt()
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "sample_per_step": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": ncpu,
                         "note": "reference is single-threaded; all-core figure of the same port in cpu_all_cores"},
        "cpu_all_cores": {"value": st["total_steps"] / dt, "unit": UNIT, "cores": ncpu},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def strong_scaling(args, cv, system, sim, PREC, ctx, rank, n, local, dev, stream, dist, interleaved_rows, scenes, iters=5):
    """ONE frame over N ranks (strong scaling).  Forms: `fused` = rows interleaved over the ranks, every pixel stored into
    every rank's complete frame over NVLink by the render kernel, 4-byte all-reduce as the barrier; `nccl` = contiguous row
    tiles + ONE ncclAllGather of the frame (what BASELINE configs[3] names).  Reference: the same frame rendered whole by one
    rank (all ranks do it at once, max over ranks).  Device time (CUDA events), max over ranks, per frame."""
    import torch
    out = {}
    token = torch.zeros(1, dtype=torch.int32, device=dev)
    keep_camera = system.camera
    for label, (Ws, Hs) in (("3840x2160", (3840, 2160)), ("7680x4320", (7680, 4320))):
        if Hs % n:
            continue
        cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, scenes.DEFAULT_FOCAL_LENGTH,
                        scenes.DEFAULT_DIAGONAL, Ws, Hs)
        system.camera = cam
        fbytes = Ws * Hs * 3
        rows = Hs // n
        mine = cv.PeerBuffer.create(ctx, fbytes)
        handles = [None] * n
        dist.all_gather_object(handles, mine.handle)
        bufs = [mine if r == rank else cv.PeerBuffer.open(ctx, handles[r], fbytes) for r in range(n)]
        gathered = mine.as_tensor(local)
        tile = torch.empty(rows * Ws * 3, dtype=torch.uint8, device=dev)
        full = torch.empty(fbytes, dtype=torch.uint8, device=dev)
        whole = torch.empty(fbytes, dtype=torch.uint8, device=dev)
        r0, r1, stride = interleaved_rows(Hs, rank, n)

        def fused():
            system.render_frames_peers([cam], *sim, r0, r1, [b.ptr for b in bufs], stream.cuda_stream, row_stride=stride, precision=PREC)
            dist.all_reduce(token)

        from curvis_b200.distributed import interleaved_blocks
        b0, b1, bstride, bw = interleaved_blocks(Hs, Ws, rank, n, BLOCK_WIDTH)

        def render_blocks(**kw):
            return system.render_frames_peers([cam], *sim, b0, b1, [b.ptr for b in bufs], stream.cuda_stream, row_stride=bstride, block_width=bw,
                                              precision=PREC, **kw)

        def fused_blocks():         # blocks of BLOCK_WIDTH pixels of a row interleaved over the ranks: every row is shared by all
            render_blocks()
            dist.all_reduce(token)

        def fused_contiguous():     # the same fused launch on CONTIGUOUS row tiles (what the NCCL form renders)
            system.render_frames_peers([cam], *sim, rank * rows, (rank + 1) * rows, [b.ptr for b in bufs], stream.cuda_stream, row_stride=1, precision=PREC)
            dist.all_reduce(token)

        def nccl():
            system.render_rows_device(*sim, rank * rows, (rank + 1) * rows, tile.data_ptr(), stream.cuda_stream, precision=PREC)
            dist.all_gather_into_tensor(full, tile)

        def single():
            system.render_rows_device(*sim, 0, Hs, whole.data_ptr(), stream.cuda_stream, precision=PREC)

        def timed(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        def timeline(render, iters=8):
            """Where a frame's time goes on each rank: `busy` = the rank's own launches (pre-pass + render + re-integration
            + peer stores), `wait` = from its last kernel to the end of the all-reduce (the slowest rank's surplus + the
            collective), per iteration, CUDA events on the launching stream; all ranks' means."""
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
            for _ in range(2):
                render(); dist.all_reduce(token)
            torch.cuda.synchronize()
            dist.barrier()
            for e in ev:
                e[0].record(); render(); e[1].record(); dist.all_reduce(token); e[2].record()
            torch.cuda.synchronize()
            busy = [e[0].elapsed_time(e[1]) for e in ev]
            wait = [e[1].elapsed_time(e[2]) for e in ev]
            mine_t = torch.tensor([sum(busy) / iters, max(busy), min(busy), sum(wait) / iters], dtype=torch.float64, device=dev)
            allr = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(n)]
            dist.all_gather(allr, mine_t)
            return {"rank_busy_ms_mean": [round(float(a[0]), 3) for a in allr], "rank_busy_ms_max": [round(float(a[1]), 3) for a in allr],
                    "rank_busy_ms_min": [round(float(a[2]), 3) for a in allr], "rank_wait_ms_mean": [round(float(a[3]), 3) for a in allr]}

        st = system.render_rows_device(*sim, 0, Hs, whole.data_ptr(), stream.cuda_stream, want_stats=True, precision=PREC)
        steps = int(st["total_steps"])
        kernel_single_ms = st["kernel_ms"]
        t_single, t_fused, t_nccl = timed(single), timed(fused), timed(nccl)
        gathered.zero_()
        dist.barrier()
        t_blocks = timed(fused_blocks)
        tl_blocks = timeline(render_blocks)
        torch.cuda.synchronize()
        bad_blocks = (gathered.view(-1, 3) != whole.view(-1, 3)).any(dim=1).sum().to(torch.int64)
        dist.all_reduce(bad_blocks)
        ctx.set_option("guard", 0)
        t_blocks_raw = timed(fused_blocks)
        ctx.set_option("guard", 1)
        gathered.zero_()
        dist.barrier()
        t_fused_contig = timed(fused_contiguous)
        torch.cuda.synchronize()
        bad_contig = (gathered.view(-1, 3) != whole.view(-1, 3)).any(dim=1).sum().to(torch.int64)
        dist.all_reduce(bad_contig)
        tl = timeline(lambda: system.render_frames_peers([cam], *sim, r0, r1, [b.ptr for b in bufs], stream.cuda_stream, row_stride=stride, precision=PREC))
        ctx.set_option("guard", 0)          # the regrouped kernel alone: what the guard band's second launch costs at this tile size
        t_fused_raw = timed(fused)
        tl_raw = timeline(lambda: system.render_frames_peers([cam], *sim, r0, r1, [b.ptr for b in bufs], stream.cuda_stream, row_stride=stride, precision=PREC))
        ctx.set_option("guard", 1)
        # this rank's own kernel inside the split frame (the slowest rank bounds the frame)
        kst = system.render_frames_peers([cam], *sim, r0, r1, [b.ptr for b in bufs], stream.cuda_stream, row_stride=stride, precision=PREC, want_stats=True)
        kms = torch.tensor([kst["kernel_ms"]], dtype=torch.float64, device=dev)
        kmax, kmin = kms.clone(), kms.clone()
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(kmin, op=dist.ReduceOp.MIN)
        kall = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(n)]
        dist.all_gather(kall, kms)
        redo_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(n)]
        dist.all_gather(redo_all, torch.tensor([kst["n_reintegrated"]], dtype=torch.int64, device=dev))
        dist.all_reduce(token)
        torch.cuda.synchronize()
        bad = torch.stack([(gathered.view(-1, 3) != whole.view(-1, 3)).any(dim=1).sum(), (full.view(-1, 3) != whole.view(-1, 3)).any(dim=1).sum()]).to(torch.int64)
        dist.all_reduce(bad)
        out[label] = {
            "ray_steps": steps, "one_rank_ms": t_single, "one_rank_kernel_ms": kernel_single_ms,
            "fused_peer_stores": {"ms_per_frame": t_fused, "speedup": t_single / t_fused, "ray_steps_per_s": steps / (t_fused * 1e-3),
                                  "rank_kernel_ms_max": float(kmax.item()), "rank_kernel_ms_min": float(kmin.item()),
                                  "rank_kernel_ms": [round(float(k.item()), 3) for k in kall],
                                  "rank_rays_reintegrated": [int(r.item()) for r in redo_all],
                                  "ms_per_frame_contiguous_tiles": t_fused_contig, "speedup_contiguous_tiles": t_single / t_fused_contig,
                                  "differing_pixels_contiguous_tiles": int(bad_contig.item()),
                                  "ms_per_frame_without_guard_band": t_fused_raw, "speedup_without_guard_band": t_single / t_fused_raw,
                                  "ideal_ms": t_single / n,
                                  "timeline": tl, "timeline_without_guard_band": tl_raw,
                                  "limiter": "fixed costs that do not shrink with the tile: a frame takes max over ranks of `timeline.rank_busy_ms` "
                                             "(that rank's launches; run-to-run spread in _min/_max: which warp a 10^4-step ray shares its scheduler with) "
                                             "+ the all-reduce (`rank_wait_ms` of the slowest rank); ms_per_frame - ms_per_frame_without_guard_band = the guard "
                                             "band's re-integration launch (one strict ray's latency, ~0.5 ms); the rest of busy - ideal_ms = the persistent "
                                             "kernel's drain tail (one ray's latency, ~0.4 ms) + row imbalance"},
            "fused_peer_stores_blocks": {"what": "the fused form with blocks of %d pixels of a row interleaved over the ranks instead of whole rows "
                                                 "(curvis_render_frames_peers_blocks): every rank owns an N-th of every row" % bw,
                                         "ms_per_frame": t_blocks, "speedup": t_single / t_blocks, "ray_steps_per_s": steps / (t_blocks * 1e-3),
                                         "ms_per_frame_without_guard_band": t_blocks_raw, "speedup_without_guard_band": t_single / t_blocks_raw,
                                         "differing_pixels_vs_one_rank": int(bad_blocks.item()), "timeline": tl_blocks},
            "nccl_all_gather": {"ms_per_frame": t_nccl, "speedup": t_single / t_nccl, "ray_steps_per_s": steps / (t_nccl * 1e-3),
                                "gather_bytes": fbytes},
            "differing_pixels_vs_one_rank": {"fused": int(bad[0].item()), "nccl": int(bad[1].item()), "pixels_checked": n * Ws * Hs},
        }
        torch.cuda.synchronize()
        del gathered
        for r, b in enumerate(bufs):
            if r != rank:
                b.close()
        dist.barrier()
        mine.close()
    system.camera = keep_camera
    return out


def oracle_parity(args, system, sim, strict_frame, fast_frame, fast_precision, n_rows=72):
    """Both kernels against the CPU oracle on `n_rows` rows spread over the frame (all host cores, records with the
    trajectory diagnostics), split into regular / chaotic rays by the survey's classifier (oracle/classify.py)."""
    import numpy as np
    from curvis_b200 import _abi, scenes
    from oracle import classify, oracle as O

    bp = scenes.decodable_background(BG_W, BG_H)
    bn = scenes.decodable_background(BG_W, BG_H, negative=True)
    cam = O.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP,
                   scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, args.width, args.height)
    stride = max(1, args.height // n_rows)
    first = stride // 2
    rows = list(range(first, args.height, stride))
    t0 = time.perf_counter()
    ref_rgb, ref_rec, _ = O.render_rows(O.metric("ellis", rho=1.0), cam, O.sim(*sim), bp, bn, row_begin=first, row_end=args.height,
                                        row_stride=stride, threads=os.cpu_count() or 1, with_records=True)
    out = {"rows": f"{first}::{stride} ({len(rows)} rows, {len(rows) * args.width} rays)", "oracle_seconds": round(time.perf_counter() - t0, 2),
           "classifier": "chaotic <=> |p_l|_final > 1.05 or min |sin theta| < 1e-3 on the ORACLE's record (SURVEY.md 8c)"}
    for name, frame, prec in (("f64", strict_frame, _abi.PRECISION_F64), ("f64_fast", fast_frame, fast_precision)):
        recs = np.concatenate([system.render_rows(*sim, r, r + 1, with_records=True, precision=prec)[1] for r in rows], axis=0)
        out[name] = classify.compare(frame[first::stride], recs, ref_rgb, ref_rec)
    out["chaotic_fraction"] = out["f64"]["chaotic_fraction"]
    out["differing_regular"] = max(out["f64_fast"]["differing_pixels_regular"], out["f64_fast"]["differing_records_regular"])
    return out


# ----------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    from curvis_b200.distributed import interleaved_blocks, interleaved_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: curvis_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = world
    Wd, Ht = args.width, args.height
    if Ht % n:
        raise SystemExit(f"height {Ht} not divisible by {n} ranks")
    rows = Ht // n
    row_begin, row_end = rank * rows, (rank + 1) * rows

    lib = _abi.load_library()
    ctx = cv.Context([local])
    bp = scenes.decodable_background(BG_W, BG_H)
    bn = scenes.decodable_background(BG_W, BG_H, negative=True)
    t_up0 = time.perf_counter()
    system = cv.RelativisticSystem(
        cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn),
        cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP,
                  scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, Wd, Ht), context=ctx)
    background_upload_ms = (time.perf_counter() - t_up0) * 1e3
    sim = (scenes.DEFAULT_MAX_ITERATIONS, scenes.DEFAULT_ESCAPE_RADIUS, scenes.DEFAULT_STEP)
    PREC = {"f64_fast": _abi.PRECISION_F64_FAST, "f64": _abi.PRECISION_F64}[args.precision]
    KERNEL = {"f64_fast": "render_rows_f64_fast<FastEllis, 1, 5, 1>", "f64": "render_rows_f64_lean<ShapeEllis, 0, 0, 1, 1>"}[args.precision]
    KERNEL_NOTE = {"f64_fast": "the longest-first instantiation; collect_long_rays runs before it and render_rows_f64_lean<ShapeEllis, 0, 0, 1, 1> in list "
                               "mode (the guard band's re-integration) after it, both inside kernel_ms", "f64": "kernel_variant 5"}[args.precision]

    stream = torch.cuda.current_stream()
    frames = [torch.empty(Ht * Wd * 3, dtype=torch.uint8, device=dev) for _ in range(n)]   # complete frames
    # this rank's tile of each frame; double-buffered so the all-gathers of step k (comm stream)
    # overlap the render kernel of step k+1 (compute stream)
    tiles_buf = [torch.empty(n * rows * Wd * 3, dtype=torch.uint8, device=dev) for _ in range(2)]
    tiles = tiles_buf[0]
    comm_stream = torch.cuda.Stream(device=dev) if n > 1 else None
    tiles_free = [torch.cuda.Event(), torch.cuda.Event()]      # recorded on the comm stream after a buffer was gathered
    step_index = [0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # synthetic camera path: frame f sits 0.02*f further out along l (a dolly move), same orientation
    cameras = [cv.Camera((0.0, scenes.DEFAULT_CAMERA_POSITION[1] + 0.02 * f, scenes.DEFAULT_CAMERA_POSITION[2], 0.0),
                         scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, Wd, Ht)
               for f in range(n)]
    tile_bytes = rows * Wd * 3

    def barrier():
        if world > 1:
            dist.barrier()

    # Overlapping the all-gathers of step k with the render kernel of step k+1 does not pay here: the render
    # kernel is persistent and saturates every scheduler, and NCCL CTAs that become co-resident with it are
    # starved by the warp scheduler (DESIGN.md sections 5-6) — measured at N=2: one step in four stalls 8-9 ms.
    # Default: the gathers run between two renders on the compute stream (N x ~0.1 ms per step); the frame
    # read-back of the e2e path is a DMA and does overlap.  CURVIS_BENCH_OVERLAP=1 restores the overlapped form.
    overlap = os.environ.get("CURVIS_BENCH_OVERLAP", "0") == "1"
    copied = [torch.cuda.Event()]                 # recorded on the comm (copy) stream after this rank's frame has been read back

    # ---- fused render + all-gather (--gather peers): every rank owns two buffers of n complete frames (steps alternate
    # between them), exported to its peers through CUDA IPC; the render kernel of every rank stores each pixel into all n
    # buffers of the step's set — its own and, over NVLink, its peers'.  No collective moves pixels; a 4-byte all-reduce
    # after the launch tells a rank that its peers' kernels have ended.
    frame_bytes = Ht * Wd * 3
    peers_mode = n > 1 and args.gather == "peers"
    peer_sets, my_frames, token = [], [], None
    if peers_mode:
        for s_ in range(2):
            mine = cv.PeerBuffer.create(ctx, n * frame_bytes)
            handles = [None] * n
            dist.all_gather_object(handles, mine.handle)
            peer_sets.append([mine if r == rank else cv.PeerBuffer.open(ctx, handles[r], n * frame_bytes) for r in range(n)])
            my_frames.append(mine.as_tensor(local))
        token = torch.zeros(1, dtype=torch.int32, device=dev)

    def peers_step(want_stats=False, readback=False):
        flush.fill_(1)                                # L2 flush between steps
        k = step_index[0]
        step_index[0] += 1
        cur = peer_sets[k & 1]
        # interleaved rows: rank g renders rows g, g+n, g+2n, ... of every frame — the same mix of short (sky) and
        # long (throat-grazing) rays on every rank; a contiguous tile of central rows holds ~3 % more steps than the mean
        # (--tiles blocks: the same with 64-pixel blocks of a row instead of rows — the two or three rows of a frame that hold
        # its 10^4-step rays are then shared by all ranks)
        if args.tiles == "blocks":
            r0, r1, stride, bw = interleaved_blocks(Ht, Wd, rank, n, BLOCK_WIDTH)
        else:
            (r0, r1, stride), bw = interleaved_rows(Ht, rank, n), 0
        st = system.render_frames_peers(cameras, *sim, r0, r1, [b.ptr for b in cur], stream.cuda_stream,
                                        want_stats=want_stats, row_stride=stride, block_width=bw, precision=PREC)
        # this rank's read-back of step k-1 (which read set (k-1)&1) must have finished before the barrier of step k lets
        # anybody launch step k+1 into that set
        stream.wait_event(copied[0])
        dist.all_reduce(token)                        # step barrier: every rank's kernel of step k has ended
        frames[:] = [my_frames[k & 1][f * frame_bytes:(f + 1) * frame_bytes] for f in range(n)]
        if readback:
            gathered = torch.cuda.Event()
            gathered.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(gathered)
                host_frames[0].copy_(frames[rank], non_blocking=True)
                copied[0] = torch.cuda.Event()
                copied[0].record(comm_stream)
        return st

    def device_step(want_stats=False, readback=False):
        """One step of the resident path: n frames; this rank's row tile of every frame in ONE
        batched launch, gathered so that every rank holds every complete frame (fused peer stores, or
        NCCL all-gathers with --gather nccl); with `readback` (the e2e path) rank r also copies
        complete frame r to its pinned host buffer."""
        if peers_mode:
            return peers_step(want_stats, readback)
        flush.fill_(1)                                # L2 flush between steps
        if n == 1:
            return system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=want_stats,
                                             precision=PREC)
        buf = step_index[0] & 1
        step_index[0] += 1
        cur = tiles_buf[buf]
        stream.wait_event(tiles_free[buf])          # the gather that last read this buffer has finished
        st = system.render_frames_device(cameras, *sim, row_begin, row_end, cur.data_ptr(), stream.cuda_stream, want_stats=want_stats,
                                         precision=PREC)
        if overlap:
            rendered = torch.cuda.Event()
            rendered.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(rendered)
                for f in range(n):
                    dist.all_gather_into_tensor(frames[f], cur[f * tile_bytes:(f + 1) * tile_bytes])
                    if readback and f == rank:
                        host_frames[0].copy_(frames[f], non_blocking=True)
                tiles_free[buf].record(comm_stream)
            return st
        stream.wait_event(copied[0])                # the previous step's read-back of frames[rank] has finished
        with dist._coalescing_manager():            # the n all-gathers as ONE NCCL group (one launch)
            for f in range(n):
                dist.all_gather_into_tensor(frames[f], cur[f * tile_bytes:(f + 1) * tile_bytes])
        tiles_free[buf].record(stream)
        if readback:
            gathered = torch.cuda.Event()
            gathered.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(gathered)
                host_frames[0].copy_(frames[rank], non_blocking=True)
                copied[0] = torch.cuda.Event()
                copied[0].record(comm_stream)
        return st

    # steps of this rank's share of one job-step (deterministic) -> total over ranks
    st = device_step(want_stats=True)
    share = torch.tensor([st["total_steps"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(share)
    steps_per_job_step = int(share.item())        # Euler steps of the n frames of one step
    frame_steps = steps_per_job_step // n

    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    barrier()
    # N > 1: the gathered frame r on rank r against the same frame rendered whole on this rank alone
    gather_check = None
    if n > 1:
        whole = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
        keep_camera, system.camera = system.camera, cameras[rank]
        system.render_rows_device(*sim, 0, Ht, whole.data_ptr(), stream.cuda_stream, want_stats=True, precision=PREC)
        system.camera = keep_camera
        bad = (whole.view(-1, 3) != frames[rank].view(-1, 3)).any(dim=1).sum().to(torch.int64).reshape(1)
        dist.all_reduce(bad)
        gather_check = {"what": "complete frame r as gathered on rank r vs. the same frame rendered whole on rank r alone, all ranks",
                        "pixels": n * Wd * Ht, "differing_pixels": int(bad.item())}
        del whole
    launches0 = lib.curvis_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            device_step()
        if comm_stream is not None:
            stream.wait_stream(comm_stream)        # the last step's all-gathers are inside the timed region
        e1.record()
        torch.cuda.synchronize()
    barrier()
    launches = lib.curvis_kernel_launch_count() - launches0
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches_t = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches_t)
    elapsed_ms = float(elapsed_ms.item())
    value = steps_per_job_step * args.steps / (elapsed_ms * 1e-3)

    # kernel-only duration of the dominant kernel (events recorded by the library around it)
    kernel_ms = []
    for _ in range(3):
        s2 = device_step(want_stats=True)
        kernel_ms.append(s2["kernel_ms"])
    kernel_ms = sum(kernel_ms) / len(kernel_ms)
    tile_steps_local = st["total_steps"]

    # ---- e2e through the host-buffer C-ABI call (per-frame parameters in, RGB8 out to host)
    # N > 1: rank r owns frame r of the step's batch (the frame it would encode / write): it copies that one
    # complete frame to its pinned host buffer, so the step's N read-backs run on N PCIe links in parallel
    host_frames = [torch.empty(Ht * Wd * 3, dtype=torch.uint8).pin_memory()] if n > 1 else None
    param_bytes = (C_sizeof(_abi.CurvisMetric) + C_sizeof(_abi.CurvisCamera) + C_sizeof(_abi.CurvisSim))

    # the caller's frame buffer, reused every step and registered once with curvis_host_register (page-locked, as
    # the contract's "pinned host memory"): the kernel stores its pixels straight into it
    host_frame = np.empty((Ht, Wd, 3), dtype=np.uint8)
    pageable_frame = np.empty((Ht, Wd, 3), dtype=np.uint8)  # an unregistered buffer, for the e2e_pageable figure
    if n == 1:
        ctx.register_host_buffer(host_frame)

    def e2e_step():
        if n == 1:
            system.render_image(*sim, out=host_frame, precision=PREC)   # curvis_render_image into the registered caller frame
        else:
            device_step(readback=True)            # the resident step + the read-back (a DMA) of this rank's frame under the next render

    e2e_step()
    torch.cuda.synchronize()
    barrier()
    trace_events = [] if os.environ.get("CURVIS_BENCH_TRACE") else None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
        if trace_events is not None:
            ev = torch.cuda.Event(enable_timing=True); ev.record(stream); trace_events.append((ev, time.perf_counter() - t0))
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_sync = time.perf_counter() - t0
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if trace_events is not None:
        print(f"[trace rank {rank}] enqueue done {t_enq * 1e3:.2f} ms, synchronized {t_sync * 1e3:.2f} ms, after barrier {float(e2e_s.item()) * 1e3:.2f} ms; "
              f"render-end gaps (ms): {[round(trace_events[i][0].elapsed_time(trace_events[i + 1][0]), 2) for i in range(len(trace_events) - 1)]}; "
              f"host enqueue times (ms): {[round(t * 1e3, 1) for _, t in trace_events]}", file=sys.stderr, flush=True)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = steps_per_job_step * args.steps / float(e2e_s.item())

    # ---- strong scaling (BASELINE configs[3] and north_star's ">= 6x at 8 GPUs" for ONE frame): one 4K and one 8K Ellis frame
    # at the default settings split over the N ranks, both exchange forms, against the same frame rendered whole by one rank
    strong = None
    if n > 1:
        strong = strong_scaling(args, cv, system, sim, PREC, ctx, rank, n, local, dev, stream, dist, interleaved_rows, scenes)

    if peers_mode:                                  # unmap the peers' buffers, then (after everybody has) free this rank's
        torch.cuda.synchronize()
        frames[:] = []
        my_frames[:] = []
        for bufs in peer_sets:
            for r, b in enumerate(bufs):
                if r != rank:
                    b.close()
        barrier()
        for bufs in peer_sets:
            bufs[rank].close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the same call into a pageable (unregistered) frame: D2H into the library's pinned staging buffer + host copy
    e2e_pageable = None
    if n == 1:
        system.render_image(*sim, out=pageable_frame, precision=PREC)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            system.render_image(*sim, out=pageable_frame, precision=PREC)
        e2e_pageable = {"value": frame_steps * args.steps / (time.perf_counter() - t0), "unit": UNIT,
                        "note": "unregistered caller frame: device frame -> pinned staging (chunked DMA) -> host copy"}

    # ---- cold e2e at N=1: re-upload both backgrounds every frame as well
    e2e_cold = None
    if n == 1:
        reps = max(2, min(args.steps, 3))
        system._upload(+1, system.background_positive)        # untimed warm-up of the upload path
        system._upload(-1, system.background_negative)
        t0 = time.perf_counter()
        for _ in range(reps):
            system._upload(+1, system.background_positive)
            system._upload(-1, system.background_negative)
            system.render_image(*sim, out=host_frame, precision=PREC)
        dt = time.perf_counter() - t0
        e2e_cold = {"value": frame_steps * reps / dt, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * BG_W * BG_H * 4 + param_bytes, "d2h_bytes_per_step": Wd * Ht * 3,
                    "note": "both backgrounds re-uploaded from pageable host memory every frame"}

    fp64_peak, fp32_peak = ctx.measure_fma_peak()
    flop = FLOP_PER_STEP["ellis"]
    kernel_rate = tile_steps_local / (kernel_ms * 1e-3)       # this rank's kernel, steps/s
    achieved_tf = kernel_rate * flop / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes_per_ray = 4 + 3                                  # one RGBA8 texel read + 3 B written
    tile_rays = rows * Wd * n
    hbm_achieved = tile_rays * alg_bytes_per_ray / (kernel_ms * 1e-3) / 1e9
    traffic, fp64_pipe = None, None
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))[KERNEL]
        traffic = prof.get("dram_bytes_per_launch")
        ipw = prof.get("fp64_warp_instructions_per_warp_step")
        if ipw:
            # instruction-level view: fp64-pipe warp-instructions issued per second vs one per two
            # cycles per scheduler (4 per SM) at the SM clock sampled during the timed region
            mhz = clocks.summary().get("sm_mhz") or 1965
            issued = kernel_rate / 32.0 * ipw
            peak_issue = sm_count * 4 * mhz * 1e6 / 2.0
            fp64_pipe = {"fp64_warp_instr_per_s": issued, "peak_warp_instr_per_s": peak_issue, "frac": issued / peak_issue,
                         "fp64_warp_instr_per_warp_step": ipw, "sm__pipe_fp64_cycles_active_pct_ncu": prof.get("sm__pipe_fp64_cycles_active_pct"),
                         "ncu_kernel_ms": prof.get("ncu_kernel_ms"), "source": prof.get("source")}
    except Exception:
        pass
    mhz_max = clocks.summary().get("sm_max_mhz") or 1965
    peak_nominal = sm_count * 64 * 2 * mhz_max * 1e6 / 1e12       # 64 fp64 FMA lanes per SM x 2 flop x f_clk
    roofline = {
        "bound": "fp64_alu",
        "bound_note": "per-ray ODE: ~1e4 flop/B and no dense contraction, so neither hbm nor tensor bounds it (DESIGN.md section 5); "
                      "the hbm figure is reported below for completeness",
        "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
        "peak_measured": fp64_peak, "peak_nominal": peak_nominal, "frac_of_nominal": achieved_tf / peak_nominal,
        "peak_nominal_note": f"{sm_count} SMs x 64 fp64 FMA lanes x 2 flop x {mhz_max} MHz",
        "traffic": traffic, "fp64_pipe": fp64_pipe,
        "peak_source": "live DFMA micro-kernel on this GPU (curvis_measure_fma_peak); MEASURED_PEAKS.json has no fp64 entry",
        "flop_per_ray_step": flop, "kernel": KERNEL, "kernel_note": KERNEL_NOTE, "kernel_ms": kernel_ms,
        "kernel_ray_steps_per_s": kernel_rate,
        "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "algorithmic_bytes_per_ray": alg_bytes_per_ray},
        "fp32_fma_peak_tflops": fp32_peak,
    }

    # ---- the operation-for-operation kernel (CURVIS_PRECISION_F64) next to the headline, the headline frame against it
    # pixel for pixel, and BOTH against the CPU oracle on strided rows, split into regular / chaotic rays
    strict_mode, parity_check, raw_fast, full_identity = None, None, None, None
    if n == 1:
        sms = []
        for _ in range(3):
            s4 = system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=True,
                                           precision=_abi.PRECISION_F64)
            sms.append(s4["kernel_ms"])
        strict_frame = frames[0].clone()
        s5 = system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=True, precision=PREC)
        fast_frame = frames[0].clone()
        differing = int((strict_frame.view(-1, 3) != fast_frame.view(-1, 3)).any(dim=1).sum().item())
        strict_mode = {"precision": "CURVIS_PRECISION_F64: one rounding per reference operation (six correctly rounded divisions, "
                                    "one square root, sincos per step)",
                       "kernel": "render_rows_f64_lean<ShapeEllis, 0, 0, 1, 1> (kernel_variant 5: the six reciprocals of a step from two seeds, the next step's shape function and sincos carried across the loop's back edge)",
                       "value": s4["total_steps"] / (min(sms) * 1e-3), "unit": UNIT,
                       "kernel_ms": min(sms), "frac_of_fp64_fma_peak": s4["total_steps"] / (min(sms) * 1e-3) * flop / 1e12 / fp64_peak}
        if args.precision == "f64_fast":
            ctx.set_option("guard", 0)
            rms = []
            for _ in range(3):
                s6 = system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=True, precision=PREC)
                rms.append(s6["kernel_ms"])
            ctx.set_option("guard", 1)
            raw_fast = {"what": "the regrouped kernel alone (ctx option guard = 0): no guard band, no re-integration — not what `value` times",
                        "kernel_ms": min(rms), "value": s6["total_steps"] / (min(rms) * 1e-3), "unit": UNIT,
                        "differing_pixels_vs_f64": int((strict_frame.view(-1, 3) != frames[0].view(-1, 3)).any(dim=1).sum().item())}
            # guard = 2: the kicked rays (stiffness >= 1) are re-integrated too, so EVERY ray of the frame carries the
            # operation-for-operation arithmetic — identity with CURVIS_PRECISION_F64 by construction, not by measurement
            ctx.set_option("guard", 2)
            gms = []
            for _ in range(3):
                s7 = system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=True, precision=PREC)
                gms.append(s7["kernel_ms"])
            ctx.set_option("guard", 1)
            full_identity = {"what": "ctx option guard = 2: every ray outside the guard band's certificate (the kicked rays included) is re-integrated "
                                     "with the CURVIS_PRECISION_F64 arithmetic — not what `value` times",
                             "kernel_ms": min(gms), "value": s7["total_steps"] / (min(gms) * 1e-3), "unit": UNIT,
                             "n_reintegrated": int(s7["n_reintegrated"]), "n_kicked": int(s7["n_kicked"]),
                             "differing_pixels_vs_f64": int((strict_frame.view(-1, 3) != frames[0].view(-1, 3)).any(dim=1).sum().item()),
                             "speedup_vs_f64_kernel": min(sms) / min(gms)}
        parity_check = {"f64_fast_vs_f64_kernel": {"pixels": Wd * Ht, "differing_pixels": differing,
                                                   "total_steps_equal": bool(s4["total_steps"] == s5["total_steps"]),
                                                   "escape_counters_equal": all(s4[k] == s5[k] for k in ("n_positive", "n_negative", "n_not_escaped", "n_clamped")),
                                                   "n_reintegrated": int(s5["n_reintegrated"]),
                                                   "reintegrated_fraction": s5["n_reintegrated"] / (Wd * Ht)}}
        if not args.no_cpu_baseline:
            parity_check["against_oracle"] = oracle_parity(args, system, sim, strict_frame.view(Ht, Wd, 3).cpu().numpy(),
                                                           fast_frame.view(Ht, Wd, 3).cpu().numpy(), PREC)
        del strict_frame, fast_frame

    # ---- opt-in fp32 mode (CURVIS_PRECISION_F32), reported next to the headline, never instead of it
    fast_mode = None
    if n == 1:
        fms = []
        for _ in range(3):
            s3 = system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=True,
                                           precision=_abi.PRECISION_F32)
            fms.append(s3["kernel_ms"])
        band = (Ht // 2 - 32, Ht // 2 + 32)
        _, r64 = system.render_rows(*sim, *band, with_records=True)
        _, r32 = system.render_rows(*sim, *band, with_records=True, precision=_abi.PRECISION_F32)
        dxy = (np.minimum(np.abs(r64["texel_x"].astype(np.int64) - r32["texel_x"].astype(np.int64)),
                          BG_W - np.abs(r64["texel_x"].astype(np.int64) - r32["texel_x"].astype(np.int64))) <= 1) & \
              (np.abs(r64["texel_y"].astype(np.int64) - r32["texel_y"].astype(np.int64)) <= 1)
        fast_mode = {
            "precision": "CURVIS_PRECISION_F32: f32 right-hand side + Kahan-compensated state (extension, off by default; DEMOTED in "
                         "round 2: 1.2x the fp64 headline for ~1e-3 of the pixels differing is not worth its tolerance — kept for A/B only)",
            "value": s3["total_steps"] / (min(fms) * 1e-3), "unit": UNIT, "kernel_ms": min(fms),
            "fp32_fma_peak_tflops": fp32_peak, "frac_of_fp32_peak": s3["total_steps"] / (min(fms) * 1e-3) * flop / 1e12 / fp32_peak,
            "deviation_vs_parity_kernel": {
                "sample": f"rows {band[0]}..{band[1]} of the frame",
                "escape_side_equal": float((r64["side"] == r32["side"]).mean()),
                "step_count_equal": float((r64["steps"] == r32["steps"]).mean()),
                "texel_equal": float(((r64["texel_x"] == r32["texel_x"]) & (r64["texel_y"] == r32["texel_y"])).mean()),
                "texel_equal_or_adjacent": float(dxy.mean()),
            },
        }

    # ---- the chart-free ("pole-safe") extension, regrouped: not the reference's scheme (another chart of the same geodesics, so
    # its frames differ from render_image's at the Euler-error level), reported next to the headline, never instead of it
    chart_free = None
    if n == 1:
        cms = []
        for _ in range(3):
            s7 = system.render_rows_device(*sim, row_begin, row_end, frames[0].data_ptr(), stream.cuda_stream, want_stats=True,
                                           precision=_abi.PRECISION_F64_FAST, coordinates=_abi.COORDINATES_CARTESIAN)
            cms.append(s7["kernel_ms"])
        chart_free = {"mode": "CURVIS_COORDINATES_CARTESIAN + CURVIS_PRECISION_F64_FAST (extension, off by default): the angular state is the unit "
                              "position vector and the conserved angular momentum — no trigonometry, no coordinate pole, no chaotic rays; 17 fp64 "
                              "instructions per step; checked against its own oracle restatement (tests/test_gpu_extensions.py)",
                      "kernel_ms": min(cms), "total_steps": int(s7["total_steps"]), "value": s7["total_steps"] / (min(cms) * 1e-3), "unit": UNIT}

    # ---- the row next to the path: render_image_efficient, what the reference's binary runs — a table of escape angles
    # (sampler on the host, its photons integrated by the kernels above) + one cheap kernel per pixel.  Wall time per 4K frame
    # through the host-buffer call into a registered frame, both metrics, table photons in F64 and in F64_FAST.
    efficient = None
    if n == 1 and args.width == W4K and args.height == H4K:
        import numpy as np
        efficient = {"what": "curvis_render_image_efficient (systems.rs:333-527), 3840x2160 into a registered host frame, wall ms per frame "
                             "(reference defaults: 100 initial points, 100 refinement passes, thresholds 1e-5)"}
        ebuf = np.empty((Ht, Wd, 3), dtype=np.uint8)
        ctx.register_host_buffer(ebuf)
        try:
            for mname, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
                esys = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), system.camera, context=ctx)
                res = {}
                for pname, prec in (("f64", _abi.PRECISION_F64), ("f64_fast", _abi.PRECISION_F64_FAST)):
                    walls = []
                    for _ in range(4):
                        t0 = time.perf_counter()
                        esys.render_image_efficient(*sim, 100, 100, 1e-5, 1e-5, out=ebuf, precision=prec)
                        walls.append((time.perf_counter() - t0) * 1e3)
                    info = esys.last_efficient_info
                    res[pname] = {"wall_ms": min(walls[1:]), "table_ms": info.get("table_ms"), "table_points": info.get("table_points")}
                    if pname == "f64":
                        ref_frame = ebuf.copy()
                    else:
                        res[pname]["differing_pixels_vs_f64_table"] = int((ebuf != ref_frame).any(axis=2).sum())
                efficient[mname] = res
        finally:
            ctx.unregister_host_buffer(ebuf)

    cpu_baseline = None
    if n == 1 and not args.no_cpu_baseline:
        one, sample = oracle_sample(args, rows_per_step=60, threads=1)
        dt, cst = one()
        cpu_baseline = {"value": cst["total_steps"] / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                        "seconds": dt, "host_cores": os.cpu_count()}

    line = {
        "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, n),
        "frames_per_sec": n * args.steps / (elapsed_ms * 1e-3), "rays_per_sec": n * Wd * Ht * args.steps / (elapsed_ms * 1e-3),
        "ray_steps_per_frame": frame_steps,
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": param_bytes * n,
                "d2h_bytes_per_step": Wd * Ht * 3 * n,
                "note": ("backgrounds are part of the scene (`&self`, uploaded once: %.1f ms); per-frame input = metric+camera+sim structs; " % background_upload_ms) +
                        ("the RGB8 frame lands in a caller buffer registered once with curvis_host_register (the kernel stores into it over PCIe)"
                         if n == 1 else "every rank renders its tiles of the N frames into all ranks' frame buffers, rank r copies complete frame r to its pinned host buffer (DMA under the next render)")},
        "strong_scaling": strong,
        "gather": (args.gather if n > 1 else None),
        "gather_check": gather_check,
        "e2e_pageable": e2e_pageable,
        "e2e_cold": e2e_cold,
        "gpu_launches": int(launches_t.item()),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "strict_mode": strict_mode,
        "raw_fast_kernel": raw_fast,
        "full_identity_mode": full_identity,
        "parity_check": parity_check,
        "f32_mode": fast_mode,
        "chart_free_mode": chart_free,
        "efficient_renderer": efficient,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def C_sizeof(t):
    import ctypes
    return ctypes.sizeof(t)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
