"""Runs every BASELINE.json config on the GPU(s) of this box and checks each against the oracle on a
bounded sample of rows (all host cores).  Writes gpurun_out/configs_r1.json.

C1a  Ellis 256x144, max_iter 200, delta 0.1, R 10          (the "200-step" PR1 parity frame, Euler)
C1b  Ellis 256x144, defaults 40000 / 0.05 / 100
C2   Ellis 1920x1080, max_iter 1000, delta 0.05, R 25
C3   Interstellar 3840x2160, max_iter 2000, delta 0.05, R 45 (early exit = the "adaptive" step count)
C4   Ellis 7680x4320, defaults, row-tiled over every visible GPU in one process
C5   video: path_through.csv at 15 fps, frames 0..299, 3840x2160, Interstellar, sim as C3 (per-pixel renderer,
     batched launches of 4 frames); parity on rows of 3 sample frames
(SURVEY.md 8d resolves the configs' under-specified settings this way.)

    python tools/run_configs.py [f64|f64_fast|f32]      (default f64_fast; the oracle row checks are exact-equality either way)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
from curvis_b200.interpolation import Interpolator
from oracle import oracle as O

PRECISION_NAME = sys.argv[1] if len(sys.argv) > 1 else "f64_fast"
PREC = {"f64": _abi.PRECISION_F64, "f64_fast": _abi.PRECISION_F64_FAST, "f32": _abi.PRECISION_F32}[PRECISION_NAME]
NCPU = os.cpu_count() or 1
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
results = []


def check_rows(kind, cam_args, sim, frame, rows):
    ocam = O.camera(*cam_args)
    bad = 0
    for y in rows:
        ref, _, _ = O.render_rows(O.metric(kind), ocam, O.sim(*sim), bp, bn, row_begin=y, row_end=y + 1, threads=NCPU, with_records=False)
        bad += int((ref[0] != frame[y]).any(axis=1).sum())
    return bad


def run_frame(name, kind, W, H, sim, ctx, sample_rows):
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=ctx)
    sysm.render_image(*sim, precision=PREC)
    best = None
    for _ in range(3):
        t = time.perf_counter(); frame = sysm.render_image(*sim, precision=PREC); wall = (time.perf_counter() - t) * 1e3
        st = sysm.last_stats
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = dict(kernel_ms=st["kernel_ms"], wall_ms=wall)
    rows = list(range(sample_rows // 2, H, max(1, H // sample_rows)))[:sample_rows] if sample_rows < H else list(range(H))
    t = time.perf_counter(); bad = check_rows(kind, cam_args, sim, frame, rows); cpu_s = time.perf_counter() - t
    r = dict(config=name, precision=PRECISION_NAME, metric=kind, W=W, H=H, sim=sim, devices=ctx.device_count(), ray_steps=st["total_steps"], **best,
             ray_steps_per_s=st["total_steps"] / best["kernel_ms"] * 1e3, frames_per_s_e2e=1e3 / best["wall_ms"],
             escaped=[st["n_positive"], st["n_negative"], st["n_not_escaped"]], oracle_rows_checked=len(rows),
             pixels_checked=len(rows) * W, differing_pixels=bad, oracle_seconds=cpu_s)
    print(json.dumps(r), flush=True)
    results.append(r)


one = cv.Context([0])
run_frame("C1a", "ellis", 256, 144, (200, 10.0, 0.1), one, 144)
run_frame("C1b", "ellis", 256, 144, (40000, 100.0, 0.05), one, 144)
run_frame("C2", "ellis", 1920, 1080, (1000, 25.0, 0.05), one, 64)
run_frame("C3", "interstellar", 3840, 2160, (2000, 45.0, 0.05), one, 16)
run_frame("C3-defaults", "interstellar", 3840, 2160, (40000, 100.0, 0.05), one, 4)
allgpus = cv.Context()
run_frame("C4", "ellis", 7680, 4320, (40000, 100.0, 0.05), allgpus, 4)

# ---- C5: video
W, H, sim = 3840, 2160, (2000, 45.0, 0.05)
it = Interpolator.from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curvis_b200", "paths", "path_through.csv"))
times = [k / 15.0 for k in range(300)]
t_acc, times15 = 0.0, []
t = it.min_time()
while t < it.max_time():
    times15.append(t); t += 1.0 / 15.0
times = times15[:300]
cams = [cv.Camera(it.camera_position(t), it.camera_forward(t), it.camera_up(t), 15.0, 43.0, W, H) for t in times]
sysm = cv.RelativisticSystem(cv.InterstellarMetric(0.1, 1e-4, 1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cams[0], context=one)
B = 4
out = torch.empty(B * H * W * 3, dtype=torch.uint8, device="cuda:0")
host = torch.empty(B * H * W * 3, dtype=torch.uint8).pin_memory()
stream = torch.cuda.current_stream().cuda_stream
keep = {0: None, 149: None, 299: None}
torch.cuda.synchronize(); t0 = time.perf_counter(); steps = 0; kernel_ms = 0.0
for b0 in range(0, 300, B):
    st = sysm.render_frames_device(cams[b0:b0 + B], *sim, 0, H, out.data_ptr(), stream, want_stats=True, precision=PREC)
    steps += st["total_steps"]; kernel_ms += st["kernel_ms"]
    host.copy_(out); torch.cuda.synchronize()
    for f in range(B):
        if b0 + f in keep:
            keep[b0 + f] = host.numpy().reshape(B, H, W, 3)[f].copy()
wall = time.perf_counter() - t0
bad = 0; checked = 0
for idx, frame in keep.items():
    t = times[idx]
    cam_args = (it.camera_position(t), it.camera_forward(t), it.camera_up(t), 15.0, 43.0, W, H)
    rows = [100, 1080, 2000]
    bad += check_rows("interstellar", cam_args, sim, frame, rows); checked += len(rows) * W
r = dict(config="C5", precision=PRECISION_NAME, metric="interstellar", W=W, H=H, sim=sim, frames=300, devices=1, batch=B, ray_steps=steps, kernel_ms=kernel_ms,
         wall_s=wall, frames_per_s=300 / wall, ray_steps_per_s=steps / kernel_ms * 1e3, pixels_checked=checked, differing_pixels=bad,
         note="frames stay in host memory (no PNG encode); 8-GPU figure = this x the measured weak-scaling efficiency")
print(json.dumps(r), flush=True)
results.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open(f"gpurun_out/configs_r1_{PRECISION_NAME}.json", "w"), indent=1)
