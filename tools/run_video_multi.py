"""BASELINE.json configs[4] on the GPUs of one box, one rank per GPU (launch with torchrun):
video path_through.csv at 15 fps, frames 0..299, 3840x2160, Interstellar metric (m=0.1, a=1e-4, rho=1),
max_iter 2000 / delta 0.05 / R 45 (C5), per-pixel renderer in CURVIS_PRECISION_F64_FAST.

A step renders N consecutive frames: rank g renders rows g, g+N, ... of all N (one launch) and the kernel
stores every pixel into the complete frames of every rank over NVLink (curvis_render_frames_peers); after
the step barrier rank r reads frame r back to pinned host memory (the frame it would encode).  Rows of
three frames are checked against the CPU oracle.  Writes gpurun_out/video_multi.json on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_video_multi.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
from curvis_b200.distributed import interleaved_rows
from curvis_b200.interpolation import Interpolator

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)

W, H, sim, FRAMES = 3840, 2160, (2000, 45.0, 0.05), 300
frame_bytes = W * H * 3
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
it = Interpolator.from_file(os.path.join(root, "curvis_b200", "paths", "path_through.csv"))
times, t = [], it.min_time()
while t < it.max_time() and len(times) < FRAMES:
    times.append(t); t += 1.0 / 15.0
cams = [cv.Camera(it.camera_position(t), it.camera_forward(t), it.camera_up(t), 15.0, 43.0, W, H) for t in times]
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
ctx = cv.Context([local])
sysm = cv.RelativisticSystem(cv.InterstellarMetric(0.1, 1e-4, 1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cams[0], context=ctx)

sets, mine = [], []
for _ in range(2):
    own = cv.PeerBuffer.create(ctx, world * frame_bytes)
    handles = [None] * world
    dist.all_gather_object(handles, own.handle)
    sets.append([own if r == rank else cv.PeerBuffer.open(ctx, handles[r], world * frame_bytes) for r in range(world)])
    mine.append(own.as_tensor(local))
token = torch.zeros(1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream()
copy_stream = torch.cuda.Stream(device=dev)
host = torch.empty(frame_bytes, dtype=torch.uint8).pin_memory()
copied = torch.cuda.Event()
keep = {}                                   # frame index -> host copy, for the oracle check (frames this rank owns)
check_frames = (0, 149, 296)
r0, r1, stride = interleaved_rows(H, rank, world)

def step(k, b0):
    global copied
    batch = cams[b0:b0 + world]
    cur = sets[k & 1]
    sysm.render_frames_peers(batch, *sim, r0, r1, [b.ptr for b in cur], stream.cuda_stream, row_stride=stride, precision=_abi.PRECISION_F64_FAST)
    stream.wait_event(copied)
    dist.all_reduce(token)                  # step barrier
    if rank < len(batch):
        done = torch.cuda.Event(); done.record(stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            host.copy_(mine[k & 1][rank * frame_bytes:(rank + 1) * frame_bytes], non_blocking=True)
            copied = torch.cuda.Event(); copied.record(copy_stream)
        if b0 + rank in check_frames:
            copied.synchronize()
            keep[b0 + rank] = host.numpy().reshape(H, W, 3).copy()

step(0, 0); torch.cuda.synchronize(); dist.barrier()          # warm-up
t0 = time.perf_counter()
for k, b0 in enumerate(range(0, FRAMES, world)):
    step(k, b0)
torch.cuda.synchronize(); dist.barrier()
wall = time.perf_counter() - t0

from oracle import oracle as O
bad = checked = 0
for idx, frame in keep.items():
    tt = times[idx]
    ocam = O.camera(it.camera_position(tt), it.camera_forward(tt), it.camera_up(tt), 15.0, 43.0, W, H)
    for y in (100, 1080, 2000):
        ref, _, _ = O.render_rows(O.metric("interstellar"), ocam, O.sim(*sim), bp, bn, row_begin=y, row_end=y + 1, threads=os.cpu_count() or 1, with_records=False)
        bad += int((ref[0] != frame[y]).any(axis=1).sum()); checked += W
stat = torch.tensor([bad, checked], dtype=torch.int64, device=dev)
dist.all_reduce(stat)
if rank == 0:
    r = dict(config="C5", precision="f64_fast", metric="interstellar", W=W, H=H, sim=sim, frames=FRAMES, gpus=world, wall_s=wall,
             frames_per_s=FRAMES / wall, pixels_checked=int(stat[1].item()), differing_pixels=int(stat[0].item()),
             note="one rank per GPU, interleaved rows, pixels stored into every rank's frames over NVLink; rank r reads frame r back to pinned "
                  "host memory under the next step's render (no PNG encode)")
    print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(r, open("gpurun_out/video_multi.json", "w"), indent=1)
torch.cuda.synchronize()
for bufs in sets:
    for r_, b in enumerate(bufs):
        if r_ != rank: b.close()
dist.barrier()
for bufs in sets:
    bufs[rank].close()
dist.destroy_process_group()
