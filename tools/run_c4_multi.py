"""BASELINE.json configs[3] on the GPUs of one box, one rank per GPU (launch with torchrun): ONE Ellis 7680x4320 frame
(default settings), (a) as written — contiguous row tiles + a single NCCL all-gather of the frame — and (b) with
interleaved rows + the fused peer stores (no collective for the pixels).  Rows are checked against the CPU oracle.
Writes gpurun_out/c4_multi.json on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
from curvis_b200.distributed import interleaved_rows, row_tile

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
W, H, sim, REPS = 7680, 4320, (40000, 100.0, 0.05), 5
PREC = _abi.PRECISION_F64_FAST
frame_bytes = W * H * 3
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
ctx = cv.Context([local])
sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=ctx)
stream = torch.cuda.current_stream()

def timed(fn):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / REPS], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

# (a) contiguous tiles + one NCCL all-gather
b, e = row_tile(H, rank, world)
tile = torch.empty((e - b) * W * 3, dtype=torch.uint8, device=dev)
frame_a = torch.empty(frame_bytes, dtype=torch.uint8, device=dev)
def nccl_frame():
    sysm.render_rows_device(*sim, b, e, tile.data_ptr(), stream.cuda_stream, precision=PREC)
    dist.all_gather_into_tensor(frame_a, tile)
ms_a = timed(nccl_frame)
st = sysm.render_rows_device(*sim, b, e, tile.data_ptr(), stream.cuda_stream, want_stats=True, precision=PREC)
steps = torch.tensor([st["total_steps"]], dtype=torch.int64, device=dev); dist.all_reduce(steps)
kernel_ms = torch.tensor([st["kernel_ms"]], dtype=torch.float64, device=dev)
kmax, kmin = kernel_ms.clone(), kernel_ms.clone()
dist.all_reduce(kmax, op=dist.ReduceOp.MAX); dist.all_reduce(kmin, op=dist.ReduceOp.MIN)

# (b) interleaved rows + fused peer stores
own = cv.PeerBuffer.create(ctx, frame_bytes)
handles = [None] * world
dist.all_gather_object(handles, own.handle)
peers = [own if r == rank else cv.PeerBuffer.open(ctx, handles[r], frame_bytes) for r in range(world)]
frame_b = own.as_tensor(local)
token = torch.zeros(1, dtype=torch.int32, device=dev)
r0, r1, stride = interleaved_rows(H, rank, world)
cam = [sysm.camera]
def peers_frame():
    sysm.render_frames_peers(cam, *sim, r0, r1, [p.ptr for p in peers], stream.cuda_stream, row_stride=stride, precision=PREC)
    dist.all_reduce(token)
ms_b = timed(peers_frame)
same = bool((frame_a == frame_b).all().item())

from oracle import oracle as O
rows = [rank * (H // world) + 7, rank * (H // world) + 333]
fa = frame_a.view(H, W, 3)
bad = 0
for y in rows:
    ref, _, _ = O.render_rows(O.metric("ellis"), O.camera(*cam_args), O.sim(*sim), bp, bn, row_begin=y, row_end=y + 1, threads=os.cpu_count() or 1, with_records=False)
    bad += int((torch.from_numpy(ref[0]).to(dev) != fa[y]).any(dim=1).sum().item())
stat = torch.tensor([bad, len(rows) * W, int(same)], dtype=torch.int64, device=dev)
dist.all_reduce(stat)
if rank == 0:
    total = int(steps.item())
    r = dict(config="C4", precision="f64_fast", metric="ellis", W=W, H=H, sim=sim, gpus=world, ray_steps=total,
             nccl_all_gather=dict(ms_per_frame=ms_a, ray_steps_per_s=total / ms_a * 1e3, frames_per_s=1e3 / ms_a,
                                  tile_kernel_ms_min=float(kmin.item()), tile_kernel_ms_max=float(kmax.item()),
                                  note="contiguous row tiles (4320/N rows per rank) + one all_gather_into_tensor of the 99.5 MB frame"),
             fused_peer_stores=dict(ms_per_frame=ms_b, ray_steps_per_s=total / ms_b * 1e3, frames_per_s=1e3 / ms_b,
                                    note="interleaved rows, pixels stored into every rank's frame over NVLink, 4-byte all-reduce as barrier"),
             frames_identical_on_all_ranks=bool(stat[2].item() == world), oracle_pixels_checked=int(stat[1].item()), differing_pixels=int(stat[0].item()))
    print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(r, open("gpurun_out/c4_multi.json", "w"), indent=1)
torch.cuda.synchronize()
del frame_b
for r_, p in enumerate(peers):
    if r_ != rank: p.close()
dist.barrier()
own.close()
dist.destroy_process_group()
