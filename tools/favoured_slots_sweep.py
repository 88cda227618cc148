"""Which hardware warp slots should claim the longest-first list?  Kernel times of CURVIS_PRECISION_F64_FAST (raw kernel, longest_first
= 1) against the ctx option "favoured_slots", whole 4K frames and one rank's tile of a frame split over 8 (rows 7::8; min / max of
8 launches).   python tools/favoured_slots_sweep.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
stream = torch.cuda.current_stream()
W, H = 3840, 2160
sim = (40000, 100.0, 0.05)
frame = torch.zeros(H * W * 3, dtype=torch.uint8, device="cuda:0")
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
out = {}
ctx.set_option("guard", 0); ctx.set_option("longest_first", 1)
for mname, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    res = {}
    for slots in (0, 4, 8, 12, 16, 64):
        ctx.set_option("favoured_slots", slots)
        fr = [system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)["kernel_ms"] for _ in range(5)][1:]
        tl = [system.render_frames_peers([cam], *sim, 7, H, [frame.data_ptr()], stream.cuda_stream, want_stats=True, row_stride=8, precision=_abi.PRECISION_F64_FAST)["kernel_ms"] for _ in range(9)][1:]
        res[f"slots{slots}"] = {"frame_ms_min": round(min(fr), 3), "frame_ms_max": round(max(fr), 3), "tile7_of_8_ms_min": round(min(tl), 3), "tile7_of_8_ms_max": round(max(tl), 3)}
    ctx.set_option("longest_first", 0)
    fr = [system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)["kernel_ms"] for _ in range(5)][1:]
    res["index_order_frame_ms_min"] = round(min(fr), 3)
    ctx.set_option("longest_first", 1)
    out[mname] = res
    print(mname, json.dumps(res), flush=True)
ctx.set_option("guard", 1); ctx.set_option("longest_first", 2); ctx.set_option("favoured_slots", 8)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
