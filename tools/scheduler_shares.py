"""How evenly do the warp schedulers share a saturated fp64 pipe?  Every warp of the persistent render kernel lives as long as the
kernel, so the Euler steps a hardware warp slot (%warpid) executed over a whole 4K frame are the issue slots its scheduler gave it
(curvis_debug_last_step_shares).  Also: the spread of one rank's tile of a frame split over 8 (rows 7::8, which hold row 1079) over
20 launches — the tail of a small launch is whichever slot its longest rays landed in.   python tools/scheduler_shares.py [out.json]"""
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
system = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
out = {}
ctx.set_option("guard", 0)
for label, prec in (("f64_fast", _abi.PRECISION_F64_FAST), ("f64", _abi.PRECISION_F64)):
    for _ in range(2):
        st = system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=prec)
    slots, sms = ctx.last_step_shares()
    used = [s for s in slots if s]
    mean = sum(used) / len(used)
    sm_used = [s for s in sms if s]
    out[label] = {"kernel_ms": st["kernel_ms"], "warp_slots_used": len(used), "slot_share_of_mean": [round(s / mean, 3) for s in slots[:max(i for i, s in enumerate(slots) if s) + 1]],
                  "sm_share_min_max_of_mean": [round(min(sm_used) / (sum(sm_used) / len(sm_used)), 3), round(max(sm_used) / (sum(sm_used) / len(sm_used)), 3)], "sms": len(sm_used)}
    print(label, json.dumps(out[label]), flush=True)
for lf in (0, 1):
    ctx.set_option("longest_first", lf)
    ms = []
    for _ in range(21):
        st = system.render_frames_peers([cam], *sim, 7, H, [frame.data_ptr()], stream.cuda_stream, want_stats=True, row_stride=8, precision=_abi.PRECISION_F64_FAST)
        ms.append(round(st["kernel_ms"], 3))
    ms = sorted(ms[1:])
    out[f"tile7_of_8_guard0_lf{lf}_ms_sorted"] = ms
    print("tile 7 of 8, longest_first", lf, ms, flush=True)
for tile in (0, 1):
    ctx.set_option("longest_first", 1)
    ms = sorted(round(system.render_frames_peers([cam], *sim, tile, H, [frame.data_ptr()], stream.cuda_stream, want_stats=True, row_stride=8, precision=_abi.PRECISION_F64_FAST)["kernel_ms"], 3) for _ in range(12))
    out[f"tile{tile}_of_8_guard0_lf1_ms_sorted"] = ms
    print("tile", tile, "of 8, longest_first 1", ms, flush=True)
for lf in (0, 1):
    ctx.set_option("longest_first", lf)
    ms = sorted(round(system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)["kernel_ms"], 3) for _ in range(6))
    out[f"frame_guard0_lf{lf}_ms_sorted"] = ms
    print("whole frame, longest_first", lf, ms, flush=True)
ctx.set_option("guard", 1); ctx.set_option("longest_first", 2)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
