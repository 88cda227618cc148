"""A/B of the re-integration launch's step loop (ctx option redo_ahead: 1 = latency form, 0 = the frame kernel's throughput form):
kernel time of CURVIS_PRECISION_F64_FAST with the guard band on one row, on one rank's tile of a frame split over 8, and on whole 4K
frames (guard 1 and 2), and that both forms give the same bytes.   python tools/redo_ahead_ab.py [out.json]"""
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
frames = [torch.zeros(H * W * 3, dtype=torch.uint8, device="cuda:0") for _ in range(2)]
out = {}
for mname, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)

    def run(case, frame, reps=4):
        ms = []
        for _ in range(reps):
            if case == "row300":
                st = system.render_rows_device(*sim, 300, 301, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)
            elif case == "frame":
                st = system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)
            else:
                st = system.render_frames_peers([cam], *sim, case, H, [frame.data_ptr()], stream.cuda_stream, want_stats=True, row_stride=8, precision=_abi.PRECISION_F64_FAST)
            ms.append(st["kernel_ms"])
        return round(min(ms[1:]), 3), int(st["n_reintegrated"])
    res = {}
    for guard in (1, 2):
        ctx.set_option("guard", guard)
        for case in ("row300", 7, 1, 0, "frame"):
            row = {}
            for ahead in (0, 1):
                ctx.set_option("redo_ahead", ahead)
                frames[ahead].zero_()
                row[f"ahead{ahead}_ms"], row["n_reintegrated"] = run(case, frames[ahead])
            torch.cuda.synchronize()
            row["differing_bytes"] = int((frames[0] != frames[1]).sum().item())
            res[f"guard{guard}_{case if isinstance(case, str) else 'tile%d_of_8' % case}"] = row
    ctx.set_option("guard", 0)
    res["guard0_frame_ms"] = run("frame", frames[0])[0]
    ctx.set_option("guard", 1); ctx.set_option("redo_ahead", 1)
    out[mname] = res
    print(mname, flush=True)
    for k, v in res.items():
        print("  ", k, json.dumps(v), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
