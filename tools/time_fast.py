"""Kernel times of the F64_FAST kernel on the 4K default frames, the knobs that matter (GPU box): python tools/time_fast.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
W, H = 3840, 2160
frame = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.current_stream()
out = {}
for mname, metric, sim in (("ellis_defaults", cv.EllisMetric(1.0), (40000, 100.0, 0.05)), ("interstellar_defaults", cv.InterstellarMetric(0.1, 1e-4, 1.0), (40000, 100.0, 0.05)),
                           ("interstellar_c3", cv.InterstellarMetric(0.1, 1e-4, 1.0), (2000, 45.0, 0.05))):
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)

    def run(reps=5, rows=(0, H), **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        ms = []
        for _ in range(reps):
            st = system.render_rows_device(*sim, rows[0], rows[1], frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)
            ms.append(st["kernel_ms"])
        return {"kernel_ms": round(min(ms[1:]), 3), "n_reintegrated": int(st["n_reintegrated"]), "total_steps": int(st["total_steps"])}
    res = {}
    for regs in (96, 128):
        for guard in (0, 1):
            for lf in (0, 1):
                res[f"regs{regs}_guard{guard}_lf{lf}"] = run(fast_regs=regs, guard=guard, longest_first=lf)
    ctx.set_option("fast_regs", 0); ctx.set_option("guard", 1)
    # an eighth of the frame around the central rows (what one of 8 ranks would render with contiguous tiles): the stragglers' tile
    for lf in (0, 1):
        res[f"rows945_1215_guard1_lf{lf}"] = run(rows=(945, 1215), longest_first=lf)
    ctx.set_option("longest_first", 2)
    out[mname] = res
    print(mname, json.dumps(res), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
