"""A/B of two builds on one box: python tools/time_ab.py <package root> (kernel ms of the 4K default frames, guard 0 / 1)."""
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes
print(cv.__file__)
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
W, H = 3840, 2160
frame = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.current_stream()
for mname, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    for guard in (0, 1):
        ctx.set_option("guard", guard)
        ms = []
        for _ in range(6):
            st = system.render_rows_device(40000, 100.0, 0.05, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)
            ms.append(round(st["kernel_ms"], 3))
        print(mname, "guard", guard, ms, flush=True)
