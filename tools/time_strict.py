"""Kernel times of CURVIS_PRECISION_F64 on the 4K default frames (GPU box): python tools/time_strict.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
W, H = 3840, 2160
frame = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.current_stream()
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
for name, metric, sim in (("ellis", cv.EllisMetric(1.0), (40000, 100.0, 0.05)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0), (40000, 100.0, 0.05)),
                          ("interstellar c3", cv.InterstellarMetric(0.1, 1e-4, 1.0), (2000, 45.0, 0.05))):
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    for variant in (4, 5):
        ctx.set_option("kernel_variant", variant)
        ms = [system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64)["kernel_ms"] for _ in range(4)]
        print(name, "variant", variant, [round(m, 2) for m in ms], flush=True)
