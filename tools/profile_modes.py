"""Dev script (GPU box, under ncu): one 4K default frame per secondary kernel — Interstellar in CURVIS_PRECISION_F64_FAST,
Ellis and Interstellar in CURVIS_PRECISION_F32 — after one warm-up frame each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
for metric, prec in ((cv.InterstellarMetric(0.1, 1e-4, 1.0), _abi.PRECISION_F64_FAST), (cv.EllisMetric(1.0), _abi.PRECISION_F32),
                     (cv.InterstellarMetric(0.1, 1e-4, 1.0), _abi.PRECISION_F32)):
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    for _ in range(2):
        sysm.render_image(40000, 100.0, 0.05, precision=prec)
    print(type(metric).__name__, prec, sysm.last_stats["kernel_ms"], sysm.last_stats["total_steps"], flush=True)
