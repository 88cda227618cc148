"""Debug: pre-pass + cached table on a mid-size frame (run under compute-sanitizer on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(256, 128), scenes.decodable_background(256, 128, True)
W, H = 640, 360
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
for metric, sim in ((cv.EllisMetric(1.0), (300, 12.0, 0.1)), (cv.InterstellarMetric(0.1, 1e-4, 1.0), (300, 12.0, 0.1))):
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    for lf in (0, 1):
        ctx.set_option("longest_first", lf)
        for guard in (0, 1):
            ctx.set_option("guard", guard)
            try:
                img = sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST)
                print(type(metric).__name__, "lf", lf, "guard", guard, sysm.last_stats["total_steps"], sysm.last_stats["n_reintegrated"], flush=True)
            except Exception as e:
                print(type(metric).__name__, "lf", lf, "guard", guard, "FAILED", e, flush=True)
    ref = sysm.render_image(*sim)
    print("differing px vs F64:", int((ref != img).any(axis=-1).sum()), flush=True)
print("ok")
