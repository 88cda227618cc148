"""ncu target: two 4K default frames of one F64_FAST kernel (first = warm-up): python tools/profile_one.py ellis|interstellar [guard]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
ctx = cv.Context([0])
ctx.set_option("guard", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
metric = cv.EllisMetric(1.0) if sys.argv[1] == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
for rep in range(2):
    system.render_image(40000, 100.0, 0.05, precision=_abi.PRECISION_F64_FAST)
    print(rep, system.last_stats["kernel_ms"], system.last_stats["total_steps"], flush=True)
