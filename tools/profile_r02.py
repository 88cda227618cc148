"""Round-2 ncu target (GPU box, under ncu): one 4K default frame per kernel after one warm-up frame each —
  Ellis F64_FAST (main + re-integration launch), Interstellar F64_FAST (main + re-integration), Ellis F64 (kernel_variant 5, the default),
  Ellis chart-free coordinates.  6 warm-up launches, then the same 6 launches to capture (ncu -s 6 -c 6)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
ellis = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
for rep in range(2):
    for label, metric, opts in (("ellis fast", cv.EllisMetric(1.0), dict(precision=_abi.PRECISION_F64_FAST)),
                                ("interstellar fast", cv.InterstellarMetric(0.1, 1e-4, 1.0), dict(precision=_abi.PRECISION_F64_FAST)),
                                ("ellis f64", cv.EllisMetric(1.0), dict(precision=_abi.PRECISION_F64)),
                                ("ellis cartesian", cv.EllisMetric(1.0), dict(coordinates=_abi.COORDINATES_CARTESIAN))):
        ellis.metric = metric
        ellis.render_image(40000, 100.0, 0.05, **opts)
        print(rep, label, ellis.last_stats["kernel_ms"], ellis.last_stats["total_steps"], ellis.last_stats["n_reintegrated"], flush=True)
