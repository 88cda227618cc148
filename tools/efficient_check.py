"""Dev script (GPU box): timing of the table-based renderer at several sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv
from curvis_b200 import scenes
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
ctx = cv.Context([0])
for kind in ("ellis", "interstellar"):
    for (W, H) in ((960, 540), (3840, 2160)):
        metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
        cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
        sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
        for _ in range(3):
            t = time.time(); sysm.render_image_efficient(40000, 100.0, 0.05, 100, 100, 1e-5, 1e-5); wall = (time.time() - t) * 1e3
        print(json.dumps(dict(kind=kind, W=W, H=H, wall_ms=wall, **sysm.last_efficient_info, black=sysm.last_stats["n_not_escaped"])), flush=True)
