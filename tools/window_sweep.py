"""Dev script (GPU box): kernel time of CURVIS_PRECISION_F64_FAST vs the "window" / "blocks_per_sm" options, 4K Ellis."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
kind = sys.argv[1] if len(sys.argv) > 1 else "ellis"
metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
sim = (40000, 100.0, 0.05)
for bps in (0, 5):
    for window in (32, 64, 96, 128, 192, 256):
        ctx.set_option("window", window); ctx.set_option("blocks_per_sm", bps)
        ms = []
        for _ in range(3):
            sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST); ms.append(sysm.last_stats["kernel_ms"])
        print(json.dumps(dict(kind=kind, blocks_per_sm=bps, window=window, kernel_ms=min(ms))), flush=True)
