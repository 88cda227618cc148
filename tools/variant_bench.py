"""Dev script (GPU box): time kernel variants / occupancy / window on the 4K Ellis frame and
check they all produce the same frame."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes

W, H = 3840, 2160
sim = (40000, 100.0, 0.05)
kind = sys.argv[1] if len(sys.argv) > 1 else "ellis"
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
ctx = cv.Context([0])
metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn),
                             cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H), context=ctx)
base = None
configs = [(0, 0, 16), (2, 0, 16), (3, 0, 16), (3, 0, 8), (3, 0, 32), (3, 0, 64), (3, 6, 16), (3, 5, 16), (3, 4, 16)]
for variant, blocks, window in configs:
    ctx.set_option("kernel_variant", variant); ctx.set_option("blocks_per_sm", blocks); ctx.set_option("window", window)
    frame = sysm.render_image(*sim)
    ms = []
    for _ in range(3):
        sysm.render_image(*sim); ms.append(sysm.last_stats["kernel_ms"])
    st = sysm.last_stats
    if base is None: base = frame
    print(json.dumps(dict(kind=kind, variant=variant, blocks_per_sm=blocks, window=window, kernel_ms=min(ms),
                          gsteps_per_s=st["total_steps"] / min(ms) / 1e6, same_as_v0=bool((frame == base).all()))), flush=True)
