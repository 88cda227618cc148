"""Dev script (GPU box): CURVIS_PRECISION_F32 on the full 4K frames — kernel time, differing pixels vs the parity
kernel, window sweep."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
sim = (40000, 100.0, 0.05)
for kind in ("ellis", "interstellar", "flat"):
    metric = {"ellis": cv.EllisMetric(1.0), "interstellar": cv.InterstellarMetric(0.1, 1e-4, 1.0), "flat": cv.FlatSphericalMetric()}[kind]
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    ref = sysm.render_image(*sim).copy(); st_ref = dict(sysm.last_stats)
    for window in (16, 32, 64):
        ctx.set_option("window", window)
        ms = []
        for _ in range(3):
            f = sysm.render_image(*sim, precision=_abi.PRECISION_F32); ms.append(sysm.last_stats["kernel_ms"])
        st = sysm.last_stats
        print(json.dumps(dict(kind=kind, window=window, kernel_ms=min(ms), ray_steps_per_s=st["total_steps"] / min(ms) * 1e3,
                              differing_pixels=int((f != ref).any(axis=2).sum()), dsteps=int(st["total_steps"]) - int(st_ref["total_steps"]),
                              counters=[st[k] - st_ref[k] for k in ("n_positive", "n_negative", "n_not_escaped")])), flush=True)
    ctx.set_option("window", 0)
