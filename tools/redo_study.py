"""Which rays does the guard band of CURVIS_PRECISION_F64_FAST send to the re-integration launch, and how long are they?
(GPU box)  python tools/redo_study.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import _abi, scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
W, H = 3840, 2160
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
for name, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    sim = (40000, 100.0, 0.05)
    rows = (0, H // 2)         # half a frame: 4.1 M rays, 330 MB of records
    ctx.set_option("guard", 0)
    _, raw = system.render_rows(*sim, *rows, with_records=True, precision=_abi.PRECISION_F64_FAST)
    ctx.set_option("guard", 1)
    _, grd = system.render_rows(*sim, *rows, with_records=True, precision=_abi.PRECISION_F64_FAST)
    st = dict(system.last_stats)
    redone = (raw["l"] != grd["l"]) | (raw["theta"] != grd["theta"]) | (raw["p_l"] != grd["p_l"]) | (raw["steps"] != grd["steps"])
    steps = grd["steps"][redone]
    k = raw["stiffness"][redone]
    out = {"metric": name, "rays": int(redone.size), "n_reintegrated_reported": int(st["n_reintegrated"]), "identified": int(redone.sum()),
           "steps_quantiles": {q: float(np.quantile(steps, q)) for q in (0.5, 0.9, 0.99, 1.0)} if steps.size else None,
           "not_escaped_among_them": int((grd["side"][redone] == 0).sum()),
           "stiffness_quantiles": {q: float(np.quantile(k, q)) for q in (0.1, 0.5, 0.9, 1.0)} if k.size else None,
           "abs_sin_final_quantiles": {q: float(np.quantile(np.abs(np.sin(grd["theta"][redone])), q)) for q in (0.01, 0.1, 0.5)} if k.size else None,
           "all_rays_steps_max": int(grd["steps"].max()), "all_rays_not_escaped": int((grd["side"] == 0).sum())}
    print(json.dumps(out), flush=True)
