"""How many of the chaotic rays of the reference's scheme become regular under the extensions (GPU box):
    python tools/extensions_report.py [out.json]
Default 1080p frames (Ellis, Interstellar): the reference's fixed-step Euler in (theta, phi); the same with the pole-adaptive
step at several tolerances; the chart-free angular state.  Classifier: oracle/classify.py (|p_l| > 1.05 or min |sin theta| <
1e-3) applied to the GPU's own records (the parity tests hold them identical to the oracle's); for the chart-free mode, which
has no theta along the trajectory, the |p_l| criterion alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import _abi, scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
W, H, sim = 1920, 1080, (40000, 100.0, 0.05)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
out = {}
for name, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    res = {}
    base_rgb = None
    for label, opts in (("euler (reference)", {}), ("adaptive tol 0.1", dict(integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=0.1)),
                        ("adaptive tol 0.03", dict(integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=0.03)),
                        ("adaptive tol 0.01", dict(integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=0.01)),
                        ("cartesian", dict(coordinates=_abi.COORDINATES_CARTESIAN)), ("rk4 delta 0.25", dict(integrator=_abi.INTEGRATOR_RK4))):
        s = (40000, 100.0, 0.25) if label.startswith("rk4") else sim
        rgb, rec = system.render_rows(*s, 0, H, with_records=True, **opts)
        st = dict(system.last_stats)
        esc = rec["side"] != 0
        big_pl = np.abs(rec["p_l"]) > 1.05
        with np.errstate(invalid="ignore"):
            small_sin = rec["min_abs_sin_theta"] < 1e-3
            kicked = ~(rec["stiffness"] < 1.0) & np.isfinite(rec["stiffness"])
        chaotic = (big_pl | small_sin) & esc
        ms = []
        for _ in range(3):
            system.render_image(*s, **opts)
            ms.append(system.last_stats["kernel_ms"])
        if base_rgb is None:
            base_rgb = rgb
        res[label] = {"chaotic_fraction": float(chaotic.mean()), "abs_p_l_above_1.05": float((big_pl & esc).mean()),
                      "min_sin_below_1e-3": float(small_sin.mean()) if np.isfinite(rec["min_abs_sin_theta"]).any() else None,
                      "kicked_stiffness_ge_1": float(kicked.mean()) if np.isfinite(rec["stiffness"]).any() else None,
                      "total_steps": int(st["total_steps"]), "kernel_ms_1080p": round(min(ms), 3), "not_escaped": int(st["n_not_escaped"]),
                      "pixels_equal_to_the_reference_scheme": float((rgb == base_rgb).all(axis=2).mean())}
    out[name] = res
    print(name, json.dumps(res), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
