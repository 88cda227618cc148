"""Dev script (GPU box): A/B of the CURVIS_PRECISION_F64_FAST variants (ctx option "fast_variant")
against the fp64 parity kernel on full 4K frames: differing pixels, step totals, kernel time.
Writes gpurun_out/fast_variant_check.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
out = []
kinds = sys.argv[1:] or ["ellis", "interstellar", "flat"]
for kind in kinds:
    metric = {"ellis": cv.EllisMetric(1.0), "interstellar": cv.InterstellarMetric(0.1, 1e-4, 1.0), "flat": cv.FlatSphericalMetric()}[kind]
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    sim = (40000, 100.0, 0.05) if kind != "flat" else (4000, 100.0, 0.05)
    ref = sysm.render_image(*sim, precision=_abi.PRECISION_F64).copy()
    ref_stats = dict(sysm.last_stats)
    res = dict(kind=kind, f64_kernel_ms=ref_stats["kernel_ms"], total_steps=ref_stats["total_steps"])
    for variant in (0, 1):
        for window in ((32, 64) if variant == 0 else (0, 64, 128, 256)):
            ctx.set_option("fast_variant", variant)
            ctx.set_option("window", window)
            ms = []
            for _ in range(3):
                fr = sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST)
                ms.append(sysm.last_stats["kernel_ms"])
            st = sysm.last_stats
            ys, xs = np.nonzero((fr != ref).any(axis=2))
            res[f"variant{variant}_window{window}"] = dict(
                kernel_ms=min(ms), ray_steps_per_s=st["total_steps"] / min(ms) * 1e3,
                differing_pixels=int(len(xs)), first=[(int(a), int(b)) for a, b in zip(xs[:8], ys[:8])],
                total_steps_equal=bool(st["total_steps"] == ref_stats["total_steps"]),
                counters_equal=all(st[k] == ref_stats[k] for k in ("n_positive", "n_negative", "n_not_escaped", "n_clamped")))
    ctx.set_option("window", 0)
    ctx.set_option("fast_variant", 1)
    print(json.dumps(res), flush=True)
    out.append(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fast_variant_check.json", "w"), indent=1)
