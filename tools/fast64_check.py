"""Dev script (GPU box): CURVIS_PRECISION_F64_FAST against the fp64 parity kernel — op-level
accuracy of its primitives, whole-frame deviation (pixels, steps, texels, state), timing.
Writes gpurun_out/fast64_check.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi

out = {}
ctx = cv.Context([0])
rng = np.random.default_rng(20251017)

# ---- op level
def ulps(got, want_ld):
    want = want_ld.astype(np.float64)
    ulp = np.spacing(np.abs(want))
    return np.abs((got.astype(np.longdouble) - want_ld) / ulp.astype(np.longdouble)).astype(np.float64)

x = np.exp(rng.uniform(-60, 60, 4_000_000)) * rng.choice([1.0], 4_000_000)
e = ulps(ctx.debug_eval(10, x), 1.0 / x.astype(np.longdouble))
out["rcp_1ulp"] = dict(max_ulp=float(e.max()), mean_ulp=float(e.mean()), frac_exact=float((ctx.debug_eval(10, x) == 1.0 / x).mean()))
th = np.concatenate([rng.uniform(-7, 7, 4_000_000), rng.uniform(-1e-3, 1e-3, 100_000) + np.pi, rng.uniform(-1e-6, 1e-6, 100_000)])
thl = th.astype(np.longdouble)
s2 = ctx.debug_eval(11, th); cs = ctx.debug_eval(12, th)
e2 = ulps(s2, np.sin(thl) ** 2); ec = ulps(cs, np.sin(thl) * np.cos(thl))
out["sin2"] = dict(max_ulp=float(e2.max()), mean_ulp=float(e2.mean()))
out["sincos_product"] = dict(max_ulp=float(ec.max()), mean_ulp=float(ec.mean()))
print(json.dumps(out), flush=True)

# ---- frames
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)

def compare(kind, W, H, sim, rows=None):
    metric = {"ellis": cv.EllisMetric(1.0), "interstellar": cv.InterstellarMetric(0.1, 1e-4, 1.0), "flat": cv.FlatSphericalMetric()}[kind]
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    r0, r1 = rows if rows else (0, H)
    f64, r64 = sysm.render_rows(*sim, r0, r1, with_records=True)
    s64 = sysm.last_stats
    ff, rf = sysm.render_rows(*sim, r0, r1, with_records=True, precision=_abi.PRECISION_F64_FAST)
    sf = sysm.last_stats
    rel = lambda k: float(np.nanmax(np.abs(r64[k] - rf[k]) / np.maximum(np.abs(r64[k]), 1e-300)))
    same_steps = r64["steps"] == rf["steps"]
    res = dict(kind=kind, W=W, H=H, rows=[r0, r1], sim=sim, rays=int(r64.size),
               differing_pixels=int((f64 != ff).any(axis=2).sum()),
               side_differs=int((r64["side"] != rf["side"]).sum()), steps_differ=int((~same_steps).sum()),
               texel_differs=int(((r64["texel_x"] != rf["texel_x"]) | (r64["texel_y"] != rf["texel_y"])).sum()),
               total_steps_equal=bool(s64["total_steps"] == sf["total_steps"]),
               counters_equal=all(s64[k] == sf[k] for k in ("n_positive", "n_negative", "n_not_escaped", "n_clamped")),
               max_rel_l=rel("l"), max_rel_pl=rel("p_l"),
               median_abs_dtheta=float(np.nanmedian(np.abs(r64["theta"] - rf["theta"])[same_steps])),
               p999_abs_dtheta=float(np.nanpercentile(np.abs(r64["theta"] - rf["theta"])[same_steps], 99.9)),
               max_abs_dtheta=float(np.nanmax(np.abs(r64["theta"] - rf["theta"])[same_steps])),
               max_abs_dphi=float(np.nanmax(np.abs(r64["phi"] - rf["phi"])[same_steps])),
               kernel_ms_f64=s64["kernel_ms"], kernel_ms_fast=sf["kernel_ms"])
    print(json.dumps(res), flush=True)
    return res

out["frames"] = [
    compare("ellis", 256, 144, (40000, 100.0, 0.05)),
    compare("ellis", 256, 144, (200, 10.0, 0.1)),
    compare("interstellar", 256, 144, (40000, 100.0, 0.05)),
    compare("flat", 256, 144, (4000, 100.0, 0.05)),
    compare("ellis", 1920, 1080, (40000, 100.0, 0.05)),
    compare("interstellar", 1920, 1080, (40000, 100.0, 0.05)),
]

# ---- full 4K frames, no records: differing pixels + timing
out["full_4k"] = []
for kind in ("ellis", "interstellar"):
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    frames, res = {}, dict(kind=kind)
    for prec, name in ((0, "f64"), (2, "f64_fast"), (1, "f32")):
        ms = []
        for _ in range(3):
            frames[name] = sysm.render_image(40000, 100.0, 0.05, precision=prec).copy(); ms.append(sysm.last_stats["kernel_ms"])
        st = sysm.last_stats
        res[name] = dict(kernel_ms=min(ms), ray_steps_per_s=st["total_steps"] / min(ms) * 1e3, total_steps=st["total_steps"])
    res["differing_pixels_fast_vs_f64"] = int((frames["f64"] != frames["f64_fast"]).any(axis=2).sum())
    res["differing_pixels_f32_vs_f64"] = int((frames["f64"] != frames["f32"]).any(axis=2).sum())
    ys, xs = np.nonzero((frames["f64"] != frames["f64_fast"]).any(axis=2))
    res["first_differences"] = [(int(a), int(b)) for a, b in zip(xs[:20], ys[:20])]
    print(json.dumps(res), flush=True)
    out["full_4k"].append(res)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fast64_check.json", "w"), indent=1)
