"""Dev script (GPU box): parity detail of the fp64 kernel vs the oracle + first timings."""
import json, math, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes
from oracle import oracle as O

def system_for(metric, W, H, bp, bn, ctx):
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP,
                    scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, W, H)
    return cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)

def compare(name, kind, sim, W=256, H=144):
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    ctx = cv.Context([0])
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    sysm = system_for(metric, W, H, bp, bn, ctx)
    frame, rec = sysm.render_rows(*sim, 0, H, with_records=True)
    st = sysm.last_stats
    cam = O.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP,
                   scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, W, H)
    ref, rrec, rst = O.render_rows(O.metric(kind), cam, O.sim(*sim), bp, bn, threads=os.cpu_count())
    same_px = (frame == ref).all(axis=2)
    same_steps = rec["steps"] == rrec["steps"]
    same_side = rec["side"] == rrec["side"]
    bit = np.ones_like(same_px)
    for f in ("l", "theta", "phi", "p_l", "p_theta", "p_phi"):
        bit &= (rec[f].view(np.uint64) == rrec[f].view(np.uint64))
    chaotic = (np.abs(rrec["p_l"]) > 1.05)
    res = dict(name=name, pixels=W * H, identical_rgb=float(same_px.mean()), identical_steps=float(same_steps.mean()),
               identical_side=float(same_side.mean()), bit_identical_state=float(bit.mean()),
               chaotic_frac=float(chaotic.mean()),
               identical_rgb_regular=float(same_px[~chaotic].mean()), identical_steps_regular=float(same_steps[~chaotic].mean()),
               gpu_steps=int(st["total_steps"]), oracle_steps=int(rst["total_steps"]), kernel_ms=st["kernel_ms"],
               gpu_counts=[st["n_positive"], st["n_negative"], st["n_not_escaped"], st["n_clamped"]],
               oracle_counts=[rst["n_positive"], rst["n_negative"], rst["n_not_escaped"], rst["n_clamped"]])
    print(json.dumps(res))
    return res

def timing(kind, W, H, sim, reps=3):
    bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
    ctx = cv.Context([0])
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    sysm = system_for(metric, W, H, bp, bn, ctx)
    for r in range(reps):
        t = time.time()
        sysm.render_image(*sim)
        wall = time.time() - t
        st = sysm.last_stats
        print(json.dumps(dict(kind=kind, W=W, H=H, sim=sim, kernel_ms=st["kernel_ms"], wall_ms=wall * 1e3,
                              steps=st["total_steps"], gsteps_per_s=st["total_steps"] / st["kernel_ms"] / 1e6,
                              counts=[st["n_positive"], st["n_negative"], st["n_not_escaped"], st["n_clamped"]])))

if __name__ == "__main__":
    out = []
    out.append(compare("C1a ellis 200/0.1/10", "ellis", (200, 10.0, 0.1)))
    out.append(compare("C1b ellis defaults", "ellis", (40000, 100.0, 0.05)))
    out.append(compare("interstellar defaults", "interstellar", (40000, 100.0, 0.05)))
    ctx = cv.Context([0])
    print("fma peak (fp64, fp32) TFLOP/s:", ctx.measure_fma_peak())
    timing("ellis", 1920, 1080, (1000, 25.0, 0.05))
    timing("ellis", 3840, 2160, (40000, 100.0, 0.05), reps=2)
    timing("interstellar", 3840, 2160, (2000, 45.0, 0.05), reps=2)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
