"""Dev script (GPU box): the FULL 4K Ellis default frame, every pixel, GPU vs the oracle on all
host cores.  Writes gpurun_out/parity_4k.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes
from oracle import oracle as O

W, H = 3840, 2160
sim = (40000, 100.0, 0.05)
kind = sys.argv[1] if len(sys.argv) > 1 else "ellis"
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
t0 = time.time()
ref, _, rst = O.render_rows(O.metric(kind), O.camera(*cam_args), O.sim(*sim), bp, bn, threads=os.cpu_count(), with_records=False)
t_cpu = time.time() - t0
out = dict(kind=kind, W=W, H=H, sim=sim, oracle_seconds=t_cpu, oracle_threads=os.cpu_count(), oracle_steps=rst["total_steps"], variants={})
ctx = cv.Context([0])
metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=ctx)
from curvis_b200 import _abi
for variant in (0, 3, "f64_fast", "f32"):
    if variant == "f64_fast":      # CURVIS_PRECISION_F64_FAST (render_f64_fast.cu)
        frame = sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST)
    elif variant == "f32":         # CURVIS_PRECISION_F32 (render_f32.cu): a tolerance mode, reported for scale
        frame = sysm.render_image(*sim, precision=_abi.PRECISION_F32)
    else:
        ctx.set_option("kernel_variant", variant)
        frame = sysm.render_image(*sim)
    st = sysm.last_stats
    diff = (frame != ref).any(axis=2)
    ys, xs = np.nonzero(diff)
    out["variants"][str(variant)] = dict(differing_pixels=int(diff.sum()), gpu_steps=int(st["total_steps"]),
                                         steps_equal=bool(st["total_steps"] == rst["total_steps"]),
                                         counts=[st["n_positive"], st["n_negative"], st["n_not_escaped"], st["n_clamped"]],
                                         oracle_counts=[rst["n_positive"], rst["n_negative"], rst["n_not_escaped"], rst["n_clamped"]],
                                         first_differences=[(int(x), int(y)) for x, y in zip(xs[:20], ys[:20])], kernel_ms=st["kernel_ms"])
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/parity_4k_{kind}.json", "w"), indent=1)
