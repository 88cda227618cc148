"""Single-wave latency of the render kernels: one row of a 4K frame (3840 rays, ~2000 steps each) leaves every scheduler with
at most one warp, so kernel_ms / steps is the dependent-chain latency of one Euler step.  (GPU box)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(2048, 1024), scenes.decodable_background(2048, 1024, negative=True)
W, H = 3840, 2160
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
system = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
buf = torch.empty(64 * W * 3, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.current_stream()
sim = (40000, 100.0, 0.05)
out = {}
for label, prec, opts in (("f64_v4", _abi.PRECISION_F64, {"kernel_variant": 4}), ("f64_v3", _abi.PRECISION_F64, {"kernel_variant": 3}),
                          ("fast_guard0", _abi.PRECISION_F64_FAST, {"guard": 0}), ("fast_guard1", _abi.PRECISION_F64_FAST, {"guard": 1})):
    for k, v in opts.items():
        ctx.set_option(k, v)
    for rows in (1, 4, 16, 64):
        ms = []
        for _ in range(4):
            st = system.render_rows_device(*sim, 300, 300 + rows, buf.data_ptr(), stream.cuda_stream, want_stats=True, precision=prec)
            ms.append(st["kernel_ms"])
        steps_per_ray = st["total_steps"] / (rows * W)
        out[f"{label}_rows{rows}"] = {"kernel_ms": round(min(ms[1:]), 3), "steps_per_ray": round(steps_per_ray), "ns_per_step_latency": round(min(ms[1:]) * 1e6 / steps_per_ray, 1),
                                      "n_reintegrated": int(st["n_reintegrated"])}
    ctx.set_option("kernel_variant", 4); ctx.set_option("guard", 1)
for label, prec in (("f64_v4", _abi.PRECISION_F64), ("fast_guard0", _abi.PRECISION_F64_FAST), ("fast_guard1", _abi.PRECISION_F64_FAST)):
    ctx.set_option("guard", 0 if label.endswith("0") else 1)
    for row in (300, 1079, 1080, 1081):
        ms = []
        for _ in range(4):
            st = system.render_rows_device(*sim, row, row + 1, buf.data_ptr(), stream.cuda_stream, want_stats=True, precision=prec)
            ms.append(st["kernel_ms"])
        out[f"{label}_row{row}"] = {"kernel_ms": round(min(ms[1:]), 3), "steps_per_ray": round(st["total_steps"] / W), "n_reintegrated": int(st["n_reintegrated"]), "n_kicked": int(st["n_kicked"])}
print(json.dumps({k: v for k, v in out.items() if "_row" in k and "rows" not in k}, indent=0))
