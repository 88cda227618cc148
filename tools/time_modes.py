"""Kernel times of every fp64 mode on the 4K default frames (GPU box): python tools/time_modes.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
W, H = 3840, 2160
frame = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.current_stream()
out = {}
for mname, metric, sim in (("ellis_defaults", cv.EllisMetric(1.0), (40000, 100.0, 0.05)), ("interstellar_defaults", cv.InterstellarMetric(0.1, 1e-4, 1.0), (40000, 100.0, 0.05)),
                           ("interstellar_c3", cv.InterstellarMetric(0.1, 1e-4, 1.0), (2000, 45.0, 0.05))):
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)

    def run(precision, reps=4, **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        ms = []
        for _ in range(reps):
            st = system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=precision)
            ms.append(st["kernel_ms"])
        return {"kernel_ms": round(min(ms[1:]), 3), "n_reintegrated": int(st["n_reintegrated"]), "n_kicked": int(st["n_kicked"]), "total_steps": int(st["total_steps"]),
                "ray_steps_per_s": st["total_steps"] / (min(ms[1:]) * 1e-3)}
    res = {}
    for regs in (96, 128):
        for guard in (0, 1, 2):
            res[f"fast_regs{regs}_guard{guard}"] = run(_abi.PRECISION_F64_FAST, fast_regs=regs, guard=guard)
    ctx.set_option("fast_regs", 96); ctx.set_option("guard", 1)
    for redo in (1, 2, 3, 5):
        res[f"fast_guard1_redo{redo}"] = run(_abi.PRECISION_F64_FAST, redo_blocks_per_sm=redo)
    ctx.set_option("redo_blocks_per_sm", 2)
    res["fast_guard1_longest_first0"] = run(_abi.PRECISION_F64_FAST, longest_first=0)
    ctx.set_option("longest_first", 2)
    for variant in (3, 4):
        res[f"f64_variant{variant}"] = run(_abi.PRECISION_F64, reps=3, kernel_variant=variant)
    ctx.set_option("kernel_variant", 4)
    res["f32"] = run(_abi.PRECISION_F32)

    def run_cart(precision):
        ms = []
        for _ in range(4):
            system.render_image(*sim, precision=precision, coordinates=_abi.COORDINATES_CARTESIAN)
            ms.append(system.last_stats["kernel_ms"])
        st = system.last_stats
        return {"kernel_ms": round(min(ms[1:]), 3), "total_steps": int(st["total_steps"]), "ray_steps_per_s": st["total_steps"] / (min(ms[1:]) * 1e-3)}
    res["cartesian_f64"] = run_cart(_abi.PRECISION_F64)
    res["cartesian_f64_fast"] = run_cart(_abi.PRECISION_F64_FAST)
    out[mname] = res
    print(mname, json.dumps(res), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
