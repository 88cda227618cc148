"""Measures what the guard band of CURVIS_PRECISION_F64_FAST has to cover (run on a GPU box):

    python tools/guard_study.py [out.json]

For a set of scenes it renders per-ray records with the operation-for-operation kernel (CURVIS_PRECISION_F64, with the
trajectory diagnostics) and with the raw regrouped kernel (CURVIS_PRECISION_F64_FAST, guard off), and reports, binned by
the ray's stiffness kappa = max (delta dphi/dlambda)^2:
  * how far the two end states are apart (direction of the lookup vector, l), also divided by the end-state factor
    (4 + 2 |d_z| / sin theta) the guard applies — i.e. the relative state error eps the band must exceed;
  * how many rays differ in step count / side / texel (the rays the guard must catch);
then the same frames with the guard on: differing rays must be 0, and the re-integrated fraction is the price.
Also times the 4K frames (guard on / off, 96 / 128 registers).
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv  # noqa: E402
from curvis_b200 import _abi, scenes  # noqa: E402

PI = math.pi


def shape_r(metric, l):
    if isinstance(metric, cv.EllisMetric):
        return np.sqrt(metric.rho ** 2 + l * l)
    al = np.abs(l)
    x = 2.0 * (al - metric.a) / (PI * metric.m)
    r = metric.rho + metric.m * (x * np.arctan(x) - np.log1p(x * x) / 2.0)
    return np.where(al > metric.a, r, metric.rho)


def lookup_direction(metric, rec):
    r = shape_r(metric, rec["l"])
    s = np.sin(rec["theta"])
    d = np.stack([rec["p_l"], rec["p_theta"] / r, rec["p_phi"] / (r * s * s)], axis=-1)
    return d, s


def study(name, metric, cam_args, sim, W, H, ctx, bp, bn, rows=None):
    cam = cv.Camera(*cam_args, scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    r0, r1 = rows if rows else (0, H)
    out = {"scene": name, "rays": (r1 - r0) * W}
    f64, rec64 = system.render_rows(*sim, r0, r1, with_records=True, precision=_abi.PRECISION_F64)
    st64 = dict(system.last_stats)
    ctx.set_option("guard", 0)
    raw, recraw = system.render_rows(*sim, r0, r1, with_records=True, precision=_abi.PRECISION_F64_FAST)
    ctx.set_option("guard", 1)
    grd, recg = system.render_rows(*sim, r0, r1, with_records=True, precision=_abi.PRECISION_F64_FAST)
    stg = dict(system.last_stats)

    kappa = rec64["stiffness"]
    same_steps = (rec64["steps"] == recraw["steps"]) & (rec64["side"] == recraw["side"])
    esc = same_steps & (rec64["side"] != 0)
    d64, s64 = lookup_direction(metric, rec64)
    dr, _ = lookup_direction(metric, recraw)
    n64 = np.linalg.norm(d64, axis=-1)
    dev_dir = np.linalg.norm(np.cross(d64, dr), axis=-1) / (n64 * np.linalg.norm(dr, axis=-1))
    factor = 4.0 + 2.0 * np.abs(d64[..., 2]) / (np.abs(s64) * n64)
    eps_dir = dev_dir / factor
    dev_l = np.abs(rec64["l"] - recraw["l"]) / (1.0 + np.abs(rec64["l"]))
    dev_th = np.abs(rec64["theta"] - recraw["theta"]) / (1.0 + np.abs(rec64["theta"]))
    dev_pl = np.abs(rec64["p_l"] - recraw["p_l"])
    eps_state = np.maximum.reduce([dev_l, dev_th, dev_pl])
    texel_diff = (rec64["texel_x"] != recraw["texel_x"]) | (rec64["texel_y"] != recraw["texel_y"])
    pix_diff_raw = (f64 != raw).any(axis=2)
    pix_diff_guard = (f64 != grd).any(axis=2)
    rec_diff_guard = (rec64["steps"] != recg["steps"]) | (rec64["side"] != recg["side"]) | \
                     (rec64["texel_x"] != recg["texel_x"]) | (rec64["texel_y"] != recg["texel_y"])
    chaotic = (np.abs(rec64["p_l"]) > 1.05) | (rec64["min_abs_sin_theta"] < 1e-3)
    out.update({
        "total_steps": int(st64["total_steps"]),
        "chaotic_fraction": float(chaotic.mean()),
        "raw": {"steps_or_side_differ": int((~same_steps).sum()), "texel_differ_same_steps": int((texel_diff & same_steps).sum()),
                "pixels_differ": int(pix_diff_raw.sum()), "pixels_differ_regular": int((pix_diff_raw & ~chaotic).sum())},
        "guarded": {"pixels_differ": int(pix_diff_guard.sum()), "records_differ": int(rec_diff_guard.sum()),
                    "n_reintegrated": int(stg["n_reintegrated"]), "reintegrated_fraction": stg["n_reintegrated"] / ((r1 - r0) * W),
                    "total_steps_equal": bool(stg["total_steps"] == st64["total_steps"])},
        "stiffness_vs_fast_kernel": {
            "max_rel_dev": float(np.nanmax(np.abs(recraw["stiffness"][same_steps] / np.maximum(kappa[same_steps], 1e-300) - 1.0))) if same_steps.any() else None},
    })
    bins = []
    edges = [0.0] + [10.0 ** e for e in range(-8, 17)] + [np.inf]
    for lo, hi in zip(edges[:-1], edges[1:]):
        m = (kappa >= lo) & (kappa < hi)
        if not m.any():
            continue
        me = m & esc
        row = {"kappa": [lo, hi if np.isfinite(hi) else None], "rays": int(m.sum()), "steps_differ": int((m & ~same_steps).sum()),
               "texel_differ": int((m & same_steps & texel_diff).sum())}
        if me.any():
            row.update({"dir_dev_max": float(dev_dir[me].max()), "dir_dev_p99": float(np.quantile(dev_dir[me], 0.99)),
                        "eps_dir_max": float(eps_dir[me].max()), "eps_dir_p999": float(np.quantile(eps_dir[me], 0.999)),
                        "eps_state_max": float(eps_state[me].max()), "eps_state_p999": float(np.quantile(eps_state[me], 0.999)),
                        "dev_l_max": float(np.abs(rec64["l"] - recraw["l"])[me].max()), "p_l_abs_max": float(np.abs(rec64["p_l"][me]).max()),
                        "dev_l_per_step_length_max": float((np.abs(rec64["l"] - recraw["l"]) / (np.abs(rec64["p_l"]) * abs(sim[2]) + 1e-300))[me].max()),
                        "dev_pl_rel_max": float((np.abs(rec64["p_l"] - recraw["p_l"]) / (np.abs(rec64["p_l"]) + 1e-300))[me].max())})
        bins.append(row)
    out["by_stiffness"] = bins
    # the rays the raw kernel got wrong: where do they sit?
    wrong = (~same_steps) | texel_diff
    if wrong.any():
        idx = np.argwhere(wrong)[:20]
        out["raw_wrong_examples"] = [{"row": int(i) + r0, "col": int(j), "kappa": float(kappa[i, j]), "min_abs_sin": float(rec64["min_abs_sin_theta"][i, j]),
                                      "steps": [int(rec64["steps"][i, j]), int(recraw["steps"][i, j])], "p_l": float(rec64["p_l"][i, j]),
                                      "dir_dev": float(dev_dir[i, j])} for i, j in idx]
    if rec_diff_guard.any() or pix_diff_guard.any():
        idx = np.argwhere(rec_diff_guard | pix_diff_guard)[:20]
        out["guard_missed_examples"] = [{"row": int(i) + r0, "col": int(j), "kappa": float(kappa[i, j]), "kappa_fast": float(recraw["stiffness"][i, j]),
                                         "min_abs_sin": float(rec64["min_abs_sin_theta"][i, j]),
                                         "steps": [int(rec64["steps"][i, j]), int(recg["steps"][i, j])], "dir_dev": float(dev_dir[i, j]),
                                         "eps_dir": float(eps_dir[i, j]), "eps_state": float(eps_state[i, j]),
                                         "texel64": [int(rec64["texel_x"][i, j]), int(rec64["texel_y"][i, j])],
                                         "texelg": [int(recg["texel_x"][i, j]), int(recg["texel_y"][i, j])]} for i, j in idx]
    return out


def timing(ctx, bp, bn):
    import torch
    out = {}
    W, H = 3840, 2160
    frame = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.current_stream()
    for mname, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
        cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, scenes.DEFAULT_FOCAL_LENGTH,
                        scenes.DEFAULT_DIAGONAL, W, H)
        system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
        sim = (40000, 100.0, 0.05)
        res = {}
        for regs, guard, redo in ((128, 0, 2), (96, 0, 2), (96, 1, 1), (96, 1, 2), (96, 1, 3), (96, 1, 5)):
            ctx.set_option("fast_regs", regs)
            ctx.set_option("guard", guard)
            ctx.set_option("redo_blocks_per_sm", redo)
            ms = []
            for _ in range(4):
                st = system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)
                ms.append(st["kernel_ms"])
            res[f"regs{regs}_guard{guard}_redo{redo}"] = {"kernel_ms": min(ms[1:]), "n_reintegrated": int(st["n_reintegrated"])}
        for variant in (3, 4):
            ctx.set_option("kernel_variant", variant)
            ms = []
            for _ in range(3):
                st = system.render_rows_device(*sim, 0, H, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64)
                ms.append(st["kernel_ms"])
            res[f"f64_strict_variant{variant}"] = {"kernel_ms": min(ms[1:])}
        ctx.set_option("fast_regs", 96)
        ctx.set_option("guard", 1)
        ctx.set_option("redo_blocks_per_sm", 2)
        out[mname] = res
    return out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/guard_study.json"
    ctx = cv.Context([0])
    bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
    dflt = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP)
    results = {"timing_4k": timing(ctx, bp, bn), "scenes": []}
    cases = [
        ("ellis_defaults_1080p", cv.EllisMetric(1.0), dflt, (40000, 100.0, 0.05), 1920, 1080, None),
        ("interstellar_defaults_1080p", cv.InterstellarMetric(0.1, 1e-4, 1.0), dflt, (40000, 100.0, 0.05), 1920, 1080, None),
        ("ellis_c2_1080p", cv.EllisMetric(1.0), dflt, (1000, 25.0, 0.05), 1920, 1080, None),
        ("interstellar_c3_4k_rows", cv.InterstellarMetric(0.1, 1e-4, 1.0), dflt, (2000, 45.0, 0.05), 3840, 2160, (900, 1260)),
        ("ellis_tilted_720p", cv.EllisMetric(2.0), ((0.0, -7.0, 1.1, 2.0), (1.0, 0.3, -0.2), (0.1, 0.0, 1.0)), (3000, 60.0, 0.05), 1280, 720, None),
        ("ellis_polar_camera_720p", cv.EllisMetric(1.0), ((0.0, 4.0, 0.3, 1.0), (-1.0, 0.1, 0.05), (0.0, 0.0, 1.0)), (40000, 100.0, 0.05), 1280, 720, None),
        ("ellis_c1a_256", cv.EllisMetric(1.0), dflt, (200, 10.0, 0.1), 256, 144, None),
        ("ellis_8k_rows", cv.EllisMetric(1.0), dflt, (40000, 100.0, 0.05), 7680, 4320, (2100, 2220)),
    ]
    for c in cases:
        results["scenes"].append(study(c[0], c[1], c[2], c[3], c[4], c[5], ctx, bp, bn, rows=c[6]))
        print(json.dumps({k: v for k, v in results["scenes"][-1].items() if k not in ("by_stiffness",)}), flush=True)
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(results, open(out_path, "w"), indent=1)
    print(json.dumps(results["timing_4k"]))


if __name__ == "__main__":
    main()
