"""Helpers shared by the -m gpu parity tests: the oracle on strided rows (all host cores, records with the trajectory
diagnostics), the same rows from the GPU with records, and the comparison split into regular / chaotic / kicked rays."""
from __future__ import annotations

import os

import numpy as np


def strided_rows(height: int, n_rows: int):
    stride = max(1, height // n_rows)
    first = stride // 2
    return first, stride, list(range(first, height, stride))


def oracle_rows(oracle, kind, metric_kwargs, cam_args, sim, bp, bn, first, stride, **sim_options):
    g = oracle.metric(kind, **metric_kwargs)
    cam = oracle.camera(*cam_args)
    H = cam_args[-1]
    return oracle.render_rows(g, cam, oracle.sim(*sim, **sim_options), bp, bn, row_begin=first, row_end=H, row_stride=stride,
                              threads=os.cpu_count() or 1, with_records=True)


def gpu_rows(system, sim, rows, **options):
    """(rgb (len(rows), W, 3), records (len(rows), W), summed stats) — one curvis_render_rows call per row."""
    frames, recs = [], []
    tot = {}
    for r in rows:
        f, rec = system.render_rows(*sim, r, r + 1, with_records=True, **options)
        frames.append(f)
        recs.append(rec)
        for k, v in system.last_stats.items():
            if k.startswith("n_") or k == "total_steps":
                tot[k] = tot.get(k, 0) + v
    return np.concatenate(frames, axis=0), np.concatenate(recs, axis=0), tot


def kicked_mask(ref_rec):
    """Rays with stiffness >= 1 on the ORACLE's record: some Euler step advanced phi by a radian or more."""
    with np.errstate(invalid="ignore"):
        return ~(ref_rec["stiffness"] < 1.0)


def report(name, cmp, kicked, bad):
    print(f"[parity] {name}: rays {cmp['rays']}, chaotic {cmp['chaotic']} ({100 * cmp['chaotic_fraction']:.2f} %), kicked {int(kicked.sum())}; "
          f"differing regular px {cmp['differing_pixels_regular']} / records {cmp.get('differing_records_regular')}, "
          f"chaotic px {cmp['differing_pixels_chaotic']} / records {cmp.get('differing_records_chaotic')}, "
          f"differing among stiffness < 1: {int((bad & ~kicked).sum())}")
