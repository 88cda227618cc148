"""Regular / chaotic classification of rays from the ORACLE's records.  TEST INFRASTRUCTURE, like
the rest of oracle/: used by tests/, bench.py's parity_check and __graft_entry__.smoke().

Forward Euler at the reference's default step in (theta, phi) coordinates (src/metrics.rs:257-262 carries 1/sin^2 theta
and cos theta / sin^3 theta) turns every ray that passes close to a coordinate pole into a "kicked" ray: its end state is an
artefact of the pole, it amplifies last-bit differences of sin/cos by many orders of magnitude, and it typically leaves
with |p_l| well above the value (1) a null geodesic has far from the throat.  The survey's rule:

    chaotic  <=>  |p_l|_final > 1.05   or   min |sin theta| along the trajectory < 1e-3

(`min_abs_sin_theta` of curvis_ray_record; the oracle tracks it when records are requested).  Parity reports split every
count by this classification: differences on regular rays are failures, chaotic rays are counted and reported.
"""
from __future__ import annotations

import numpy as np

P_L_LIMIT = 1.05
MIN_SIN_LIMIT = 1e-3


def chaotic_mask(rec) -> np.ndarray:
    """Boolean array, True where the oracle's record marks the ray chaotic.  NaN states count as chaotic."""
    p_l = np.abs(rec["p_l"])
    with np.errstate(invalid="ignore"):
        return ~(p_l <= P_L_LIMIT) | ~(rec["min_abs_sin_theta"] >= MIN_SIN_LIMIT)


def compare(gpu_rgb, gpu_rec, ref_rgb, ref_rec) -> dict:
    """Integer results of a GPU tile against the oracle's, split by the classification.  Arrays are (rows, W[, 3])."""
    chaotic = chaotic_mask(ref_rec)
    pix = (gpu_rgb != ref_rgb).any(axis=-1)
    out = {"rays": int(chaotic.size), "chaotic": int(chaotic.sum()), "chaotic_fraction": float(chaotic.mean()) if chaotic.size else 0.0,
           "differing_pixels_regular": int((pix & ~chaotic).sum()), "differing_pixels_chaotic": int((pix & chaotic).sum())}
    if gpu_rec is not None:
        integers = (gpu_rec["steps"] != ref_rec["steps"]) | (gpu_rec["side"] != ref_rec["side"]) | \
                   (gpu_rec["texel_x"] != ref_rec["texel_x"]) | (gpu_rec["texel_y"] != ref_rec["texel_y"])
        out["differing_records_regular"] = int((integers & ~chaotic).sum())
        out["differing_records_chaotic"] = int((integers & chaotic).sum())
    return out
