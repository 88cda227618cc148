"""Generates tests/golden/*.npz from the CPU oracle (oracle/curvis_oracle.c) in THIS container.
The reference is Rust and cannot be built here (no cargo), so these vectors pin the ORACLE
(regression + cross-machine reproducibility), not the reference; see oracle header
"PARITY UNPINNED".  Run:  python tools/make_golden.py
"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from curvis_b200 import scenes
from oracle import oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (metric kind, metric kwargs, W, H, (max_iter, R, delta), camera position, forward, up, bg size)
    "ellis_c1a_64x36": ("ellis", {}, 64, 36, (200, 10.0, 0.1), scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, (1024, 512)),
    "ellis_defaults_48x27": ("ellis", {}, 48, 27, (40000, 100.0, 0.05), scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, (1024, 512)),
    "interstellar_defaults_48x27": ("interstellar", {}, 48, 27, (40000, 100.0, 0.05), scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, (1024, 512)),
    "flat_40x30": ("flat", {}, 40, 30, (500, 20.0, 0.1), (0.0, 5.0, 1.2, 0.3), (1.0, 0.2, 0.1), (0.0, 0.1, 1.0), (512, 256)),
    "ellis_tilted_33x17": ("ellis", {"rho": 2.0}, 33, 17, (3000, 60.0, 0.05), (0.0, -7.0, 1.1, 2.0), (1.0, 0.3, -0.2), (0.1, 0.0, 1.0), (777, 333)),
}


def scene(name):
    kind, mk, W, H, sim, pos, fwd, up, (bw, bh) = CASES[name]
    g = O.metric(kind, **mk)
    cam = O.camera(pos, fwd, up, scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, W, H)
    bp = scenes.noise_background(bw, bh, seed=20251017)
    bn = scenes.noise_background(bw, bh, seed=20251018)
    return g, cam, O.sim(*sim), bp, bn


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        g, cam, s, bp, bn = scene(name)
        rgb, rec, st = O.render_rows(g, cam, s, bp, bn)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), rgb=rgb, rec=rec,
                            total_steps=np.uint64(st["total_steps"]),
                            counts=np.array([st["n_positive"], st["n_negative"], st["n_not_escaped"], st["n_clamped"]], dtype=np.uint64))
        print(name, rgb.shape, st["total_steps"], os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    main()
