"""Round-2 final session's new launch paths, small, for compute-sanitizer (no torch: starts quickly): block tiles, the longest-first
list claimed by no / all / the favoured warp slots in both fp64 kernels, the re-integration list walked from either end with guard = 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi, distributed

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(256, 128), scenes.decodable_background(256, 128, True)
W, H, sim = 256, 144, (300, 12.0, 0.1)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
for metric in (cv.EllisMetric(1.0), cv.InterstellarMetric(0.1, 1e-4, 1.0)):
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    ref = sysm.render_image(*sim, precision=_abi.PRECISION_F64).copy()
    for guard, lf, slots in ((2, 1, 8), (2, 0, 8), (1, 1, 0), (1, 1, 64)):
        ctx.set_option("guard", guard); ctx.set_option("longest_first", lf); ctx.set_option("favoured_slots", slots)
        a = sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST)
        b = sysm.render_image(*sim, precision=_abi.PRECISION_F64)
        assert (b == ref).all() and (guard != 2 or (a == ref).all())
        print(type(metric).__name__, guard, lf, slots, sysm.last_stats["total_steps"], int((a != ref).sum()), flush=True)
    ctx.set_option("guard", 1); ctx.set_option("longest_first", 2); ctx.set_option("favoured_slots", 8)
    buf = cv.PeerBuffer.create(ctx, W * H * 3)
    for prec in (_abi.PRECISION_F64, _abi.PRECISION_F64_FAST):
        total = 0
        for g in range(3):
            b0, b1, stride, bw = distributed.interleaved_blocks(H, W, g, 3, 32)
            total += sysm.render_frames_peers([cam], *sim, b0, b1, [buf.ptr], want_stats=True, row_stride=stride, block_width=bw, precision=prec)["n_rays"]
        assert total == W * H
    buf.close()
print("ok")
