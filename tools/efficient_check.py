"""Dev script (GPU box): timing of the table-based renderer at several sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
ctx = cv.Context([0])
for kind in ("ellis", "interstellar"):
    for (W, H) in ((960, 540), (3840, 2160)):
        metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
        cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
        sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
        for _ in range(3):
            t = time.time(); sysm.render_image_efficient(40000, 100.0, 0.05, 100, 100, 1e-5, 1e-5); wall = (time.time() - t) * 1e3
        buf = np.empty((H, W, 3), dtype=np.uint8)
        ctx.register_host_buffer(buf)
        for _ in range(3):
            t = time.time(); sysm.render_image_efficient(40000, 100.0, 0.05, 100, 100, 1e-5, 1e-5, out=buf); wall_reg = (time.time() - t) * 1e3
        ref = buf.copy(); info_f64 = dict(sysm.last_efficient_info)
        for _ in range(3):
            t = time.time(); sysm.render_image_efficient(40000, 100.0, 0.05, 100, 100, 1e-5, 1e-5, out=buf, precision=_abi.PRECISION_F64_FAST); wall_fast = (time.time() - t) * 1e3
        fast = dict(wall_ms=wall_fast, table_ms=sysm.last_efficient_info["table_ms"], table_points=sysm.last_efficient_info["table_points"],
                    differing_pixels_vs_f64=int((buf != ref).any(axis=2).sum()))
        sysm.last_efficient_info = info_f64
        ctx.unregister_host_buffer(buf)
        print(json.dumps(dict(kind=kind, W=W, H=H, wall_ms=wall, wall_ms_registered_frame=wall_reg, f64_fast_table=fast, **sysm.last_efficient_info,
                              black=sysm.last_stats["n_not_escaped"])), flush=True)
