"""Dev script (GPU box): wall time of curvis_render_image (kernel + read-back into a pageable host frame), 4K Ellis
defaults, pageable vs registered caller frame vs zero-copy stores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi
ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, True)
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 3840, 2160)
sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
sim = (40000, 100.0, 0.05)
out = np.empty((2160, 3840, 3), dtype=np.uint8)
def run(label):
    sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST, out=out)
    t0 = time.perf_counter(); k = []
    for _ in range(10):
        sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST, out=out); k.append(sysm.last_stats["kernel_ms"])
    dt = (time.perf_counter() - t0) / 10 * 1e3
    print(json.dumps(dict(mode=label, wall_ms_per_frame=dt, kernel_ms=float(np.mean(k)), exposed_ms=dt - float(np.mean(k)))), flush=True)

run("pageable caller frame (staging copy)")
t0 = time.perf_counter(); ctx.register_host_buffer(out); print("register: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
run("registered caller frame (one DMA)")
ctx.set_option("zero_copy", 1)
run("registered caller frame, kernel stores into it (zero_copy)")
ctx.set_option("zero_copy", 0)
run("registered caller frame (one DMA)")
