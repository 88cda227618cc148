"""Small frames through every per-ray kernel (for compute-sanitizer memcheck / racecheck on the GPU box):
the three precisions x three metrics, records on, a registered host frame (zero-copy stores), the fused peer-store launch, the table-based renderer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import curvis_b200 as cv
from curvis_b200 import scenes, _abi

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(256, 128), scenes.decodable_background(256, 128, True)
W, H = 96, 54
cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
out = np.empty((H, W, 3), dtype=np.uint8)
ctx.register_host_buffer(out)
for metric, sim in ((cv.EllisMetric(1.0), (300, 12.0, 0.1)), (cv.InterstellarMetric(0.1, 1e-4, 1.0), (300, 12.0, 0.1)),
                    (cv.FlatSphericalMetric(), (400, 12.0, 0.1))):
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
    for prec in (_abi.PRECISION_F64, _abi.PRECISION_F64_FAST, _abi.PRECISION_F32):
        sysm.render_image(*sim, precision=prec, out=out)
        frame, rec = sysm.render_rows(*sim, 3, 41, with_records=True, precision=prec)
        assert (frame == out[3:41]).all()
        print(type(metric).__name__, prec, sysm.last_stats["total_steps"], flush=True)
# fused render + gather into two peer buffers (interleaved rows), and the table-based renderer into the registered frame
import torch
bufs = [cv.PeerBuffer.create(ctx, 2 * W * H * 3) for _ in range(2)]
cams = [cam, cv.Camera((0.0, 5.5, 1.4, 0.2), scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)]
sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
for prec in (_abi.PRECISION_F64, _abi.PRECISION_F64_FAST, _abi.PRECISION_F32):
    for g in range(3):
        sysm.render_frames_peers(cams, 300, 12.0, 0.1, g, H, [b.ptr for b in bufs], want_stats=True, row_stride=3, precision=prec)
torch.cuda.synchronize()
assert (bufs[0].as_tensor() == bufs[1].as_tensor()).all()
for b in bufs:
    b.close()
for prec in (_abi.PRECISION_F64, _abi.PRECISION_F64_FAST):
    sysm.render_image_efficient(2000, 30.0, 0.05, 30, 30, 1e-4, 1e-4, out=out, precision=prec)
# round 2: the guard band's second launch in both modes, the extensions, two streams on one context
for guard in (2, 1):
    ctx.set_option("guard", guard)
    for metric in (cv.EllisMetric(1.0), cv.InterstellarMetric(0.1, 1e-4, 1.0)):
        sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
        sysm.render_image(3000, 60.0, 0.05, precision=_abi.PRECISION_F64_FAST, out=out)
        print("guard", guard, type(metric).__name__, sysm.last_stats["n_reintegrated"], sysm.last_stats["n_kicked"], flush=True)
sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)
for opts in (dict(integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=0.01), dict(coordinates=_abi.COORDINATES_CARTESIAN),
             dict(frame=_abi.FRAME_WORLD), dict(frame=_abi.FRAME_WORLD_QUIRK, precision=_abi.PRECISION_F64_FAST), dict(integrator=_abi.INTEGRATOR_RK4)):
    sysm.render_image(3000, 60.0, 0.05, out=out, **opts)
    sysm.render_rows(3000, 60.0, 0.05, 5, 20, with_records=True, **opts)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
tiles = [torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0") for _ in range(4)]
for i in range(4):
    sysm.render_rows_device(3000, 60.0, 0.05, 0, H, tiles[i].data_ptr(), streams[i & 1].cuda_stream, precision=_abi.PRECISION_F64_FAST)
torch.cuda.synchronize()
assert all((t == tiles[0]).all() for t in tiles)
# round 2, later: the longest-first pre-pass + list (launches of >= 2^15 rays), the cached Interstellar table, the strict kernel's atan / ln tables
W2, H2 = 256, 144
cam2 = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W2, H2)
for lf in (1, 2, 0):
    ctx.set_option("longest_first", lf)
    for metric in (cv.EllisMetric(1.0), cv.InterstellarMetric(0.1, 1e-4, 1.0)):
        sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam2, context=ctx)
        a = sysm.render_image(400, 12.0, 0.1, precision=_abi.PRECISION_F64_FAST)
        b = sysm.render_image(400, 12.0, 0.1, precision=_abi.PRECISION_F64)
        assert (a == b).all()
        print("longest_first", lf, type(metric).__name__, sysm.last_stats["total_steps"], flush=True)
ctx.set_option("longest_first", 2)
ctx.unregister_host_buffer(out)
print("ok")
