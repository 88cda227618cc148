"""Parity of curvis_render_image_efficient (the table-based renderer the curvis binary runs,
src/systems.rs:333-527) against the oracle's restatement.  The equatorial integrations are
bit-identical (theta stays exactly pi/2, so no transcendental enters the Euler loop) and the
table is finalised on the host with the platform libm, so the TABLE must equal the oracle's
exactly; the per-pixel pass uses device acos/sin/cos, so pixels may differ only where a 1-ulp
change crosses a texel boundary: the bar is >= 99.99 % identical RGB."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DEFAULT_SAMPLING = (100, 100, 1e-5, 1e-5)      # main.rs:46-47 passes sampling_initial_nums twice


@pytest.mark.parametrize("kind,sim,W,H", [("ellis", (40000, 100.0, 0.05), 256, 144), ("interstellar", (40000, 100.0, 0.05), 192, 108),
                                          ("ellis", (1000, 25.0, 0.05), 160, 90)])
def test_efficient_matches_oracle(gpu_ctx, oracle, kind, sim, W, H):
    import curvis_b200 as cv
    from curvis_b200 import scenes
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    frame, dbg = sysm.render_image_efficient(*sim, *DEFAULT_SAMPLING, debug=True)
    info = sysm.last_efficient_info
    ref, rinfo, alpha, angle, space = oracle.render_image_efficient(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(*sim), bp, bn,
                                                                    *DEFAULT_SAMPLING, debug=True)
    assert (info["table_points"], info["table_evaluations"], info["table_steps"]) == \
        (rinfo["table_points"], rinfo["table_evaluations"], rinfo["table_steps"])
    assert sysm.last_stats["total_steps"] == rinfo["table_steps"]
    same = (frame == ref).all(axis=2)
    assert same.mean() >= 0.9999, f"{(~same).sum()} pixels differ"
    np.testing.assert_allclose(dbg[..., 0], alpha, rtol=0, atol=1e-14)
    ok = np.isfinite(angle)
    np.testing.assert_allclose(dbg[..., 1][ok], angle[ok], rtol=1e-12, atol=1e-12)
    assert (dbg[..., 2] == space).all() or (np.isnan(space) == np.isnan(dbg[..., 2])).all()
    st = sysm.last_stats
    assert st["n_positive"] + st["n_negative"] + st["n_not_escaped"] == W * H
    assert st["n_positive"] == int((space == 1.0).sum()) and st["n_negative"] == int((space == -1.0).sum())


def test_efficient_tilted_camera_and_errors(gpu_ctx, oracle):
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.noise_background(512, 256, 21), scenes.noise_background(512, 256, 22)
    cam_args = ((0.0, 6.0, 1.2, 0.7), (-1.0, 0.2, 0.1), (0.0, 0.1, 1.0), 20.0, 43.0, 120, 80)
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.5), cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    frame = sysm.render_image_efficient(40000, 100.0, 0.05, 50, 50, 1e-4, 1e-4)
    ref, _ = oracle.render_image_efficient(oracle.metric("ellis", rho=1.5), oracle.camera(*cam_args), oracle.sim(40000, 100.0, 0.05), bp, bn,
                                           50, 50, 1e-4, 1e-4)
    assert (frame == ref).all(axis=2).mean() >= 0.9995
    # the table integrated by the regrouped fp64 kernel (precision = F64_FAST): same sampler decisions, same frame
    info = dict(sysm.last_efficient_info)
    fast = sysm.render_image_efficient(40000, 100.0, 0.05, 50, 50, 1e-4, 1e-4, precision=_abi.PRECISION_F64_FAST)
    assert sysm.last_efficient_info["table_points"] == info["table_points"]
    assert sysm.last_efficient_info["table_evaluations"] == info["table_evaluations"]
    assert (fast == frame).all(axis=2).mean() >= 0.9999
    with pytest.raises(cv.CurvisError) as e32:
        sysm.render_image_efficient(40000, 100.0, 0.05, 50, 50, 1e-4, 1e-4, precision=_abi.PRECISION_F32)
    assert e32.value.code == _abi.ERR_UNSUPPORTED
    # into a registered caller frame (one DMA, no staging copy): the same bytes
    buf = np.zeros((80, 120, 3), dtype=np.uint8)
    gpu_ctx.register_host_buffer(buf)
    try:
        sysm.render_image_efficient(40000, 100.0, 0.05, 50, 50, 1e-4, 1e-4, out=buf)
        assert (buf == frame).all()
    finally:
        gpu_ctx.unregister_host_buffer(buf)
    with pytest.raises(cv.CurvisError) as e:                    # camera outside the escape radius (systems.rs:122-124)
        sysm.render_image_efficient(100, 2.0, 0.05, 50, 50, 1e-4, 1e-4)
    assert e.value.code == _abi.ERR_CAMERA_OUTSIDE_RADIUS
    with pytest.raises(cv.CurvisError):
        sysm.render_image_efficient(100, 100.0, 0.05, 2, 50, 1e-4, 1e-4)
