"""CURVIS_INTEGRATOR_RK4 (extension; the reference only has forward Euler, src/metrics.rs:283-297).
Its oracle is oracle_step_rk4 (same right-hand side, stated operation order).  CPU: the oracle's RK4
is 4th-order accurate and agrees with a fine Euler integration; GPU: RGB8 / side / steps / texels
identical to the RK4 oracle, final state within the trig tolerance."""
import math
import os

import numpy as np
import pytest


def _integrate(oracle, g, x, p, delta, n, rk4):
    for _ in range(n):
        x, p = oracle.step(g, x, p, delta, rk4=rk4)
    return x, p


def test_rk4_oracle_order_of_accuracy(oracle):
    g = oracle.metric("ellis")
    x0, p0 = oracle.new_photon(g, (0, 5, math.pi / 2, 0), (math.cos(2.8), 0.2, math.sin(2.8)))
    ref_x, ref_p = _integrate(oracle, g, x0, p0, 0.0025, 8000, True)          # T = 20
    errs = []
    for delta, n in ((0.5, 40), (0.25, 80), (0.125, 160)):
        x, p = _integrate(oracle, g, x0, p0, delta, n, True)
        errs.append(max(np.abs(x - ref_x)[1:].max(), np.abs(p - ref_p).max()))
    assert errs[0] / errs[1] > 10 and errs[1] / errs[2] > 10                     # ~16x per halving
    xe, pe = _integrate(oracle, g, x0, p0, 0.05, 400, False)
    assert errs[0] < 0.05 * np.abs(xe - ref_x)[1:].max()                         # RK4 at 10x the step beats Euler
    # conserved momenta stay bit-exact, null norm drift is tiny compared with Euler's
    x, p = _integrate(oracle, g, x0, p0, 0.5, 40, True)
    assert p[0] == p0[0] and p[3] == p0[3]
    assert abs(oracle.squared_norm_cov(g, p, x)) < 1e-2 * abs(oracle.squared_norm_cov(g, pe, xe))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_rk4_gpu_matches_rk4_oracle(gpu_ctx, oracle, kind):
    """BASELINE.json configs[0] names "256x144, 200 RK4 steps": with delta 0.5 a ray needs ~202 steps to go
    from l = 5 past the default escape radius 100 (at 200 almost every ray is NotEscaped), so the frame uses 220."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    W, H, sim = 256, 144, (220, 100.0, 0.5)
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    frame, rec = sysm.render_rows(*sim, 0, H, with_records=True, integrator=_abi.INTEGRATOR_RK4)
    ref, rrec, rst = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(*sim, integrator=_abi.INTEGRATOR_RK4), bp, bn,
                                        threads=os.cpu_count() or 1)
    assert (frame == ref).all()
    for f in ("side", "steps", "texel_x", "texel_y"):
        assert (rec[f] == rrec[f]).all(), f
    assert sysm.last_stats["total_steps"] == rst["total_steps"] and rst["n_not_escaped"] < 0.05 * W * H
    ok = (rrec["side"] != 0) & np.isfinite(rrec["l"])
    np.testing.assert_allclose(rec["l"][ok], rrec["l"][ok], rtol=1e-9)
    np.testing.assert_allclose(rec["p_l"][ok], rrec["p_l"][ok], rtol=1e-9, atol=1e-9)
    # the Euler default is untouched by the option, and differs from RK4 (coarser but different trajectories)
    euler = sysm.render_rows(40000, 100.0, 0.05, 0, H)
    eref, _, _ = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(40000, 100.0, 0.05), bp, bn, threads=os.cpu_count() or 1,
                                    with_records=False)
    assert (euler == eref).all() and (euler != frame).any()
    with pytest.raises(cv.CurvisError):
        sysm.render_image(*sim, integrator=_abi.INTEGRATOR_RK4, precision=_abi.PRECISION_F32)
    with pytest.raises(cv.CurvisError):
        sysm.render_image(*sim, integrator=5)
