"""CURVIS_PRECISION_F32 (opt-in fast mode, an extension with no reference counterpart): fp32
right-hand side + Kahan-compensated state.  It is not bit-comparable with the reference, so the
bar is a stated tolerance against the fp64 parity kernel and the oracle:

  * escape side identical on >= 99.99 % of rays, step count within +-1 on >= 99.99 % (equal on >= 99.9 %);
  * end direction within 1e-5 rad (BASELINE.json north_star) on >= 99 % of rays, median < 1e-7 rad;
  * texel index equal or 8-neighbour-adjacent (wrap in x) on >= 99.9 % of rays.
The parity mode stays the default; nothing here relaxes its tests."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TWO_OVER_PI = np.longdouble(2) / (4 * np.arctan(np.longdouble(1)))      # in x87 long double


def _dir_ellis(rec, rho=1.0):
    s = np.sin(rec["theta"])
    r = np.sqrt(rho * rho + rec["l"] ** 2)
    return np.stack([rec["p_l"], rec["p_theta"] / r, rec["p_phi"] / (r * s * s)], -1)   # metrics.rs:339-349


def _angle(a, b):
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=-1), (a * b).sum(-1))


@pytest.mark.parametrize("sim,W,H", [((40000, 100.0, 0.05), 256, 144), ((1000, 25.0, 0.05), 480, 270)])
def test_fast_mode_tolerance_vs_parity_kernel_and_oracle(gpu_ctx, oracle, sim, W, H):
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    f64, r64 = sysm.render_rows(*sim, 0, H, with_records=True)
    f32, r32 = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F32)
    _, ref, _ = oracle.render_rows(oracle.metric("ellis"), oracle.camera(*cam_args), oracle.sim(*sim), bp, bn, threads=os.cpu_count() or 1)
    for name, base in (("parity kernel", r64), ("oracle", ref)):
        assert (base["side"] == r32["side"]).mean() >= 0.9999, name
        dsteps = np.abs(base["steps"].astype(np.int64) - r32["steps"].astype(np.int64))
        assert (dsteps <= 1).mean() >= 0.9999 and (dsteps == 0).mean() >= 0.999, name
        esc = (base["side"] != 0) & (r32["side"] != 0)
        ang = _angle(_dir_ellis(base)[esc], _dir_ellis(r32)[esc])
        assert (ang <= 1e-5).mean() >= 0.99, (name, float((ang <= 1e-5).mean()))
        assert np.median(ang) < 1e-7, name
        dx = np.abs(base["texel_x"].astype(np.int64) - r32["texel_x"].astype(np.int64))
        dx = np.minimum(dx, 4096 - dx)
        dy = np.abs(base["texel_y"].astype(np.int64) - r32["texel_y"].astype(np.int64))
        assert ((dx <= 1) & (dy <= 1))[esc].mean() >= 0.999, name
    assert (f64 == f32).all(axis=2).mean() >= 0.99          # most pixels are even byte-identical
    assert sysm.last_stats["n_rays"] == W * H


def test_fast_mode_other_metrics_and_edges(gpu_ctx):
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.decodable_background(2048, 1024), scenes.decodable_background(2048, 1024, True)
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 192, 108)
    for metric in (cv.InterstellarMetric(0.1, 1e-4, 1.0), cv.FlatSphericalMetric()):
        sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
        _, r64 = sysm.render_rows(4000, 100.0, 0.05, 0, 108, with_records=True)
        _, r32 = sysm.render_rows(4000, 100.0, 0.05, 0, 108, with_records=True, precision=_abi.PRECISION_F32)
        ok = np.isfinite(r64["l"])
        assert (r64["side"] == r32["side"])[ok].mean() >= 0.999
        assert (np.abs(r64["steps"].astype(np.int64) - r32["steps"].astype(np.int64)) <= 1)[ok].mean() >= 0.999
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
    frame = sysm.render_image(0, 100.0, 0.05, precision=_abi.PRECISION_F32)      # zero iterations
    assert (frame == 0).all() and sysm.last_stats["total_steps"] == 0
    frame = sysm.render_image(50, 100.0, 0.05, precision=_abi.PRECISION_F32)     # nobody escapes
    assert (frame == 0).all() and sysm.last_stats["n_not_escaped"] == 192 * 108
    with pytest.raises(cv.CurvisError):
        sysm.render_image(10, 100.0, 0.05, precision=7)


def test_fp32_shape_table(gpu_ctx):
    """The fp32 edition of the Interstellar shape-function table (csrc/shape_table.h: 16 intervals per binade,
    degree 3): F(x) = x atan x - ln(1+x^2)/2 and G(x) = atan x within 4 fp32 ulp inside [2^-10, 2^16), and the
    library composition outside it (absolute bar for tiny x, as in the fp64 test)."""
    rng = np.random.default_rng(11)
    x = np.exp(rng.uniform(np.log(2.0 ** -10), np.log(2.0 ** 16), 500_000)).astype(np.float32).astype(np.float64)
    xl = x.astype(np.longdouble)
    want_f, want_g = xl * np.arctan(xl) - np.log1p(xl * xl) / 2, TWO_OVER_PI * np.arctan(xl)

    def ulps32(got, want):
        return np.abs((got.astype(np.longdouble) - want) / np.spacing(np.abs(want.astype(np.float32))).astype(np.longdouble)).astype(np.float64)

    assert ulps32(gpu_ctx.debug_eval(15, x), want_f).max() <= 4.0
    assert ulps32(gpu_ctx.debug_eval(16, x), want_g).max() <= 4.0
    tiny = np.exp(rng.uniform(np.log(1e-5), np.log(2.0 ** -10), 5_000)).astype(np.float32).astype(np.float64)
    huge = np.exp(rng.uniform(np.log(2.0 ** 16), np.log(1e9), 5_000)).astype(np.float32).astype(np.float64)
    tl, hl = tiny.astype(np.longdouble), huge.astype(np.longdouble)
    assert np.abs(gpu_ctx.debug_eval(15, tiny).astype(np.longdouble) - (tl * np.arctan(tl) - np.log1p(tl * tl) / 2)).max() <= 2e-7
    assert ulps32(gpu_ctx.debug_eval(15, huge), hl * np.arctan(hl) - np.log1p(hl * hl) / 2).max() <= 8.0
    assert ulps32(gpu_ctx.debug_eval(16, np.concatenate([tiny, huge])), TWO_OVER_PI * np.arctan(np.concatenate([tl, hl]))).max() <= 4.0
