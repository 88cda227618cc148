"""CURVIS_SAMPLING_BILINEAR (extension; the reference only has the nearest u8 lookup).  Its
oracle is the fp32 restatement in oracle/curvis_oracle.c (oracle_bilinear_tap):

  * fed the SAME continuous coordinates, the device tap is bit-identical (0 ULP <= the 1 ULP
    north_star asks) — float4 texels, 128-bit loads, three fmaf lerps;
  * through the whole pipeline the coordinates themselves come from acos/atan2, which differ by
    an ulp of fp64 between device and glibc, so unrounded colours agree to 2^-15 on the 0..255
    scale (1 ULP of the largest texel value ~ 1.5e-5) and the rounded RGB8 on >= 99.99 % of pixels.
Nearest stays the default; with nearest sampling the float output equals the texel bytes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(cv, scenes, ctx, W=192, H=108, bw=301, bh=157):
    bp, bn = scenes.noise_background(bw, bh, 41), scenes.noise_background(bw, bh, 42)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=ctx)
    return sysm, bp, bn, cam_args


def test_tap_is_bit_identical_on_equal_coordinates(gpu_ctx, oracle):
    import curvis_b200 as cv
    from curvis_b200 import scenes
    sysm, bp, bn, _ = _scene(cv, scenes, gpu_ctx)
    rng = np.random.default_rng(5)
    W, H = bp.shape[1], bp.shape[0]
    fx = np.concatenate([rng.uniform(0, W, 4000), [0.0, 0.25, 0.5, W - 0.5, W - 0.25, np.nextafter(W, 0), 17.5, 3.0]])
    fy = np.concatenate([rng.uniform(0, H, 4000), [0.0, 0.25, 0.5, H - 0.5, H - 0.1, H, 0.5, H / 2]])
    for side, bg in ((1, bp), (-1, bn)):
        got = sysm.debug_bilinear(side, fx, fy)
        ref = oracle.bilinear_tap(bg, fx, fy)
        assert got.tobytes() == ref.tobytes()
    # at texel centres the tap returns the texel itself; across the x seam it wraps, in y it clamps
    got = sysm.debug_bilinear(1, np.array([10.5, 0.0, 5.5, 5.5]), np.array([20.5, 7.5, 0.0, float(H)]))
    assert (got[0] == bp[20, 10].astype(np.float32)).all()
    assert (got[1] == 0.5 * (bp[7, W - 1].astype(np.float32) + bp[7, 0].astype(np.float32))).all()
    assert (got[2] == bp[0, 5].astype(np.float32)).all() and (got[3] == bp[H - 1, 5].astype(np.float32)).all()


def test_bilinear_frame_against_oracle(gpu_ctx, oracle):
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    sysm, bp, bn, cam_args = _scene(cv, scenes, gpu_ctx)
    W, H = cam_args[5], cam_args[6]
    sim = (40000, 100.0, 0.05)
    frame = sysm.render_image(*sim, sampling=_abi.SAMPLING_BILINEAR)
    taps = sysm.render_rows_rgba32f(*sim, 0, H, sampling=_abi.SAMPLING_BILINEAR)
    ref_f = np.zeros((H, W, 4), np.float32)
    ref, _, _ = oracle.render_rows(oracle.metric("ellis"), oracle.camera(*cam_args), oracle.sim(*sim, sampling=_abi.SAMPLING_BILINEAR), bp, bn,
                                   threads=os.cpu_count() or 1, rgba32f=ref_f)
    assert np.abs(taps - ref_f).max() <= 2.0 ** -15 * 255.0 / 128.0 + 1e-4        # ~1 ULP of the texel scale (+ slack for steep lerps)
    assert (np.abs(taps - ref_f) <= 2.0 ** -15).mean() >= 0.999
    assert (frame == ref).all(axis=2).mean() >= 0.9999
    q = np.clip(np.rint(taps[..., :3]), 0, 255).astype(np.uint8)
    assert (q == frame).all()                                                        # RGB8 = rounded taps
    # nearest mode: float output = texel bytes; RGB8 unchanged by the refactor
    near = sysm.render_image(*sim)
    near_f = sysm.render_rows_rgba32f(*sim, 0, H)
    assert (near_f[..., :3] == near.astype(np.float32)).all() and set(np.unique(near_f[..., 3])) <= {255.0}
    ref_n, _, _ = oracle.render_rows(oracle.metric("ellis"), oracle.camera(*cam_args), oracle.sim(*sim), bp, bn, threads=os.cpu_count() or 1)
    assert (near == ref_n).all()
    assert (frame != near).any()                                                      # the filter does something
