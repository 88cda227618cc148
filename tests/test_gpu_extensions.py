"""The extensions of SURVEY.md 8f row f4 on the GPU, each against its own oracle restatement (tests/test_oracle_extensions.py
checks those restatements on the CPU): the pole-adaptive Euler step, the chart-free ("pole-safe") angular state, the
world-frame lookup.  Integer results must be identical to the oracle's; the end state within the tolerance stated."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(kind, W, H, bg=(4096, 2048)):
    import curvis_b200 as cv
    from curvis_b200 import scenes
    bp, bn = scenes.decodable_background(*bg), scenes.decodable_background(*bg, negative=True)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    return metric, cam_args, bp, bn


def _integers_differ(rgb, rec, ref_rgb, ref_rec):
    return (rgb != ref_rgb).any(axis=-1) | (rec["steps"] != ref_rec["steps"]) | (rec["side"] != ref_rec["side"]) | \
           (rec["texel_x"] != ref_rec["texel_x"]) | (rec["texel_y"] != ref_rec["texel_y"])


@pytest.mark.parametrize("kind,tol", [("ellis", 0.03), ("ellis", 0.01), ("interstellar", 0.01)])
def test_adaptive_step_matches_its_oracle(gpu_ctx, oracle, kind, tol):
    import curvis_b200 as cv
    from curvis_b200 import _abi
    from oracle import classify
    W, H, sim = 256, 144, (40000, 100.0, 0.05)
    metric, cam_args, bp, bn = _scene(kind, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    opts = dict(integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=tol)
    rgb, rec = system.render_rows(*sim, 0, H, with_records=True, **opts)
    st = dict(system.last_stats)
    ref_rgb, ref_rec, ref_st = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(*sim, **opts), bp, bn,
                                                  threads=os.cpu_count() or 1)
    bad = _integers_differ(rgb, rec, ref_rgb, ref_rec)
    chaotic = classify.chaotic_mask(ref_rec)
    print(f"[adaptive {kind} tol {tol}] chaotic {chaotic.mean():.4f}, differing {int(bad.sum())} (regular {int((bad & ~chaotic).sum())}), "
          f"steps {st['total_steps']} vs {ref_st['total_steps']}")
    assert int((bad & ~chaotic).sum()) == 0
    assert int(bad.sum()) <= 2
    assert abs(st["total_steps"] - ref_st["total_steps"]) <= 40000 * int(bad.sum())
    assert (system.render_image(*sim, **opts) == rgb).all()
    # with a tolerance no step reaches, the kernel is the reference's Euler step bit for bit
    base = system.render_rows(*sim, 0, H, with_records=True)
    loose = system.render_rows(*sim, 0, H, with_records=True, integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=1e300)
    assert (base[0] == loose[0]).all() and base[1].tobytes() == loose[1].tobytes()


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_cartesian_coordinates_match_their_oracle(gpu_ctx, oracle, kind):
    import curvis_b200 as cv
    from curvis_b200 import _abi
    W, H, sim = 256, 144, (40000, 100.0, 0.05)
    metric, cam_args, bp, bn = _scene(kind, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    opts = dict(coordinates=_abi.COORDINATES_CARTESIAN)
    rgb, rec = system.render_rows(*sim, 0, H, with_records=True, **opts)
    st = dict(system.last_stats)
    ref_rgb, ref_rec, ref_st = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(*sim, **opts), bp, bn,
                                                  threads=os.cpu_count() or 1)
    bad = _integers_differ(rgb, rec, ref_rgb, ref_rec)
    print(f"[cartesian {kind}] differing {int(bad.sum())} of {W * H}; |p_l| max {np.abs(rec['p_l']).max():.4f}; steps {st['total_steps']}")
    assert int(bad.sum()) == 0
    assert st["total_steps"] == ref_st["total_steps"]
    if kind == "ellis":      # only IEEE + - * / sqrt in the loop: the radial state is the oracle's, bit for bit
        assert rec["l"].tobytes() == ref_rec["l"].tobytes() and rec["p_l"].tobytes() == ref_rec["p_l"].tobytes()
        assert rec["p_phi"].tobytes() == ref_rec["p_phi"].tobytes()
    np.testing.assert_allclose(rec["theta"], ref_rec["theta"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(rec["p_theta"], ref_rec["p_theta"], rtol=1e-9, atol=1e-12)
    assert (np.abs(rec["p_l"][rec["side"] != 0]) <= 1.05).all()            # no kicked rays in this chart
    assert (system.render_image(*sim, **opts) == rgb).all()


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_regrouped_cartesian_kernel_against_the_cartesian_oracle(gpu_ctx, oracle, kind):
    """CURVIS_COORDINATES_CARTESIAN + CURVIS_PRECISION_F64_FAST: the chart-free scheme regrouped for the fp64 pipe (17 fp64
    instructions per Ellis step).  No guard band here, so the bar is the raw regrouped kernel's: integers identical to the
    chart-free oracle on >= 99.99 % of the rays (a ~1e-13 state difference flips a truncation now and then), state within
    1e-9, no kicked ray, and the same frame from every entry point."""
    import curvis_b200 as cv
    from curvis_b200 import _abi
    W, H, sim = 320, 180, (40000, 100.0, 0.05)
    metric, cam_args, bp, bn = _scene(kind, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    opts = dict(coordinates=_abi.COORDINATES_CARTESIAN, precision=_abi.PRECISION_F64_FAST)
    rgb, rec = system.render_rows(*sim, 0, H, with_records=True, **opts)
    st = dict(system.last_stats)
    ref_rgb, ref_rec, ref_st = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args),
                                                  oracle.sim(*sim, coordinates=_abi.COORDINATES_CARTESIAN), bp, bn, threads=os.cpu_count() or 1)
    bad = _integers_differ(rgb, rec, ref_rgb, ref_rec)
    print(f"[cartesian fast {kind}] differing {int(bad.sum())} of {W * H}; steps {st['total_steps']} vs {ref_st['total_steps']}")
    assert int(bad.sum()) <= max(1, int(1e-4 * W * H))
    assert abs(st["total_steps"] - ref_st["total_steps"]) <= int(bad.sum()) * 2
    ok = ~bad & (ref_rec["side"] != 0)
    np.testing.assert_allclose(rec["l"][ok], ref_rec["l"][ok], rtol=1e-9)
    np.testing.assert_allclose(rec["p_l"][ok], ref_rec["p_l"][ok], rtol=1e-9)
    np.testing.assert_allclose(rec["theta"][ok], ref_rec["theta"][ok], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(rec["p_phi"][ok], ref_rec["p_phi"][ok], rtol=1e-12, atol=1e-15)
    assert (np.abs(rec["p_l"][rec["side"] != 0]) <= 1.05).all()
    assert (system.render_image(*sim, **opts) == rgb).all()
    # exotic: zero iterations, NotEscaped budget, negative step
    for s2 in ((0, 100.0, 0.05), (300, 100.0, 0.05), (3000, 60.0, -0.05)):
        a = system.render_rows(*s2, 0, H, with_records=True, **opts)
        b = system.render_rows(*s2, 0, H, with_records=True, coordinates=_abi.COORDINATES_CARTESIAN)
        assert ((a[0] != b[0]).any(axis=2) | (a[1]["steps"] != b[1]["steps"]) | (a[1]["side"] != b[1]["side"])).sum() <= max(1, int(1e-4 * W * H)), s2


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
@pytest.mark.parametrize("frame_name", ["FRAME_WORLD", "FRAME_WORLD_QUIRK"])
def test_world_frame_matches_its_oracle(gpu_ctx, oracle, kind, frame_name):
    import curvis_b200 as cv
    from curvis_b200 import _abi
    from oracle import classify
    frame = getattr(_abi, frame_name)
    W, H, sim = 256, 144, (40000, 100.0, 0.05)
    metric, cam_args, bp, bn = _scene(kind, W, H)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    for precision in (_abi.PRECISION_F64, _abi.PRECISION_F64_FAST):
        rgb, rec = system.render_rows(*sim, 0, H, with_records=True, frame=frame, precision=precision)
        ref_rgb, ref_rec, _ = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(*sim, frame=frame), bp, bn,
                                                 threads=os.cpu_count() or 1)
        bad = _integers_differ(rgb, rec, ref_rgb, ref_rec)
        chaotic = classify.chaotic_mask(ref_rec)
        print(f"[{frame_name} {kind} precision {precision}] differing {int(bad.sum())} (regular {int((bad & ~chaotic).sum())}) of {W * H}")
        assert int((bad & ~chaotic).sum()) == 0
        # the rotation reads sin / cos of the END POSITION's theta and phi, which wander to 1e8 on kicked rays: CUDA's and
        # glibc's reductions of such arguments agree to an ulp of the angle, i.e. ~1e-8 rad — enough to flip a texel now and then
        assert int(bad.sum()) <= max(2, int(2e-3 * chaotic.sum())), (int(bad.sum()), int(chaotic.sum()))
        if precision == _abi.PRECISION_F64:      # the frame mode changes the lookup only: same photons
            base = system.render_rows(*sim, 0, H, with_records=True)
            for f in ("l", "theta", "phi", "p_l", "p_theta", "p_phi", "steps", "side"):
                assert rec[f].tobytes() == base[1][f].tobytes(), f


def test_world_frame_central_column_carries_the_table_angle(gpu_ctx, oracle):
    """The cross-renderer check of tests/test_oracle_extensions.py, on the GPU: for the pixels whose ray stays on the equator
    the texel looked up in CURVIS_FRAME_WORLD_QUIRK is the texel of the direction compute_escape_angle works with."""
    import math
    import curvis_b200 as cv
    from curvis_b200 import _abi
    W, H, sim = 256, 144, (40000, 100.0, 0.05)
    metric, cam_args, bp, bn = _scene("ellis", W, H, bg=(8192, 4096))
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    rgb, rec = system.render_rows(*sim, 0, H, with_records=True, frame=_abi.FRAME_WORLD_QUIRK)
    g, cam = oracle.metric("ellis"), oracle.camera(*cam_args)
    col = W // 2
    checked = 0
    for row in range(H):
        d = oracle.outward_vector(cam, col, row)
        alpha = math.atan2(d[2], d[0])
        side, angle, steps = oracle.compute_escape_angle(g, 5.0, alpha, 0.05, 40000, 100.0)
        assert (side, steps) == (int(rec["side"][row, col]), int(rec["steps"][row, col]))
        if side == 0:
            continue
        # compute_escape_angle's `angle` is the 3-D angle between the world direction w and the x axis (systems.rs:246-251:
        # acos of w.x after normalisation, mirrored for w.y < 0); the texel the GPU looked up gives w's azimuth and polar
        # angle to half a texel: cos(angle) = cos(azimuth) sin(polar)
        az = (0.5 - (int(rec["texel_x"][row, col]) + 0.5) / 8192.0) * 2 * math.pi
        pol = (int(rec["texel_y"][row, col]) + 0.5) / 4096.0 * math.pi
        got = math.acos(max(-1.0, min(1.0, math.cos(az) * math.sin(pol))))
        if math.sin(az) < 0.0:
            got = 2 * math.pi - got
        assert abs(got - angle) <= 3 * (2 * math.pi / 8192), (row, angle, got)
        checked += 1
    assert checked >= 100
