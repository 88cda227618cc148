"""Every BASELINE.json config at its FULL size, on the GPU through the C ABI, against the CPU oracle on >= 64 rows spread
over the frame (all host cores), for BOTH fp64 modes:

  * CURVIS_PRECISION_F64 — RGB8, escape side, step count and texel index identical on every ray;
  * CURVIS_PRECISION_F64_FAST (what bench.py times) — identical on every ray the oracle classifies as regular
    (oracle/classify.py) AND on every ray with stiffness < 1 (the rays its guard band covers by
    construction); chaotic / kicked rays are counted and printed, and their differing fraction is bounded.

C1a / C1b (256x144) live in test_gpu_parity.py / test_gpu_fast64.py; here: C2 (Ellis 1080p 1000 / 0.05 / 25), C3
(Interstellar 4K 2000 / 0.05 / 45), the Interstellar 4K frame at the default settings, C4's frame (Ellis 7680x4320,
defaults) and the bench frame (Ellis 4K defaults).  C5's camera path is covered in test_gpu_driver.py.
"""
import numpy as np
import pytest

from parity_util import gpu_rows, kicked_mask, oracle_rows, report, strided_rows

pytestmark = pytest.mark.gpu

CONFIGS = {
    "C2_ellis_1080p_1000_0.05_25": ("ellis", {}, 1920, 1080, (1000, 25.0, 0.05)),
    "C3_interstellar_4k_2000_0.05_45": ("interstellar", {}, 3840, 2160, (2000, 45.0, 0.05)),
    "interstellar_4k_defaults": ("interstellar", {}, 3840, 2160, (40000, 100.0, 0.05)),
    "bench_ellis_4k_defaults": ("ellis", {}, 3840, 2160, (40000, 100.0, 0.05)),
    "C4_ellis_8k_defaults": ("ellis", {}, 7680, 4320, (40000, 100.0, 0.05)),
}
N_ROWS = 64


@pytest.fixture(scope="module")
def backgrounds():
    from curvis_b200 import scenes
    return scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_config_against_oracle(gpu_ctx, oracle, backgrounds, name):
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    from oracle import classify
    kind, mk, W, H, sim = CONFIGS[name]
    bp, bn = backgrounds
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, scenes.DEFAULT_FOCAL_LENGTH,
                scenes.DEFAULT_DIAGONAL, W, H)
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    first, stride, rows = strided_rows(H, N_ROWS)
    assert len(rows) >= 64
    ref_rgb, ref_rec, ref_st = oracle_rows(oracle, kind, mk, cam_args, sim, bp, bn, first, stride)
    kicked = kicked_mask(ref_rec)

    # ---- CURVIS_PRECISION_F64: identity, every ray
    rgb, rec, st = gpu_rows(system, sim, rows)
    cmp = classify.compare(rgb, rec, ref_rgb, ref_rec)
    bad = (rgb != ref_rgb).any(axis=-1) | (rec["steps"] != ref_rec["steps"]) | (rec["side"] != ref_rec["side"]) | \
          (rec["texel_x"] != ref_rec["texel_x"]) | (rec["texel_y"] != ref_rec["texel_y"])
    report(name + " F64", cmp, kicked, bad)
    assert int(bad.sum()) == 0, f"{name} F64: {int(bad.sum())} rays differ from the oracle"
    assert st["total_steps"] == ref_st["total_steps"]
    for k in ("n_positive", "n_negative", "n_not_escaped", "n_clamped"):
        assert st[k] == ref_st[k], k
    # the trajectory diagnostics the classifier reads (GPU sin/cos are not glibc's: tolerance)
    fin = np.isfinite(ref_rec["stiffness"]) & (ref_rec["stiffness"] < 1e6)
    np.testing.assert_allclose(rec["stiffness"][fin], ref_rec["stiffness"][fin], rtol=1e-6)
    np.testing.assert_allclose(rec["min_abs_sin_theta"][fin], ref_rec["min_abs_sin_theta"][fin], rtol=1e-6, atol=1e-300)

    # ---- CURVIS_PRECISION_F64_FAST
    rgb, rec, st = gpu_rows(system, sim, rows, precision=_abi.PRECISION_F64_FAST)
    cmp = classify.compare(rgb, rec, ref_rgb, ref_rec)
    bad = (rgb != ref_rgb).any(axis=-1) | (rec["steps"] != ref_rec["steps"]) | (rec["side"] != ref_rec["side"]) | \
          (rec["texel_x"] != ref_rec["texel_x"]) | (rec["texel_y"] != ref_rec["texel_y"])
    report(name + " F64_FAST", cmp, kicked, bad)
    assert cmp["differing_pixels_regular"] == 0 and cmp["differing_records_regular"] == 0, cmp
    assert int((bad & ~kicked).sum()) == 0, "a ray with stiffness < 1 differs: the guard band did not hold"
    assert int(bad.sum()) <= max(1, int(1e-5 * bad.size)), f"{int(bad.sum())} kicked rays differ"
    assert abs(st["total_steps"] - ref_st["total_steps"]) <= int(bad.sum()) * sim[0]
    # the fast kernel's own count of kicked rays agrees with the oracle's stiffness (its monitor rounds w up by <= 2^-20)
    assert abs(st["n_kicked"] - int(kicked.sum())) <= max(2, int(1e-4 * kicked.sum()))


def test_full_identity_mode_guard_2(gpu_ctx, oracle, backgrounds):
    """ctx option "guard" = 2: kicked rays are re-integrated too, so EVERY ray of the fast frame equals the
    operation-for-operation kernel's — checked on a full 1080p Interstellar frame and against the oracle's rows."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = backgrounds
    W, H, sim = 1920, 1080, (40000, 100.0, 0.05)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    system = cv.RelativisticSystem(cv.InterstellarMetric(0.1, 1e-4, 1.0), cv.SphericalImage(bp), cv.SphericalImage(bn),
                                   cv.Camera(*cam_args), context=gpu_ctx)
    strict, rec0 = system.render_rows(*sim, 0, H, with_records=True)
    st0 = dict(system.last_stats)
    gpu_ctx.set_option("guard", 2)
    try:
        fast, rec1 = system.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
        st1 = dict(system.last_stats)
    finally:
        gpu_ctx.set_option("guard", 1)
    assert (strict == fast).all()
    for f in ("steps", "side", "texel_x", "texel_y"):
        assert (rec0[f] == rec1[f]).all(), f
    for k in ("total_steps", "n_positive", "n_negative", "n_not_escaped", "n_clamped"):
        assert st0[k] == st1[k], k
    assert st1["n_reintegrated"] >= st1["n_kicked"] > 0
    print(f"[parity] guard=2: {st1['n_reintegrated']} of {W * H} rays re-integrated ({st1['n_kicked']} kicked)")


def test_full_re_integration_list_falls_back_in_line(gpu_ctx):
    """A ray that finds the re-integration list full is re-integrated by the fast kernel itself (plain operators): same
    frame, same counters.  Forced here with the test knob "redo_capacity_limit"."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.noise_background(512, 256, 3), scenes.noise_background(512, 256, 4)
    W, H, sim = 320, 180, (40000, 100.0, 0.05)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    for metric in (cv.EllisMetric(1.0), cv.InterstellarMetric(0.1, 1e-4, 1.0)):
        system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
        strict, rec0 = system.render_rows(*sim, 0, H, with_records=True)
        st0 = dict(system.last_stats)
        gpu_ctx.set_option("guard", 2)
        gpu_ctx.set_option("redo_capacity_limit", 37)
        try:
            fast, rec1 = system.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
            st1 = dict(system.last_stats)
        finally:
            gpu_ctx.set_option("guard", 1)
            gpu_ctx.set_option("redo_capacity_limit", 0)
        assert st1["n_reintegrated"] > 37
        assert (strict == fast).all()
        for f in ("steps", "side", "texel_x", "texel_y"):
            assert (rec0[f] == rec1[f]).all(), f
        for k in ("total_steps", "n_positive", "n_negative", "n_not_escaped", "n_clamped"):
            assert st0[k] == st1[k], k
