"""CURVIS_PRECISION_F64_FAST — fp64 with the right-hand side regrouped around one reciprocal per step
(render_f64_fast.cu) plus the guard band that sends every ray near a decision boundary back through the
operation-for-operation arithmetic.  Its bar, against the same oracle as the parity kernel:

* integer / byte results (RGB8, escape side, step count, texel index, counters): IDENTICAL to the oracle on every ray
  the oracle classifies as regular (oracle/classify.py) and on every ray with stiffness < 1 (the rays
  the guard band covers by construction); chaotic / kicked rays are counted and printed, and at most 1e-5 of all rays
  may differ (measured: 1 pixel in 21 M).  With the context option "guard" = 2 every ray is identical
  (tests/test_gpu_baseline_configs.py::test_full_identity_mode_guard_2);
* floating-point state: final l and p_l within rtol 1e-9 on >= 99.9 % of escaped rays (the rest are
  the kicked rays, where 1 ulp is amplified up to 1e10-fold), end direction within 1e-5 rad
  (BASELINE.json north_star) on >= 99.99 %;
* exotic cases (NaN rays of the Flat metric, NotEscaped, zero iterations, negative delta): exactly
  the parity kernel's output — such rays are re-integrated.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

MAX_DIFFERING_FRACTION = 1e-5
TWO_OVER_PI = np.longdouble(2) / (4 * np.arctan(np.longdouble(1)))      # in x87 long double
END_DIRECTION_TOL_RAD = 1e-5


def _ulps(got, want_ld):
    want = want_ld.astype(np.float64)
    return np.abs((got.astype(np.longdouble) - want_ld) / np.spacing(np.abs(want)).astype(np.longdouble)).astype(np.float64)


def test_primitives_accuracy(gpu_ctx):
    """The reciprocal without the correction step is <= 1 ulp (and correctly rounded on > 99 % of
    operands); sin^2 and sin*cos from one reduction are within 4 ulp."""
    rng = np.random.default_rng(20251017)
    x = np.exp(rng.uniform(-200, 200, 2_000_000)) * rng.choice([-1.0, 1.0], 2_000_000)
    got = gpu_ctx.debug_eval(10, x)
    e = _ulps(got, 1.0 / x.astype(np.longdouble))
    assert e.max() <= 1.0 and (got == 1.0 / x).mean() > 0.99
    th = np.concatenate([rng.uniform(-7, 7, 2_000_000), np.pi + rng.uniform(-1e-3, 1e-3, 50_000), rng.uniform(-1e-6, 1e-6, 50_000),
                         rng.uniform(-1e4, 1e4, 100_000)])
    thl = th.astype(np.longdouble)
    assert _ulps(gpu_ctx.debug_eval(11, th), np.sin(thl) ** 2).max() <= 4.0
    # sin*cos = sin(2 theta)/2 has zeros at multiples of pi/2, where the relative error of any
    # finite-precision reduction is unbounded: absolute bound there, relative elsewhere
    cs, want = gpu_ctx.debug_eval(12, th), np.sin(thl) * np.cos(thl)
    away = np.abs(want) > 1e-3
    assert _ulps(cs[away], want[away]).max() <= 4.0
    assert np.abs(cs[~away] - want[~away].astype(np.float64)).max() <= 1e-18 + 4e-16 * np.abs(th[~away]).max()


def test_interstellar_shape_table_on_device(gpu_ctx):
    """F(x) = x atan x - ln(1+x^2)/2 and G(x) = atan x as the fast kernel evaluates them (table inside
    [2^-10, 2^16), library functions outside): <= 2 ulp inside, <= 6 ulp of the composition outside;
    bit-identical to the host evaluation of the same table."""
    import ctypes as C
    from curvis_b200 import _abi
    rng = np.random.default_rng(7)
    x = np.concatenate([np.exp(rng.uniform(np.log(2.0 ** -10), np.log(2.0 ** 16), 1_000_000)), np.ldexp(1.0, np.arange(-10, 16))])
    xl = x.astype(np.longdouble)
    f, g = gpu_ctx.debug_eval(13, x), gpu_ctx.debug_eval(14, x)
    assert _ulps(f, xl * np.arctan(xl) - np.log1p(xl * xl) / 2).max() <= 2.0
    assert _ulps(g, TWO_OVER_PI * np.arctan(xl)).max() <= 2.0
    hf, hg = np.empty_like(x), np.empty_like(x)
    dp = C.POINTER(C.c_double)
    assert _abi.load_library().curvis_debug_shape_table_host(x.ctypes.data_as(dp), hf.ctypes.data_as(dp), hg.ctypes.data_as(dp), x.size) == 1
    assert f.tobytes() == hf.tobytes() and g.tobytes() == hg.tobytes()
    # outside the table: the reference's own expression x atan x - ln(1 + x*x)/2 through the library
    # functions.  For tiny x its 1 + x*x rounds, so the bar there is absolute (r = rho + m F only
    # sees F's absolute error); for huge x it is relative.
    tiny, huge = np.exp(rng.uniform(np.log(1e-6), np.log(2.0 ** -10), 10_000)), np.exp(rng.uniform(np.log(2.0 ** 16), np.log(1e12), 10_000))
    tl, hl = tiny.astype(np.longdouble), huge.astype(np.longdouble)
    assert np.abs(gpu_ctx.debug_eval(13, tiny).astype(np.longdouble) - (tl * np.arctan(tl) - np.log1p(tl * tl) / 2)).max() <= 2.3e-16
    assert _ulps(gpu_ctx.debug_eval(13, huge), hl * np.arctan(hl) - np.log1p(hl * hl) / 2).max() <= 6.0
    assert _ulps(gpu_ctx.debug_eval(14, np.concatenate([tiny, huge])), TWO_OVER_PI * np.arctan(np.concatenate([tl, hl]))).max() <= 3.0
    assert (gpu_ctx.debug_eval(13, np.array([0.0, -1.0, np.nan])) == 0).all() and (gpu_ctx.debug_eval(14, np.array([0.0, -1.0, np.nan])) == 0).all()


def test_fast_variants_render_the_same_frames(gpu_ctx):
    """ctx option "fast_variant": (sin theta, cos theta) rotated by the step's dtheta (1, default)
    against sin/cos from theta every step (0) — same RGB8, sides, step counts and texels."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    cases = [(cv.EllisMetric(1.0), (40000, 100.0, 0.05)), (cv.InterstellarMetric(0.1, 1e-4, 1.0), (40000, 100.0, 0.05)),
             (cv.FlatSphericalMetric(), (4000, 100.0, 0.05)), (cv.EllisMetric(1.0), (200, 10.0, 0.1))]
    W, H = 320, 180
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    try:
        for metric, sim in cases:
            sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
            out = {}
            for variant in (0, 1):
                gpu_ctx.set_option("fast_variant", variant)
                out[variant] = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
            (f0, r0), (f1, r1) = out[0], out[1]
            bad = (f0 != f1).any(axis=2) | (r0["side"] != r1["side"]) | (r0["steps"] != r1["steps"]) | \
                  (r0["texel_x"] != r1["texel_x"]) | (r0["texel_y"] != r1["texel_y"])
            assert int(bad.sum()) <= int(MAX_DIFFERING_FRACTION * W * H), (type(metric).__name__, sim, int(bad.sum()))
    finally:
        gpu_ctx.set_option("fast_variant", 1)


def test_scheduling_options_do_not_change_a_ray(gpu_ctx):
    """ctx options "longest_first" (pre-pass list of the predicted stragglers, claimed first) and "fast_regs" (96 / 128 registers:
    5 / 4 resident CTAs per SM) only change the ORDER in which rays are integrated: every record — state, steps, side, texel —
    and every counter is byte-identical, and the pre-pass shows up as one more kernel launch.  Frames: the default camera
    (stragglers next to the central row) and a tilted one (the pole-grazing wedge is oblique), batched frames included."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    lib = _abi.load_library()
    bp, bn = scenes.noise_background(1024, 512, 5), scenes.noise_background(1024, 512, 6)
    W, H, sim = 480, 270, (40000, 100.0, 0.05)
    cams = [cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H),
            cv.Camera((0.0, -4.0, 1.1, 2.0), (0.9, 0.3, -0.2), (0.1, 0.2, 1.0), 12.0, 43.0, W, H)]
    try:
        for metric in (cv.EllisMetric(1.0), cv.InterstellarMetric(0.1, 1e-4, 1.0)):
            for cam in cams:
                sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
                out = {}
                # (the third entry: "favoured_slots" — 0: no warp takes the list before the index walk is exhausted, i.e. the path a
                # launch takes whose warps all sit in unfavoured hardware slots; 64: every warp takes it first)
                for lf, regs, slots in ((0, 0, 8), (1, 0, 8), (2, 0, 8), (1, 128, 8), (1, 0, 0), (1, 0, 64), (1, 0, 3)):
                    gpu_ctx.set_option("longest_first", lf)
                    gpu_ctx.set_option("fast_regs", regs)
                    gpu_ctx.set_option("favoured_slots", slots)
                    n0 = lib.curvis_kernel_launch_count()
                    frame, rec = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
                    out[(lf, regs, slots)] = (frame, rec, dict(sysm.last_stats), lib.curvis_kernel_launch_count() - n0)
                f0, r0, s0, n_launch0 = out[(0, 0, 8)]
                assert n_launch0 == 2                                  # render + re-integration
                assert out[(1, 0, 8)][3] == 3 and out[(2, 0, 8)][3] == 3      # + the pre-pass (129,600 rays: under 64 per lane)
                for key, (f, r, st, _) in out.items():
                    assert f.tobytes() == f0.tobytes() and r.tobytes() == r0.tobytes(), (type(metric).__name__, key)
                    for k in ("total_steps", "n_positive", "n_negative", "n_not_escaped", "n_clamped", "n_reintegrated", "n_kicked"):
                        assert st[k] == s0[k], (key, k)
    finally:
        gpu_ctx.set_option("longest_first", 2)
        gpu_ctx.set_option("fast_regs", 0)
        gpu_ctx.set_option("favoured_slots", 8)


def _compare_with_oracle(frame, rec, ref_frame, ref_rec, name):
    n = ref_rec.size
    bad = ((frame != ref_frame).any(axis=2) | (rec["side"] != ref_rec["side"]) | (rec["steps"] != ref_rec["steps"]) |
           (rec["texel_x"] != ref_rec["texel_x"]) | (rec["texel_y"] != ref_rec["texel_y"]))
    n_bad = int(bad.sum())
    assert n_bad <= max(0, int(MAX_DIFFERING_FRACTION * n)), f"{name}: {n_bad} of {n} rays differ from the oracle"
    if "min_abs_sin_theta" in ref_rec.dtype.names:       # live oracle records (the golden fixtures predate the diagnostics)
        from oracle import classify
        chaotic = classify.chaotic_mask(ref_rec)
        with np.errstate(invalid="ignore"):
            kicked = ~(ref_rec["stiffness"] < 1.0)
        print(f"[parity] {name}: {n} rays, chaotic {int(chaotic.sum())}, kicked {int(kicked.sum())}, differing {n_bad} "
              f"(regular {int((bad & ~chaotic).sum())}, stiffness < 1 {int((bad & ~kicked).sum())})")
        assert int((bad & ~chaotic).sum()) == 0, f"{name}: a regular ray differs from the oracle"
        assert int((bad & ~kicked).sum()) == 0, f"{name}: a ray with stiffness < 1 differs from the oracle (guard band)"
    assert rec["p_phi"].tobytes() == ref_rec["p_phi"].tobytes(), f"{name}: p_phi is conserved and never recomputed"
    ok = (ref_rec["side"] != 0) & ~bad & np.isfinite(ref_rec["l"]) & np.isfinite(ref_rec["theta"])
    if ok.any():
        for f in ("l", "p_l"):
            rel = np.abs(rec[f][ok] - ref_rec[f][ok]) / np.maximum(np.abs(ref_rec[f][ok]), 1e-300)
            assert (rel <= 1e-9).mean() >= 0.999, f"{name}: {f} deviates (p99.9 = {np.percentile(rel, 99.9)})"
        a = np.stack([rec["p_l"], rec["p_theta"], rec["p_phi"]], -1)[ok]
        b = np.stack([ref_rec["p_l"], ref_rec["p_theta"], ref_rec["p_phi"]], -1)[ok]
        ang = np.arctan2(np.linalg.norm(np.cross(a, b), axis=-1), (a * b).sum(-1))
        assert (ang <= END_DIRECTION_TOL_RAD).mean() >= 0.9999, f"{name}: end direction"
    return n_bad


@pytest.mark.parametrize("name", ["ellis_c1a_64x36", "ellis_defaults_48x27", "interstellar_defaults_48x27",
                                  "flat_40x30", "ellis_tilted_33x17"])
def test_golden_scenes(gpu_ctx, oracle, name):
    """The golden fixtures of the parity kernel, rendered in fast mode."""
    import curvis_b200 as cv
    import make_golden
    from curvis_b200 import _abi, scenes
    kind, mk, W, H, sim, pos, fwd, up, _ = make_golden.CASES[name]
    _, _, _, bp, bn = make_golden.scene(name)
    metric = {"ellis": lambda: cv.EllisMetric(mk.get("rho", 1.0)), "interstellar": lambda: cv.InterstellarMetric(0.1, 1e-4, 1.0),
              "flat": cv.FlatSphericalMetric}[kind]()
    cam = cv.Camera(pos, fwd, up, scenes.DEFAULT_FOCAL_LENGTH, scenes.DEFAULT_DIAGONAL, W, H)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
    frame, rec = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
    st = sysm.last_stats
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    n_bad = _compare_with_oracle(frame, rec, gold["rgb"], gold["rec"], name)
    if n_bad == 0:
        assert st["total_steps"] == int(gold["total_steps"])
        assert [st["n_positive"], st["n_negative"], st["n_not_escaped"], st["n_clamped"]] == gold["counts"].tolist()
    assert (sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST) == frame).all()


@pytest.mark.parametrize("kind,sim", [("ellis", (200, 10.0, 0.1)), ("ellis", (40000, 100.0, 0.05)),
                                      ("interstellar", (40000, 100.0, 0.05))])
def test_baseline_config_256x144(gpu_ctx, oracle, kind, sim):
    """BASELINE.json configs[0] (C1a / C1b) and the Interstellar default frame against the live oracle."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    W, H = 256, 144
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    metric = cv.EllisMetric(1.0) if kind == "ellis" else cv.InterstellarMetric(0.1, 1e-4, 1.0)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    frame, rec = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
    ref_frame, ref_rec, ref_st = oracle.render_rows(oracle.metric(kind), oracle.camera(*cam_args), oracle.sim(*sim), bp, bn,
                                                    threads=os.cpu_count() or 1)
    n_bad = _compare_with_oracle(frame, rec, ref_frame, ref_rec, f"{kind} {sim}")
    assert abs(sysm.last_stats["total_steps"] - ref_st["total_steps"]) <= n_bad * sim[0]


def test_exotic_rays_match_the_parity_kernel(gpu_ctx):
    """NaN rays (Flat metric through l = 0), NotEscaped, zero iterations, negative step, a camera on
    the polar axis: rays outside the fast step's operand window take the parity step, so the frames
    and the per-ray integers equal the parity kernel's."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.noise_background(128, 64, 1), scenes.noise_background(128, 64, 2)
    cases = [
        (cv.FlatSphericalMetric(), (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 40, 30), (4000, 100.0, 0.05)),
        (cv.EllisMetric(1.0), (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 40, 24), (100, 100.0, 0.05)),
        (cv.EllisMetric(1.0), (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 40, 24), (0, 100.0, 0.05)),
        (cv.EllisMetric(1.0), (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 40, 24), (140, 10.0, 0.1)),
        (cv.EllisMetric(1.0), (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 40, 24), (3000, 100.0, -0.05)),
        (cv.EllisMetric(2.0), ((0.0, 3.0, 1e-9, 0.3), (-1.0, 0.2, 0.1), (0.0, 0.0, 1.0), 15.0, 43.0, 31, 17), (3000, 50.0, 0.05)),
        (cv.InterstellarMetric(0.1, 1e-4, 1.0), ((0.0, 0.00005, 1.2, 0.3), (-1.0, 0.2, 0.1), (0.0, 0.0, 1.0), 15.0, 43.0, 31, 17), (3000, 50.0, 0.05)),
    ]
    for metric, cam_args, sim in cases:
        H = cam_args[-1]
        sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
        f0, r0 = sysm.render_rows(*sim, 0, H, with_records=True)
        s0 = dict(sysm.last_stats)
        f1, r1 = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
        s1 = sysm.last_stats
        name = f"{type(metric).__name__} {sim}"
        assert (f0 == f1).all(), name
        for f in ("side", "steps", "texel_x", "texel_y"):
            assert (r0[f] == r1[f]).all(), (name, f)
        for k in ("total_steps", "n_positive", "n_negative", "n_not_escaped", "n_clamped"):
            assert s0[k] == s1[k], (name, k)


def test_full_4k_frame_against_the_parity_kernel(gpu_ctx, oracle):
    """BASELINE metric config (Ellis 3840x2160 defaults): the fast frame against the parity kernel's
    (itself identical to the oracle on all 8,294,400 pixels, profiles/r01_parity_full_4k_ellis.json),
    row tiles == whole frame, determinism, and strided rows against the oracle directly."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    from curvis_b200.distributed import row_tile
    W, H = 3840, 2160
    sim = (40000, 100.0, 0.05)
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=gpu_ctx)
    strict = sysm.render_image(*sim).copy()
    st_strict = dict(sysm.last_stats)
    fast = sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST).copy()
    st_fast = dict(sysm.last_stats)
    differing = int((strict != fast).any(axis=2).sum())
    assert differing <= int(MAX_DIFFERING_FRACTION * W * H), differing
    assert abs(st_fast["total_steps"] - st_strict["total_steps"]) <= differing * sim[0]
    assert (sysm.render_image(*sim, precision=_abi.PRECISION_F64_FAST) == fast).all()      # deterministic
    tiles = [sysm.render_rows(*sim, *row_tile(H, r, 8), precision=_abi.PRECISION_F64_FAST) for r in range(8)]
    assert (np.concatenate(tiles, axis=0) == fast).all()
    ref, _, _ = oracle.render_rows(oracle.metric("ellis"), oracle.camera(*cam_args), oracle.sim(*sim), bp, bn, row_begin=11, row_end=H,
                                   row_stride=271, threads=os.cpu_count() or 1, with_records=False)
    assert int((ref != fast[11:H:271]).any(axis=2).sum()) <= 1


def test_fast64_rejects_rk4(gpu_ctx):
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp = scenes.noise_background(64, 32, 1)
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 16, 9)
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bp), cam, context=gpu_ctx)
    with pytest.raises(cv.CurvisError) as e:
        sysm.render_image(10, 100.0, 0.05, precision=_abi.PRECISION_F64_FAST, integrator=_abi.INTEGRATOR_RK4)
    assert e.value.code == _abi.ERR_UNSUPPORTED


def test_interstellar_inverse_table_on_device(gpu_ctx):
    """The per-metric table of 1/r^2 and r'/r^3 (as functions of z = |l| - a) that fast_variant 1 reads (csrc/shape_table.h): the
    device evaluation equals the host evaluation of the same table bit for bit (tests/test_abi_host.py holds the host side to
    <= 2 ulp of long double), and the table is rebuilt when the metric parameters change."""
    import ctypes as C
    import curvis_b200 as cv
    from curvis_b200 import _abi
    lib = _abi.load_library()
    rng = np.random.default_rng(11)
    x = np.concatenate([np.exp(rng.uniform(np.log(2.0 ** -44), np.log(2.0 ** 14), 500_000)), np.array([0.0, -1.0, 2.0 ** -45, 1e-300]),
                        np.ldexp(1.0, np.arange(-44, 14))])
    dp = C.POINTER(C.c_double)
    for rho, m in ((1.0, 0.1), (2.0, 0.37), (1.0, 0.1)):
        metric = cv.InterstellarMetric(m, 1e-4, rho)
        y, g, hy, hg = (np.empty_like(x) for _ in range(4))
        mc = metric.as_c()
        _abi.check(lib.curvis_debug_inverse_shape(gpu_ctx.ptr, C.byref(mc), x.ctypes.data_as(dp), y.ctypes.data_as(dp), g.ctypes.data_as(dp), x.size), gpu_ctx.ptr)
        assert lib.curvis_debug_inverse_table_host(rho, m, x.ctypes.data_as(dp), hy.ctypes.data_as(dp), hg.ctypes.data_as(dp), x.size) == 1
        assert y.tobytes() == hy.tobytes() and g.tobytes() == hg.tobytes(), (rho, m)


def test_guard_band_holds_on_random_scenes(gpu_ctx):
    """The guard band's budget was calibrated on eight scenes (tools/guard_study.py).  Here: 16 seeded random scenes the
    calibration never saw — camera position on either side of the throat and at any polar angle, random orientation, focal
    length, throat size, metric, step, escape radius — fast kernel against the operation-for-operation kernel, every ray:
    integers identical wherever the strict kernel's stiffness is < 1; kicked rays counted, at most 1e-5 of all rays differ."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    rng = np.random.default_rng(20261017)
    bp, bn = scenes.noise_background(2048, 1024, 7), scenes.noise_background(2048, 1024, 8)
    W, H = 480, 270
    total, differing_kicked, kicked_total, reintegrated = 0, 0, 0, 0
    for scene in range(16):
        rho = float(rng.uniform(0.5, 3.0))
        metric = cv.EllisMetric(rho) if scene % 2 == 0 else cv.InterstellarMetric(float(rng.uniform(0.03, 0.5)), float(rng.uniform(1e-5, 0.05)), rho)
        l0 = float(rng.uniform(2.0, 12.0) * rng.choice([-1.0, 1.0]))
        pos = (0.0, l0, float(rng.uniform(0.15, np.pi - 0.15)), float(rng.uniform(0.0, 2 * np.pi)))
        fwd = rng.normal(size=3)
        fwd[0] = -abs(fwd[0]) * np.sign(l0) - 0.5 * np.sign(l0)          # roughly towards the throat
        up = rng.normal(size=3)
        cam = cv.Camera(pos, tuple(fwd), tuple(up), float(rng.uniform(10.0, 40.0)), 43.0, W, H)
        sim = (int(rng.integers(3000, 40000)), float(rng.uniform(abs(l0) + 5.0, 120.0)), float(rng.choice([0.02, 0.05, 0.1])))
        system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
        f0, r0 = system.render_rows(*sim, 0, H, with_records=True)
        s0 = dict(system.last_stats)
        f1, r1 = system.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64_FAST)
        s1 = dict(system.last_stats)
        bad = (f0 != f1).any(axis=2) | (r0["steps"] != r1["steps"]) | (r0["side"] != r1["side"]) | \
              (r0["texel_x"] != r1["texel_x"]) | (r0["texel_y"] != r1["texel_y"])
        with np.errstate(invalid="ignore"):
            kicked = ~(r0["stiffness"] < 1.0)
        assert int((bad & ~kicked).sum()) == 0, (scene, type(metric).__name__, pos, sim, int((bad & ~kicked).sum()))
        total += bad.size
        differing_kicked += int(bad.sum())
        kicked_total += int(kicked.sum())
        reintegrated += s1["n_reintegrated"]
        assert abs(s1["total_steps"] - s0["total_steps"]) <= int(bad.sum()) * sim[0]
    print(f"[guard] 16 random scenes, {total} rays: kicked {kicked_total}, re-integrated {reintegrated}, differing (all kicked) {differing_kicked}")
    assert differing_kicked <= max(1, int(1e-5 * total))


def test_step_shares_account_for_every_step(gpu_ctx):
    """curvis_debug_last_step_shares: the per-warp-slot and per-SM step counts of the last launch (what tools/scheduler_shares.py reads
    the schedulers' unfairness from) each add up to the launch's total_steps, re-integration launch included; no slot beyond the
    resident warps of an SM is used."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    bp, bn = scenes.noise_background(512, 256, 3), scenes.noise_background(512, 256, 4)
    W, H, sim = 320, 180, (40000, 100.0, 0.05)
    cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
    for prec in (_abi.PRECISION_F64_FAST, _abi.PRECISION_F64):
        sysm.render_rows(*sim, 0, H, precision=prec)
        slots, sms = gpu_ctx.last_step_shares()
        total = sysm.last_stats["total_steps"]
        assert sum(slots) == total and sum(sms) == total and total > 0
        assert all(s == 0 for s in slots[32:])          # at most 5 CTAs x 4 warps per SM: slots 0..19 (0..31 with headroom)
        assert sum(1 for s in sms if s) > 100           # the launch covered the GPU
