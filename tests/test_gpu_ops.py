"""Op-level tests of the device math primitives through curvis_debug_eval (GPU).

The unguarded Newton-Raphson reciprocal / division / square root of csrc/ieee_f64.cuh must be
bit-identical to IEEE-754 correctly rounded results (numpy on the host) for operands in their
safe window; the in-kernel sincos must stay within 2 ulp of the correctly rounded value (the
CUDA math library's own bound)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 1 << 22


def _operands(rng, n, lo_exp, hi_exp):
    mant = rng.uniform(1.0, 2.0, n)
    exp = rng.integers(lo_exp, hi_exp, n)
    sign = rng.choice([-1.0, 1.0], n)
    return sign * np.ldexp(mant, exp)


def test_unguarded_division_is_correctly_rounded(gpu_ctx):
    rng = np.random.default_rng(11)
    for lo, hi in ((-3, 3), (-60, 60), (-380, 380)):
        a, b = _operands(rng, N, lo, hi), _operands(rng, N, lo, hi)
        got = gpu_ctx.debug_eval(1, a, b)
        assert got.tobytes() == (a / b).tobytes(), f"division differs from IEEE in window 2^[{lo},{hi})"
    # near-one quotients and exactly representable quotients (rounding boundary cases)
    a = _operands(rng, N, -2, 2)
    b = a * (1.0 + rng.integers(-4, 5, N) * 2.0 ** -52)
    assert gpu_ctx.debug_eval(1, a, b).tobytes() == (a / b).tobytes()
    k = rng.integers(1, 1 << 20, N).astype(np.float64)
    m = rng.integers(1, 1 << 20, N).astype(np.float64)
    assert gpu_ctx.debug_eval(1, k * m, m).tobytes() == k.tobytes()
    # the compiler's own operator agrees too (sanity of the comparison itself)
    assert gpu_ctx.debug_eval(7, a, b).tobytes() == (a / b).tobytes()


def test_unguarded_reciprocal_and_sqrt_are_correctly_rounded(gpu_ctx):
    rng = np.random.default_rng(12)
    for lo, hi in ((-3, 3), (-100, 100), (-390, 390)):
        x = _operands(rng, N, lo, hi)
        assert gpu_ctx.debug_eval(0, x).tobytes() == (1.0 / x).tobytes()
        ax = np.abs(x)
        assert gpu_ctx.debug_eval(2, ax).tobytes() == np.sqrt(ax).tobytes()
    sq = rng.integers(1, 1 << 26, N).astype(np.float64)
    assert gpu_ctx.debug_eval(2, sq * sq).tobytes() == sq.tobytes()          # perfect squares
    near = np.nextafter(sq * sq, np.inf)
    assert gpu_ctx.debug_eval(2, near).tobytes() == np.sqrt(near).tobytes()   # just above a perfect square


def test_step_shaped_operands(gpu_ctx):
    """Operands shaped like the Euler step's: r^2 = rho^2 + l^2, sin^2, their products."""
    rng = np.random.default_rng(13)
    l = rng.uniform(-100, 100, N)
    s = np.sin(rng.uniform(0, np.pi, N))
    r2 = 1.0 + l * l
    r = np.sqrt(r2)
    assert gpu_ctx.debug_eval(2, r2).tobytes() == r.tobytes()
    assert gpu_ctx.debug_eval(1, l, r).tobytes() == (l / r).tobytes()
    assert gpu_ctx.debug_eval(0, r2 * (s * s)).tobytes() == (1.0 / (r2 * (s * s))).tobytes()
    c = np.cos(rng.uniform(0, np.pi, N))
    den = r2 * ((s * s) * s)
    assert gpu_ctx.debug_eval(1, c, den).tobytes() == (c / den).tobytes()


def _ulp_error(got, x, fn):
    import mpmath as mp
    mp.mp.dps = 40
    worst = 0.0
    for g, xi in zip(got, x):
        exact = fn(mp.mpf(float(xi)))
        ulp = float(np.spacing(abs(float(exact)))) if exact != 0 else 5e-324
        worst = max(worst, abs(float((mp.mpf(float(g)) - exact) / ulp)))
    return worst


def test_in_kernel_sincos_accuracy(gpu_ctx):
    import mpmath as mp
    rng = np.random.default_rng(14)
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 20000), rng.uniform(-300, 300, 20000), rng.uniform(-1e6, 1e6, 5000),
                        np.arange(-40, 41) * (np.pi / 2), np.array([0.0, 1e-300, -1e-10, 0.5, 1.0, 2.0, 3.0, 1e9])])
    s, c = gpu_ctx.debug_eval(3, x), gpu_ctx.debug_eval(4, x)
    assert _ulp_error(s, x, mp.sin) <= 2.0
    assert _ulp_error(c, x, mp.cos) <= 2.0
    # large arguments fall back to the library's Payne-Hanek path
    big = np.concatenate([rng.uniform(-1e15, 1e15, 2000), np.array([2.0 ** 30, -2.0 ** 40, 1e300])])
    assert _ulp_error(gpu_ctx.debug_eval(5, big), big, mp.sin) <= 2.0
    assert _ulp_error(gpu_ctx.debug_eval(6, big), big, mp.cos) <= 2.0
    # non-finite
    assert np.isnan(gpu_ctx.debug_eval(5, np.array([np.nan, np.inf, -np.inf]))).all()
    # against numpy on a large sample: never more than 2 ulp apart (numpy/glibc is < 1 ulp)
    xx = rng.uniform(-200, 200, N)
    sd = np.abs(gpu_ctx.debug_eval(3, xx) - np.sin(xx)) / np.spacing(np.abs(np.sin(xx)))
    cd = np.abs(gpu_ctx.debug_eval(4, xx) - np.cos(xx)) / np.spacing(np.abs(np.cos(xx)))
    assert sd.max() <= 2.0 and cd.max() <= 2.0


def test_kernel_variants_render_identical_frames(gpu_ctx, oracle):
    """The tuning knobs never change a result: all kernel variants, occupancies and window sizes
    give the same RGB8 / steps / texels, equal to the oracle's."""
    import curvis_b200 as cv
    from curvis_b200 import scenes
    W, H = 160, 90
    bp, bn = scenes.noise_background(512, 256, 1), scenes.noise_background(512, 256, 2)
    cam_args = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    ref, ref_rec, _ = oracle.render_rows(oracle.metric("ellis"), oracle.camera(*cam_args), oracle.sim(40000, 100.0, 0.05), bp, bn,
                                         threads=8)
    ctx = cv.Context([0])
    sysm = cv.RelativisticSystem(cv.EllisMetric(1.0), cv.SphericalImage(bp), cv.SphericalImage(bn), cv.Camera(*cam_args), context=ctx)
    for variant, blocks, window in [(0, 0, 16), (1, 0, 16), (2, 0, 16), (3, 0, 16), (3, 3, 7), (3, 1, 64), (2, 2, 1),
                                    (4, 0, 16), (4, 2, 5), (5, 0, 16), (5, 3, 7), (5, 1, 1), (5, 0, 0)]:
        ctx.set_option("kernel_variant", variant)
        ctx.set_option("blocks_per_sm", blocks)
        ctx.set_option("window", window)
        frame, rec = sysm.render_rows(40000, 100.0, 0.05, 0, H, with_records=True)
        assert (frame == ref).all(), (variant, blocks, window)
        for f in ("steps", "side", "texel_x", "texel_y"):
            assert (rec[f] == ref_rec[f]).all(), (variant, blocks, window, f)
    with pytest.raises(cv.CurvisError):
        ctx.set_option("no_such_option", 1)


def test_strict_interstellar_atan_and_log_on_device(gpu_ctx):
    """atan x and ln(1 + x^2) as the CURVIS_PRECISION_F64 Interstellar step evaluates them (csrc/geodesic_f64.cuh:
    ShapeInterstellar::atan_log): inside the tables bit-identical to the host evaluation of the same tables
    (tests/test_abi_host.py holds that to <= 1.5 ulp of long double), outside them the CUDA library."""
    import ctypes as C
    from curvis_b200 import _abi
    lib = _abi.load_library()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(23)
    x = np.concatenate([np.exp(rng.uniform(np.log(2.0 ** -10), np.log(2.0 ** 16), 1_000_000)), np.ldexp(1.0, np.arange(-10, 16))])
    at, lg = gpu_ctx.debug_eval(17, x), gpu_ctx.debug_eval(18, x)
    y = 1.0 + x * x                                   # the reference's two roundings (metrics.rs:468)
    hat, hlg = np.empty_like(x), np.empty_like(x)
    assert lib.curvis_debug_fn_table_host(0, x.ctypes.data_as(dp), hat.ctypes.data_as(dp), x.size) == 1
    assert lib.curvis_debug_fn_table_host(1, y.ctypes.data_as(dp), hlg.ctypes.data_as(dp), x.size) == 1
    assert at.tobytes() == hat.tobytes() and lg.tobytes() == hlg.tobytes()
    # outside the tables: the library functions (<= 2 ulp)
    out = np.concatenate([np.exp(rng.uniform(np.log(1e-12), np.log(2.0 ** -10), 10_000)), np.exp(rng.uniform(np.log(2.0 ** 16), np.log(1e12), 10_000))])
    ol = out.astype(np.longdouble)
    for got, want in ((gpu_ctx.debug_eval(17, out), np.arctan(ol)), (gpu_ctx.debug_eval(18, out), np.log((1.0 + out * out).astype(np.longdouble)))):
        err = np.abs((got.astype(np.longdouble) - want) / np.spacing(np.abs(want.astype(np.float64))).astype(np.longdouble)).astype(np.float64)
        assert err.max() <= 2.0


def test_strict_kernel_longest_first_does_not_change_a_ray(gpu_ctx):
    """CURVIS_PRECISION_F64 launches of >= 2^15 rays run the longest-first pre-pass too (kernel_variant 5): the listed rays are
    claimed first by the favoured warp slots, the index walk skips them.  Only the ORDER changes: frame, records and counters are
    byte-identical with the list off, with nobody favoured (the list is taken last) and with everybody favoured; the pre-pass
    shows up as one more launch."""
    import curvis_b200 as cv
    from curvis_b200 import _abi, scenes
    lib = _abi.load_library()
    bp, bn = scenes.noise_background(512, 256, 7), scenes.noise_background(512, 256, 8)
    W, H, sim = 320, 180, (40000, 100.0, 0.05)
    try:
        for metric in (cv.EllisMetric(1.0), cv.InterstellarMetric(0.1, 1e-4, 1.0)):
            for cam in (cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H),
                        cv.Camera((0.0, -4.0, 1.1, 2.0), (0.9, 0.3, -0.2), (0.1, 0.2, 1.0), 12.0, 43.0, W, H)):
                sysm = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=gpu_ctx)
                out = {}
                for lf, slots in ((0, 8), (2, 8), (1, 0), (1, 64), (1, 5)):
                    gpu_ctx.set_option("longest_first", lf)
                    gpu_ctx.set_option("favoured_slots", slots)
                    n0 = lib.curvis_kernel_launch_count()
                    frame = sysm.render_rows(*sim, 0, H, precision=_abi.PRECISION_F64).copy()
                    out[(lf, slots)] = (frame, dict(sysm.last_stats), lib.curvis_kernel_launch_count() - n0)
                f0, s0, n_launch0 = out[(0, 8)]
                assert n_launch0 == 1 and out[(2, 8)][2] == 2 and out[(1, 0)][2] == 2
                for key, (f, st, _) in out.items():
                    assert f.tobytes() == f0.tobytes(), (type(metric).__name__, key)
                    for k in ("total_steps", "n_positive", "n_negative", "n_not_escaped", "n_clamped"):
                        assert st[k] == s0[k], (key, k)
                # records launches keep the index order (no pre-pass) and agree with the frame above
                frame_r, rec = sysm.render_rows(*sim, 0, H, with_records=True, precision=_abi.PRECISION_F64)
                assert frame_r.tobytes() == f0.tobytes() and int(rec["steps"].sum()) == s0["total_steps"]
    finally:
        gpu_ctx.set_option("longest_first", 2)
        gpu_ctx.set_option("favoured_slots", 8)
