"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/curvis_gpu.h
declares, its host-side setup helpers equal the oracle bit for bit, its error behaviour mirrors
the reference's panics, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "curvis_gpu.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(curvis_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from curvis_b200 import _abi
    names = declared_functions()
    assert len(names) >= 13
    out = subprocess.check_output(["nm", "-D", "--defined-only", _abi.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (curvis_[a-z0-9_]+)", out))
    assert set(names) <= exported, f"missing: {set(names) - exported}"
    assert set(names) == set(_abi.EXPORTED_SYMBOLS)
    for n in names:
        getattr(lib, n)
    assert lib.curvis_abi_version() == _abi.ABI_VERSION == 4


def test_struct_layouts_match_header():
    from curvis_b200 import _abi
    assert C.sizeof(_abi.CurvisMetric) == 32
    assert C.sizeof(_abi.CurvisCamera) == 4 * 8 + 9 * 8 + 3 * 8 + 8
    assert C.sizeof(_abi.CurvisSim) == 56
    assert C.sizeof(_abi.CurvisStats) == 10 * 8
    assert C.sizeof(_abi.CurvisRayRecord) == 80 == np.dtype(_abi.RAY_RECORD_DTYPE).itemsize


def test_library_embeds_sm100a_kernels(lib):
    from curvis_b200 import _abi
    out = subprocess.check_output(["cuobjdump", "-lelf", _abi.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_orientation_and_camera_equal_oracle_bitwise(lib, oracle):
    import curvis_b200 as cv
    rng = np.random.default_rng(7)
    cases = [((1, 0, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1)), ((1, 1, 0), (-1, -1, 1)), ((1, 0, 1), (1, 1, 1))]
    cases += [(tuple(rng.uniform(-1, 1, 3)), tuple(rng.uniform(-1, 1, 3))) for _ in range(50)]
    for f, u in cases:
        o = cv.Orientation(f, u)
        rot, inv, up = oracle.orientation(f, u)
        assert o.rotation_matrix().tobytes() == rot.tobytes()
        assert o.inverse_rotation_matrix().tobytes() == inv.tobytes()
        assert o.up().tobytes() == up.tobytes()
        cam = cv.Camera((0, 5, 1.0, 0.5), f, u, 15.0, 43.0, 1920, 1080).as_c()
        ocam = oracle.camera((0, 5, 1.0, 0.5), f, u, 15.0, 43.0, 1920, 1080)
        assert bytes(cam) == bytes(ocam)


def test_reference_orientation_kats_through_the_abi(lib):     # algebra.rs:143-209 via the product
    import curvis_b200 as cv
    assert (cv.Orientation((1, 0, 0), (0, 0, 1)).rotation_matrix() == np.eye(3)).all()
    assert cv.Orientation((1, 0, 0), (1, 0, 1)).up().tolist() == [0, 0, 1]
    assert cv.Orientation((1, 1, 0), (-1, -1, 1)).up().tolist() == [0, 0, 1]
    assert cv.Orientation((1, 0, 1), (1, 1, 1)).up().tolist() == [0, 1, 0]
    assert cv.Orientation((1, 1, 0), (-1, -1, 1)).forward().tolist() == [1, 1, 0]   # forward kept as given


def test_error_behaviour_mirrors_reference_panics(lib):
    import curvis_b200 as cv
    from curvis_b200 import _abi
    with pytest.raises(cv.CurvisError) as e:                   # algebra.rs:19-21
        cv.Orientation((1, 0, 0), (-2, 0, 0))
    assert e.value.code == _abi.ERR_PARALLEL_VECTORS and "parallel" in e.value.message
    for bad in (dict(focal_length=0.0), dict(sensor_diagonal=-1.0), dict(resolution_width=0)):   # cameras.rs:94-102
        kw = dict(focal_length=15.0, sensor_diagonal=43.0, resolution_width=16, resolution_height=9)
        kw.update(bad)
        with pytest.raises(cv.CurvisError) as e:
            cv.Camera((0, 5, 1, 0), (-1, 0, 0), (0, 0, 1), **kw)
        assert e.value.code == _abi.ERR_INVALID_ARGUMENT
    with pytest.raises(cv.CurvisError) as e:                   # metrics.rs:407-409
        cv.EllisMetric(0.0)
    assert e.value.code == _abi.ERR_INVALID_METRIC and "rho" in e.value.message
    for m, a, rho in ((0, 1, 1), (1, -1, 1), (1, 1, 0)):        # metrics.rs:443-456
        with pytest.raises(cv.CurvisError):
            cv.InterstellarMetric(m, a, rho)
    cv.InterstellarMetric(0.1, 1e-4, 1.0)
    cv.FlatSphericalMetric()


def test_null_arguments_do_not_crash(lib):
    from curvis_b200 import _abi
    assert lib.curvis_orientation(None, None, None, None, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.curvis_camera_init(None, None, None, None, 1.0, 1.0, 1, 1) == _abi.ERR_INVALID_ARGUMENT
    assert lib.curvis_metric_validate(None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.curvis_ctx_create(None, 0, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.curvis_set_background(None, 1, None, 1, 1, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.curvis_render_image(None, None, None, None, None, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.curvis_ctx_device_count(None) == 0
    lib.curvis_ctx_destroy(None)
    assert isinstance(lib.curvis_last_error(None), bytes)


def test_no_cpu_fallback_without_a_gpu(lib):
    """Without a CUDA device the product refuses to run: no CPU path exists behind the ABI."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import curvis_b200 as cv
    from curvis_b200 import _abi
    with pytest.raises(cv.CurvisError) as e:
        cv.Context()
    assert e.value.code == _abi.ERR_NO_DEVICE
    assert "no CPU fallback" in e.value.message


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under curvis_b200/ may reference it."""
    pkg = os.path.join(ROOT, "curvis_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "curvis_oracle" not in text and "liboracle" not in text, f


def test_spherical_image_conversions(lib):
    import curvis_b200 as cv
    rgb = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    img = cv.SphericalImage(rgb)
    assert img.rgba8.shape == (2, 3, 4) and (img.rgba8[..., 3] == 255).all() and (img.rgba8[..., :3] == rgb).all()
    assert img.dimensions() == (3, 2)
    assert (img.orientation().inverse_rotation_matrix() == np.eye(3)).all()
    luma = cv.SphericalImage(np.full((2, 2), 9, np.uint8))
    assert (luma.rgba8[..., :3] == 9).all()
    with pytest.raises(TypeError):
        cv.SphericalImage(np.zeros((2, 2, 3), np.float32))


def test_interstellar_shape_table_host(lib):
    """The piecewise degree-5 table of F(x) = x atan x - ln(1+x^2)/2 and G(x) = (2/pi) atan x that
    CURVIS_PRECISION_F64_FAST uploads (csrc/shape_table.h; replaces the atan + ln of
    InterstellarMetric::r / r_derivative, reference src/metrics.rs:461-485), evaluated on the host
    with the kernel's arithmetic, against x87 long double: <= 2 ulp, i.e. the class of a libm."""
    import ctypes as C
    import numpy as np
    rng = np.random.default_rng(20251017)
    edges = np.ldexp(1.0, np.arange(-10, 17))
    x = np.concatenate([np.exp(rng.uniform(np.log(2.0 ** -10), np.log(2.0 ** 16), 400_000)),
                        edges[:-1], np.nextafter(edges[1:], 0.0), np.nextafter(edges[:-1], np.inf),
                        np.ldexp(1.0 + np.arange(128) / 128.0, 3), np.nextafter(np.ldexp(1.0 + np.arange(1, 129) / 128.0, 3), 0.0)])
    f, g = np.empty_like(x), np.empty_like(x)
    dp = C.POINTER(C.c_double)
    assert lib.curvis_debug_shape_table_host(x.ctypes.data_as(dp), f.ctypes.data_as(dp), g.ctypes.data_as(dp), x.size) == 1
    xl = x.astype(np.longdouble)
    TWO_OVER_PI = np.longdouble(2) / (4 * np.arctan(np.longdouble(1)))      # in x87 long double
    want_f, want_g = xl * np.arctan(xl) - np.log1p(xl * xl) / 2, TWO_OVER_PI * np.arctan(xl)

    def ulps(got, want):
        return np.abs((got.astype(np.longdouble) - want) / np.spacing(np.abs(want.astype(np.float64))).astype(np.longdouble)).astype(np.float64)

    assert ulps(f, want_f).max() <= 2.0
    assert ulps(g, want_g).max() <= 2.0
    out = np.array([0.0, 2.0 ** -11, 2.0 ** 16, np.inf, np.nan, -1.0])
    fo, go = np.empty_like(out), np.empty_like(out)
    assert lib.curvis_debug_shape_table_host(out.ctypes.data_as(dp), fo.ctypes.data_as(dp), go.ctypes.data_as(dp), out.size) == 0
    assert np.isnan(fo).all() and np.isnan(go).all()


@pytest.mark.parametrize("rho,m", [(1.0, 0.1), (2.5, 0.03), (0.7, 1.5)])
def test_interstellar_inverse_table_host(lib, rho, m):
    """The per-metric table the default fast kernel reads (csrc/shape_table.h: build_interstellar_inverse_table):
    U(z) = 1 / (rho + m F(x))^2 and H(z) = G(x) / (rho + m F(x))^3 at x = 2 z / (pi m), z = |l| - a — the two combinations the
    regrouped step needs — evaluated on the host with the kernel's arithmetic, against x87 long double.  Degree 5 on 128
    intervals per binade: U ~ z^-2 and H ~ z^-3 have seventh Taylor coefficients 7 and 28, so the interpolation error reaches 8
    resp. 28-48 units of 2^-53 (relative) at the START of a binade and falls 64-fold towards its end — r.m.s. 0.8 resp. 2.9
    units, the size of the rounding of the Horner evaluation.  (256 intervals per binade hold both below 2 units everywhere,
    but the table then falls out of the L1 and the 4K frame costs 5 % more; the fast kernel's deviation from the
    operation-for-operation kernel is the same with either: tools/guard_study_interstellar.py.)  Everything below 2^-44 — zero
    and negative z, the plateau |l| <= a of metrics.rs:470 / :482 — reads r = rho, r' = 0."""
    import ctypes as C
    import numpy as np
    rng = np.random.default_rng(20261017)
    edges = np.ldexp(1.0, np.arange(-44, 15))
    z = np.concatenate([np.exp(rng.uniform(np.log(2.0 ** -44), np.log(2.0 ** 14), 400_000)),
                        np.exp(rng.uniform(np.log(0.01), np.log(200.0), 400_000)),
                        edges[:-1], np.nextafter(edges[1:], 0.0), np.nextafter(edges[:-1], np.inf)])
    y, g = np.empty_like(z), np.empty_like(z)
    dp = C.POINTER(C.c_double)
    assert lib.curvis_debug_inverse_table_host(rho, m, z.ctypes.data_as(dp), y.ctypes.data_as(dp), g.ctypes.data_as(dp), z.size) == 1
    PI = 4 * np.arctan(np.longdouble(1))
    xl = z.astype(np.longdouble) * (np.longdouble(2) / (PI * np.longdouble(m)))
    r = np.longdouble(rho) + np.longdouble(m) * (xl * np.arctan(xl) - np.log1p(xl * xl) / 2)
    want_y = 1 / (r * r)
    want_g = (np.longdouble(2) / PI) * np.arctan(xl) / (r * r * r)

    def units(got, want):      # relative error in units of 2^-53
        return (np.abs((got.astype(np.longdouble) - want) / want) * np.longdouble(2.0 ** 53)).astype(np.float64)

    eu, eh = units(y, want_y), units(g, want_g)
    assert eu.max() <= 16.0 and np.sqrt((eu ** 2).mean()) <= 2.0, (eu.max(), np.sqrt((eu ** 2).mean()))
    assert eh.max() <= 56.0 and np.sqrt((eh ** 2).mean()) <= 5.5, (eh.max(), np.sqrt((eh ** 2).mean()))
    low = np.array([0.0, -0.0, -1.0, -1e300, 2.0 ** -45, 5e-324])
    yo, go = np.empty_like(low), np.empty_like(low)
    assert lib.curvis_debug_inverse_table_host(rho, m, low.ctypes.data_as(dp), yo.ctypes.data_as(dp), go.ctypes.data_as(dp), low.size) == 1
    assert (yo == float(1 / (np.longdouble(rho) * np.longdouble(rho)))).all() and (go == 0.0).all()
    out = np.array([2.0 ** 14, 1e30])
    yo, go = np.empty_like(out), np.empty_like(out)
    assert lib.curvis_debug_inverse_table_host(rho, m, out.ctypes.data_as(dp), yo.ctypes.data_as(dp), go.ctypes.data_as(dp), out.size) == 0
    assert np.isnan(yo).all()


def test_strict_interstellar_atan_and_log_tables_host(lib):
    """atan and ln as the CURVIS_PRECISION_F64 Interstellar step evaluates them (csrc/shape_table.h: degree-5 pieces, 128 per
    binade), on the host with the kernel's arithmetic against x87 long double: <= 1.5 ulp for atan on [2^-10, 2^16); for ln on
    [1, 2^33) <= 1.6 ulp or, next to 1 where ln -> 0, 2^-55 absolute (ln(1 + x^2) enters r as m ln / 2 beside rho: only its
    absolute error is seen, and the reference's own 1 + x*x has already rounded x^2 to 2^-53 there).  Arguments outside the tables are refused."""
    import ctypes as C
    import numpy as np
    rng = np.random.default_rng(5)
    dp = C.POINTER(C.c_double)

    def run(which, x):
        out = np.empty_like(x)
        rc = lib.curvis_debug_fn_table_host(which, x.ctypes.data_as(dp), out.ctypes.data_as(dp), x.size)
        return rc, out

    x = np.concatenate([np.exp(rng.uniform(np.log(2.0 ** -10), np.log(2.0 ** 16), 1_000_000)), np.ldexp(1.0, np.arange(-10, 16)),
                        np.nextafter(np.ldexp(1.0, np.arange(-9, 17)), 0.0)])
    rc, at = run(0, x)
    assert rc == 1
    want = np.arctan(x.astype(np.longdouble))
    err = np.abs((at.astype(np.longdouble) - want) / np.spacing(np.abs(want.astype(np.float64))).astype(np.longdouble)).astype(np.float64)
    assert err.max() <= 1.5, err.max()        # (the CUDA library documents 2 ulp for atan, 1 for log)
    y = np.concatenate([np.exp(rng.uniform(0.0, np.log(2.0 ** 33), 1_000_000)), 1.0 + np.exp(rng.uniform(np.log(2.0 ** -52), 0.0, 200_000)),
                        np.array([1.0, np.nextafter(1.0, 2.0), 2.0, np.nextafter(2.0 ** 33, 0.0)])])
    rc, lg = run(1, y)
    assert rc == 1
    want = np.log(y.astype(np.longdouble))
    abs_err = np.abs(lg.astype(np.longdouble) - want).astype(np.float64)
    ulp = np.spacing(np.abs(want.astype(np.float64)))
    assert (abs_err <= np.maximum(1.6 * ulp, 2.0 ** -55)).all(), (abs_err / np.maximum(1.6 * ulp, 2.0 ** -55)).max()
    for which, bad in ((0, np.array([2.0 ** -11, 2.0 ** 16, 0.0, -1.0])), (1, np.array([0.5, 2.0 ** 33, 0.0, -2.0]))):
        rc, out = run(which, bad)
        assert rc == 0 and np.isnan(out).all()


def test_quotient_from_a_correctly_rounded_reciprocal_is_correctly_rounded():
    """The CURVIS_PRECISION_F64 Interstellar step forms x = 2 (|l| - a) / (pi m) as q0 = a y; rem = fma(-b, q0, a); q = fma(rem, y, q0)
    with y = RN(1 / b) supplied by the host (csrc/geodesic_f64.cuh: ShapeInterstellar::eval_fast): that is RN(a / b).  Checked here
    with exact rational arithmetic standing in for the two fused operations."""
    from fractions import Fraction
    import numpy as np
    rng = np.random.default_rng(99)
    a = np.ldexp(rng.uniform(1.0, 2.0, 20_000), rng.integers(-40, 10, 20_000))
    b = np.concatenate([np.ldexp(rng.uniform(1.0, 2.0, 19_000), rng.integers(-20, 20, 19_000)),
                        np.pi * rng.uniform(1e-3, 10.0, 1_000)])                 # pi m for plausible m
    for ai, bi in zip(a.tolist(), b.tolist()):
        y = 1.0 / bi
        q0 = ai * y
        rem = float(Fraction(ai) - Fraction(bi) * Fraction(q0))                   # one rounding (the first FMA; nearly always exact)
        q = float(Fraction(rem) * Fraction(y) + Fraction(q0))                     # one rounding (the second FMA)
        assert q == ai / bi
