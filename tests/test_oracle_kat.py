"""Pins the CPU oracle (oracle/curvis_oracle.c) against every known-answer test the reference
holds on the render_image path, against SURVEY.md 8c's independent probe values, and against
the committed golden fixtures.  CPU only."""
import math
import os

import numpy as np
import pytest

PI = math.pi
EPS = 2e12 * np.finfo(np.float64).eps   # the reference's own tolerance (algebra.rs:254, :305)


# ---------------------------------------------------------------- src/algebra.rs:143-309
def test_orientation_constructor(oracle):                      # algebra.rs:143-151
    rot, inv, up = oracle.orientation((1, 0, 0), (0, 0, 1))
    assert up.tolist() == [0.0, 0.0, 1.0]


@pytest.mark.parametrize("forward,up,expect_up", [             # algebra.rs:153-176 (exact assert_eq!)
    ((1, 0, 0), (1, 0, 1), (0, 0, 1)),
    ((1, 1, 0), (-1, -1, 1), (0, 0, 1)),
    ((1, 0, 1), (1, 1, 1), (0, 1, 0)),
])
def test_orientation_constructor_non_orthogonal(oracle, forward, up, expect_up):
    _, _, up_o = oracle.orientation(forward, up)
    assert up_o.tolist() == [float(v) for v in expect_up]


@pytest.mark.parametrize("up", [(1, 0, 0), (-1, 0, 0)])        # algebra.rs:178-198 (#[should_panic])
def test_orientation_constructor_parallel_panics(oracle, up):
    with pytest.raises(ValueError):
        oracle.orientation((1, 0, 0), up)


def test_rotation_matrix_xz_is_exact_identity(oracle):         # algebra.rs:200-209
    rot, inv, _ = oracle.orientation((1, 0, 0), (0, 0, 1))
    assert (rot == np.eye(3)).all() and (inv == np.eye(3)).all()


def test_rotation_matrix_times_inverse_is_identity(oracle):    # algebra.rs:211-235
    rng = np.random.default_rng(1)
    for _ in range(200):
        f, u = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        rot, inv, _ = oracle.orientation(f, u)
        np.testing.assert_allclose(rot @ inv, np.eye(3), atol=1e-14)
        # the rotation takes x to forward-hat and z to the orthogonalised up
        np.testing.assert_allclose(rot @ [1, 0, 0], f / np.linalg.norm(f), atol=1e-14)


def test_vector3_from_theta_phi_kats(oracle):                  # algebra.rs:259-282 (13 KATs)
    s = 1.0 / math.sqrt(2.0)
    kats = [
        ((0.0, 0.0), (0, 0, 1)), ((PI / 2, 0.0), (1, 0, 0)), ((PI, 0.0), (0, 0, -1)),
        ((PI / 2, PI / 4), (s, s, 0)), ((-PI / 2, PI / 4), (-s, -s, 0)),
        ((PI / 2, -PI / 4), (s, -s, 0)), ((-PI / 2, -PI / 4), (-s, s, 0)),
        ((PI / 2, PI / 2), (0, 1, 0)), ((-PI / 2, PI / 2), (0, -1, 0)),
        ((PI / 2, 3 * PI / 4), (-s, s, 0)), ((PI / 2, PI), (-1, 0, 0)),
        ((PI / 2, 5 * PI / 4), (-s, -s, 0)), ((PI / 2, 3 * PI / 2), (0, -1, 0)), ((PI / 2, 7 * PI / 4), (s, -s, 0)),
    ]
    for (theta, phi), want in kats:
        got = oracle.vector3_from_theta_phi(theta, phi)
        np.testing.assert_allclose(got, want, atol=1e-15, rtol=0)   # approx's default epsilon is f64::EPSILON-scale


def test_theta_phi_from_vector3_round_trip(oracle):            # algebra.rs:284-309
    rng = np.random.default_rng(2)
    for _ in range(1000):
        theta, phi, r = rng.uniform(0, PI), rng.uniform(0, 2 * PI), rng.uniform(0.1, 5.0)
        v = (r * math.sin(theta) * math.cos(phi), r * math.sin(theta) * math.sin(phi), r * math.cos(theta))
        t2, p2 = oracle.theta_phi_from_vector3(v)
        assert abs(theta - t2) <= EPS and abs(phi - p2) <= EPS


def test_normalize_theta_phi(oracle):                          # algebra.rs:106-116
    assert oracle.normalize_theta_phi(-0.5, 0.25) == (0.5, 0.25 + PI)
    t, p = oracle.normalize_theta_phi(1.0, -0.5)
    assert t == 1.0 and p == math.fmod(-0.5, 2 * PI) + 2 * PI
    t, p = oracle.normalize_theta_phi(1.0, 7.0)
    assert p == math.fmod(7.0, 2 * PI)


# ---------------------------------------------------------------- src/metrics.rs:509-573
def test_photon_normalization_and_direction_in_plane(oracle):  # metrics.rs:512-541
    g = oracle.metric("ellis", rho=1.0)
    pos = (0.0, 5.0, PI / 2, 0.0)
    d = (math.cos(PI / 4), 0.0, math.sin(PI / 4))
    x, p = oracle.new_photon(g, pos, d)
    assert abs(oracle.squared_norm_cov(g, p, pos)) <= 1e-15     # assert_relative_eq!(norm, 0.0)
    np.testing.assert_allclose(oracle.direction(g, p, x), d, rtol=1e-15, atol=1e-15)


def test_photon_stays_null_when_norm_is_taken_at_current_position(oracle):
    """metrics.rs:543-570 evaluates the norm at the INITIAL position (:567) and cannot pass as
    written; the meaningful invariant is the null norm at the photon's
    own position, which forward Euler keeps to O(delta)."""
    g = oracle.metric("ellis", rho=1.0)
    pos = (0.0, 5.0, PI / 2, 0.0)
    d = (math.cos(PI / 4), 0.0, math.sin(PI / 4))
    x, p = oracle.new_photon(g, pos, d)
    for _ in range(100):
        x, p = oracle.step(g, x, p, 0.01)
    assert abs(oracle.squared_norm_cov(g, p, x)) < 1e-3
    assert x[2] == PI / 2 and p[2] == pytest.approx(0.0, abs=1e-15)   # equatorial rays stay equatorial


# ---------------------------------------------------------------- src/images.rs:353-398 (disabled KATs)
@pytest.mark.parametrize("v,want", [
    ((1.234, 0, 0), (PI / 2, 0.0)), ((-1.234, 0, 0), (PI / 2, PI)),
    ((0, 1.234, 0), (PI / 2, PI / 2)), ((0, -1.234, 0), (PI / 2, 3 * PI / 2)),
    ((0, 0, 1.234), (0.0, 0.0)), ((0, 0, -1.234), (PI, 0.0)),
    ((1.234, 1.234, 0), (PI / 2, PI / 4)), ((-1.234, -1.234, 0), (PI / 2, 5 * PI / 4)),
])
def test_theta_phi_of_image_no_orientation(oracle, v, want):
    _, _, theta, phi = oracle.texel_from_vector3(v, 64, 32)
    assert (theta, phi) == want


def test_theta_phi_of_image_with_orientation(oracle):          # images.rs:400-452 (forward = y, up = z)
    _, inv, _ = oracle.orientation((0, 1, 0), (0, 0, 1))
    cases = [((1.234, 0, 0), (PI / 2, 3 * PI / 2)), ((-1.234, 0, 0), (PI / 2, PI / 2)),
             ((0, 1.234, 0), (PI / 2, 0.0)), ((0, -1.234, 0), (PI / 2, PI)), ((1.234, 1.234, 0), (PI / 2, 7 * PI / 4))]
    for v, want in cases:
        _, _, theta, phi = oracle.texel_from_vector3(v, 64, 32, inv_rot=inv)
        assert theta == pytest.approx(want[0], abs=1e-15) and phi == pytest.approx(want[1], abs=1e-15)


def test_texel_mapping(oracle):                                # images.rs:115-121
    W, H = 4096, 2048
    assert oracle.texel_from_vector3((1, 0, 0), W, H)[:2] == (W // 2, H // 2)     # +x: image centre
    assert oracle.texel_from_vector3((-1, 0, 0), W, H)[:2] == (0, H // 2)         # -x: left edge (phi = pi)
    assert oracle.texel_from_vector3((0, 0, 1), W, H)[:2] == (W // 2, 0)          # +z: top row
    x, y, _, _ = oracle.texel_from_vector3((0, 0, -1), W, H)                      # -z: theta = pi -> row H:
    assert y == H                                                                 # the reference panics here


# ---------------------------------------------------------------- cameras.rs
def test_camera_sensor_and_centre_ray(oracle):                 # cameras.rs:107-110, :150-172
    cam = oracle.camera((0, 5, PI / 2, 0), (-1, 0, 0), (0, 0, 1), 15.0, 43.0, 960, 540)
    assert math.hypot(cam.sensor_width, cam.sensor_height) == pytest.approx(43.0, rel=1e-15)
    assert cam.sensor_width / cam.sensor_height == pytest.approx(960 / 540, rel=1e-15)
    assert oracle.outward_vector(cam, 480, 270, world=False).tolist() == [1.0, 0.0, 0.0]
    assert oracle.outward_vector(cam, 480, 270, world=True).tolist() == [-1.0, 0.0, 0.0]
    top_left = oracle.outward_vector(cam, 0, 0, world=False)                      # pixel corner, no +0.5 offset
    want = np.array([15.0, 0.5 * cam.sensor_width, 0.5 * cam.sensor_height])
    np.testing.assert_allclose(top_left, want / np.linalg.norm(want), rtol=1e-15)


# ---------------------------------------------------------------- SURVEY.md 8c probe KATs
def _default_scene(oracle, kind, W=256, H=144):
    from curvis_b200 import scenes
    cam = oracle.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    return oracle.metric(kind), cam, bp, bn


def test_survey_kats_ellis_defaults(oracle):
    g, cam, bp, bn = _default_scene(oracle, "ellis")
    rows = {0: None, 10: None, 20: None, 72: None, 100: None}
    for y in rows:
        _, rec, _ = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn, row_begin=y, row_end=y + 1)
        rows[y] = rec[0]
    r = rows[0][0]
    assert (r["side"], r["steps"]) == (1, 2029)
    assert (r["l"], r["theta"], r["phi"]) == (100.0130035140307, 0.7745759973021238, 2.534370512752706)
    assert (r["p_l"], r["p_theta"], r["p_phi"]) == (1.0165499732647159, 3.0404267705480215, 2.0501957015850265)
    assert (r["texel_x"], r["texel_y"]) == (2028, 997)
    r = rows[20][37]
    assert (r["side"], r["steps"], r["l"], r["theta"], r["phi"], r["p_l"]) == \
        (1, 2031, 100.01458604180449, 0.9193504617324602, 2.695523926135355, 1.0233708288211019)
    assert (r["texel_x"], r["texel_y"]) == (2029, 1005)
    r = rows[10][128]
    assert (r["side"], r["steps"], r["theta"], r["phi"], r["p_l"]) == (1, 2058, PI / 2, 2.6956786224431037, 1.0199099023838596)
    assert (r["texel_x"], r["texel_y"]) == (2048, 1007)
    r = rows[100][200]
    assert (r["side"], r["steps"], r["l"], r["theta"], r["phi"]) == (1, 2008, 100.00913300522924, 2.106575623909251, -2.9113396020524367)
    assert (r["texel_x"], r["texel_y"]) == (2066, 1033)
    r = rows[72][128]   # centre pixel: straight through the throat
    assert (r["side"], r["steps"], r["l"], r["p_l"], r["p_theta"]) == (-1, 2101, -100.04999999999644, -1.0, 0.0)
    assert (r["texel_x"], r["texel_y"]) == (0, 1024)


def test_survey_kats_ellis_c1a_and_interstellar(oracle):
    g, cam, bp, bn = _default_scene(oracle, "ellis")
    _, rec, st = oracle.render_rows(g, cam, oracle.sim(200, 10.0, 0.1), bp, bn, threads=os.cpu_count() or 1)
    r = rec[0, 0]
    assert (r["side"], r["steps"], r["l"], r["theta"], r["phi"], r["p_l"], r["p_theta"]) == \
        (1, 121, 10.020834497802767, 0.5507179175913904, 2.030336665198834, 0.9442389880643146, 1.6854117878746213)
    assert (r["texel_x"], r["texel_y"]) == (1933, 594)
    assert (rec[100, 200]["side"], rec[100, 200]["steps"], rec[100, 200]["texel_x"], rec[100, 200]["texel_y"]) == (1, 134, 2224, 1152)
    assert (rec[72, 128]["side"], rec[72, 128]["steps"]) == (-1, 151)
    assert st["total_steps"] == 4867638 and st["n_not_escaped"] == 12
    g = oracle.metric("interstellar")
    _, rec0, _ = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn, row_begin=0, row_end=1)
    r = rec0[0, 0]
    assert (r["side"], r["steps"], r["l"], r["theta"], r["phi"], r["p_l"]) == \
        (1, 2046, 100.02135227588496, 0.7718020463907622, 2.5300634511071087, 1.014923807588282)
    assert (r["texel_x"], r["texel_y"]) == (2026, 994)
    _, rec72, _ = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn, row_begin=72, row_end=73)
    assert (rec72[0, 128]["side"], rec72[0, 128]["steps"]) == (-1, 2101)


def test_survey_frame_total_ellis_defaults(oracle):
    g, cam, bp, bn = _default_scene(oracle, "ellis")
    _, _, st = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn, threads=os.cpu_count() or 1, with_records=False)
    assert st["total_steps"] == 72225185            # two independent restatements agree
    assert (st["n_positive"], st["n_negative"], st["n_not_escaped"]) == (35555, 1309, 0)


# ---------------------------------------------------------------- golden fixtures
def _golden_names():
    d = os.path.join(os.path.dirname(__file__), "golden")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz"))


@pytest.mark.parametrize("name", _golden_names())
def test_oracle_reproduces_golden(oracle, name):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools"))
    import make_golden
    g, cam, s, bp, bn = make_golden.scene(name)
    rgb, rec, st = oracle.render_rows(g, cam, s, bp, bn, threads=os.cpu_count() or 1)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    assert (rgb == gold["rgb"]).all()
    # bit-exact final states, NaNs included.  The fixtures were written with the ABI-3 record (the first ten fields); they are
    # kept as they are, so the refactored oracle (diagnostics, frames, coordinates) is pinned to the round-1 one.
    for field in gold["rec"].dtype.names:
        assert rec[field].tobytes() == gold["rec"][field].tobytes(), field
    assert st["total_steps"] == int(gold["total_steps"])


# ---------------------------------------------------------------- oracle self-consistency
def test_escape_edge_cases(oracle):
    g = oracle.metric("ellis")
    x, p = oracle.new_photon(g, (0, 5, PI / 2, 0), (1, 0, 0))
    assert oracle.escape_photon(g, x, p, 0.05, 0, 100.0)[:2] == (0, 0)               # zero iterations -> NotEscaped
    assert oracle.escape_photon(g, (0, 101.0, PI / 2, 0), p, 0.05, 10, 100.0)[0] == -2   # systems.rs:122-124 panic
    side, steps, xf, pf = oracle.escape_photon(g, x, p, 0.05, 40000, 100.0)
    assert side == 1 and steps == 1901 and pf[1] == 1.0                              # radial ray: l = 5 + 0.05 k > 100
    side, steps, _, _ = oracle.escape_photon(g, x, p, 0.05, 1900, 100.0)
    assert side == 0 and steps == 1900
