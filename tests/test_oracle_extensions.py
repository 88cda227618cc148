"""CPU tests of the oracle's restatements of the extensions that have no reference counterpart —
each restatement IS the oracle of its CUDA kernel (tests/test_gpu_extensions.py):

  * CURVIS_FRAME_WORLD / _WORLD_QUIRK: escaped_photon_to_world_direction (src/systems.rs:144-187) applied to the photon of
    render_image, and the one cross-check the reference itself offers between its two renderers: for an equatorial photon
    the rotated direction IS what compute_escape_angle (src/systems.rs:203-261) turns into the table of
    render_image_efficient;
  * CURVIS_INTEGRATOR_EULER_ADAPTIVE and CURVIS_COORDINATES_CARTESIAN: how many of the chaotic rays become regular.
"""
import math
import os

import numpy as np
import pytest

PI = math.pi


def _scene(oracle, kind="ellis", W=96, H=54):
    from curvis_b200 import scenes
    bp, bn = scenes.decodable_background(2048, 1024), scenes.decodable_background(2048, 1024, negative=True)
    cam = oracle.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    return oracle.metric(kind), cam, bp, bn


def _angle_like_compute_escape_angle(d):
    """src/systems.rs:246-251 on a world direction: normalise, vx, vy, angle in [0, 2 pi)."""
    n = math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    w = (d[0] / n, d[1] / n, d[2] / n)
    vx = (w[0] * 1.0 + w[1] * 0.0) + w[2] * 0.0
    vy = (w[0] * 0.0 + w[1] * 1.0) + w[2] * 0.0
    return math.acos(vx) if vy >= 0.0 else 2.0 * PI - math.acos(vx)


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_world_frame_of_equatorial_photons_is_the_table_function(oracle, kind):
    """Cross-renderer check.  A photon fired in the equatorial plane from (l, pi/2, 0) with tangent direction
    (cos a, 0, sin a) is both a ray of render_image (a pixel with no theta component) and a sample of the table of
    render_image_efficient.  CURVIS_FRAME_WORLD_QUIRK — escaped_photon_to_world_direction on the escaped photon — must
    reproduce compute_escape_angle's angle bit for bit; CURVIS_FRAME_WORLD (frame_field_33 instead of _22 for the phi
    component, the fix of metrics.rs:347) must give the same bits too on the equator, where sin(theta) = 1 exactly."""
    from curvis_b200 import _abi
    g = oracle.metric(kind)
    l0 = 5.0
    for alpha in np.linspace(-0.1 * PI, 1.1 * PI, 97):
        d = (math.cos(alpha), 0.0, math.sin(alpha))
        x, p = oracle.new_photon(g, (0.0, l0, PI / 2.0, 0.0), d)
        side, steps, xf, pf = oracle.escape_photon(g, x, p, 0.05, 40000, 100.0)
        side2, angle, steps2 = oracle.compute_escape_angle(g, l0, float(alpha), 0.05, 40000, 100.0)
        assert (side, steps) == (side2, steps2)
        if side == 0:
            continue
        assert xf[2] == PI / 2.0                                  # theta never moves: cos(pi/2) p_phi^2 / ... adds 6e-17 * ... = 0? see below
        for frame in (_abi.FRAME_WORLD_QUIRK, _abi.FRAME_WORLD):
            w = oracle.lookup_direction(g, xf, pf, frame)
            assert _angle_like_compute_escape_angle(w) == angle, (alpha, frame)


def test_world_frame_pixels_against_the_efficient_renderer(oracle):
    """The same cross-check at frame level: the pixels of render_image whose camera ray has no theta component (the central
    column of the default camera) carry, in CURVIS_FRAME_WORLD_QUIRK, the escape angle that render_image_efficient
    interpolates from its table for that pixel — equal up to the table's interpolation error."""
    import ctypes as C
    from curvis_b200 import _abi
    g, cam, bp, bn = _scene(oracle, W=96, H=54)
    sim = oracle.sim(40000, 100.0, 0.05)
    _, _, alpha, angle, space = oracle.render_image_efficient(g, cam, sim, bp, bn, debug=True)
    col = 96 // 2                                                # w = x / W - 0.5 = 0: no theta component
    worst = 0.0
    checked = 0
    for row in range(54):
        d = oracle.outward_vector(cam, col, row)
        assert d[1] == 0.0
        x, p = oracle.new_photon(g, tuple(cam.position), d)
        side, steps, xf, pf = oracle.escape_photon(g, x, p, 0.05, 40000, 100.0)
        if side == 0 or space[row, col] not in (1.0, -1.0):
            continue
        w = oracle.lookup_direction(g, xf, pf, _abi.FRAME_WORLD_QUIRK)
        a = _angle_like_compute_escape_angle(w)
        if d[2] < 0.0:
            a = 2.0 * PI - a            # the table holds sin(alpha) >= 0; the efficient renderer mirrors the orbit by flipping its axis
        assert float(space[row, col]) == float(side)
        worst = max(worst, abs(a - angle[row, col]))
        checked += 1
    assert checked >= 40
    print(f"[world frame] central column: brute-force escape angle vs the efficient renderer's interpolated one, worst |diff| = {worst:.2e} rad over {checked} pixels")
    assert worst < 2e-3, worst                                   # the table's interpolation error (measured 5e-4; a texel of an 8192-wide image is 7.7e-4)


def test_world_frame_off_axis_rays_differ_from_the_efficient_renderer_by_the_parallax(oracle):
    """Documented limit: escaped_photon_to_world_direction rotates the tangent frame by the MINIMAL
    rotation x -> position, which maps the theta / phi components onto the world axes correctly only on the equator; and
    render_image_efficient places every orbit in the plane spanned by the camera direction and the pixel direction READ AS
    WORLD VECTORS, while (theta, phi) are polar coordinates about the world z axis.  So away from the central column the two
    renderers look up different texels by construction — here the angle between their lookup directions is measured and
    shown to be of the order of the pixel's off-axis angle, not of a rounding error."""
    from curvis_b200 import _abi
    g, cam, bp, bn = _scene(oracle, W=96, H=54)
    # brute force, world frame, whole frame
    sim_w = oracle.sim(40000, 100.0, 0.05, frame=_abi.FRAME_WORLD)
    rgb_w, rec_w, _ = oracle.render_rows(g, cam, sim_w, bp, bn, threads=os.cpu_count() or 1)
    rgb_e, _ = oracle.render_image_efficient(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn)
    same = (rgb_w == rgb_e).all(axis=2)
    print(f"[world frame] identical pixels brute-force(WORLD) vs efficient: {same.mean():.4f}; central column: {same[:, 48].mean():.4f}")
    assert same.mean() < 0.5                                     # they are different images away from the equatorial column


def test_adaptive_step_equals_euler_far_from_the_poles_and_regularises_the_rest(oracle):
    from curvis_b200 import _abi
    from oracle import classify
    g, cam, bp, bn = _scene(oracle, W=128, H=72)
    base = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn, threads=os.cpu_count() or 1)
    # an unreachable tolerance never cuts a step: bit-identical records
    loose = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05, integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=1e300),
                               bp, bn, threads=os.cpu_count() or 1)
    assert (base[0] == loose[0]).all() and base[1].tobytes() == loose[1].tobytes() and base[2]["total_steps"] == loose[2]["total_steps"]
    chaotic0 = classify.chaotic_mask(base[1]).mean()
    tight = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05, integrator=_abi.INTEGRATOR_EULER_ADAPTIVE, step_tolerance=0.01),
                               bp, bn, threads=os.cpu_count() or 1)
    chaotic1 = classify.chaotic_mask(tight[1]).mean()
    print(f"[adaptive] chaotic fraction {chaotic0:.4f} -> {chaotic1:.4f} at step_tolerance 0.01; steps {base[2]['total_steps']} -> {tight[2]['total_steps']}")
    assert chaotic0 > 0.15 and chaotic1 < 0.02
    assert np.nanmax(tight[1]["stiffness"]) <= 0.01 ** 2 * (1 + 1e-9)      # no step advanced phi by more than the tolerance
    assert (np.abs(tight[1]["p_l"][tight[1]["side"] != 0]) <= 1.05).all()
    assert tight[2]["total_steps"] < 1.5 * base[2]["total_steps"]


def test_cartesian_coordinates_have_no_chaotic_rays_and_agree_on_the_regular_ones(oracle):
    from curvis_b200 import _abi
    from oracle import classify
    g, cam, bp, bn = _scene(oracle, W=128, H=72)
    base = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05), bp, bn, threads=os.cpu_count() or 1)
    cart = oracle.render_rows(g, cam, oracle.sim(40000, 100.0, 0.05, coordinates=_abi.COORDINATES_CARTESIAN), bp, bn, threads=os.cpu_count() or 1)
    r0, r1 = base[1], cart[1]
    esc = r1["side"] != 0
    assert (np.abs(r1["p_l"][esc]) <= 1.05).all()               # nothing is kicked: |p_l| -> 1 + O(delta) for every ray
    regular = ~classify.chaotic_mask(r0)
    assert (r0["side"][regular] == r1["side"][regular]).all()
    # same scheme, same step, different chart: end directions agree to the Euler discretisation error, O(delta)
    def direction(rec):
        r = np.sqrt(1.0 + rec["l"] ** 2)
        s = np.sin(rec["theta"])
        return np.stack([rec["p_l"], rec["p_theta"] / r, rec["p_phi"] / (r * s * s)], -1)
    a, b = direction(r0)[regular], direction(r1)[regular]
    ang = np.arccos(np.clip((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1)), -1, 1))
    print(f"[cartesian] chaotic fraction {1 - regular.mean():.4f} -> 0; regular rays: median angle to the spherical result {np.median(ang):.2e} rad, p99 {np.quantile(ang, 0.99):.2e}")
    assert np.median(ang) < 2e-3 and np.quantile(ang, 0.9) < 1e-2      # (rays the survey rule calls regular but that pass within ~0.05 of a pole deviate more)
    # angular momentum is conserved exactly by the scheme: p_phi = J_z never changes
    assert np.allclose(r1["p_phi"], r0["p_phi"], rtol=1e-13, atol=1e-15)
