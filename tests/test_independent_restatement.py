"""A second, independent restatement of the reference's per-ray arithmetic — written in Python
straight from the formulas of src/metrics.rs:223-334, src/cameras.rs:107-172, src/systems.rs:115-139
and src/images.rs:115-121, scalar floats, different structure from oracle/curvis_oracle.c — must
agree with the C oracle BIT FOR BIT (both run on glibc's libm through the same process).  The
reference cannot be executed here, so two independently written restatements agreeing to the last
bit is the strongest available check that the oracle says what the reference says."""
import math

import numpy as np
import pytest

PI = math.pi


class Ellis:
    def __init__(self, rho): self.rho = rho
    def r(self, l): return math.sqrt(self.rho * self.rho + l * l)
    def r2(self, l): return self.rho * self.rho + l * l
    def dr(self, l): return l / self.r(l)


class Interstellar:
    def __init__(self, m, a, rho): self.m, self.a, self.rho = m, a, rho
    def _x(self, l): return 2.0 * (abs(l) - self.a) / (PI * self.m)
    def r(self, l):
        if abs(l) > self.a:
            x = self._x(l)
            return self.rho + self.m * (x * math.atan(x) - math.log(1.0 + x * x) / 2.0)
        return self.rho
    def r2(self, l): return self.r(l) * self.r(l)
    def dr(self, l):
        if abs(l) > self.a:
            return (2.0 / PI) * math.copysign(1.0, l) * math.atan(self._x(l))
        return 0.0


def normalize(v):
    n = math.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    return [v[0] / n, v[1] / n, v[2] / n]


def pixel_ray(rot, focal, sw, sh, W, H, px, py):
    h = 0.5 - (py / H)
    w = (px / W) - 0.5
    v = normalize([focal * 1.0, -sw * w, sh * h])
    return [(rot[i][0] * v[0] + rot[i][1] * v[1]) + rot[i][2] * v[2] for i in range(3)]


def trace(metric, pos, direction, delta, max_iter, R):
    d = normalize(direction)
    t, l, th, ph = pos
    p_t, p_l, p_th, p_ph = 1.0, d[0], d[1] * metric.r(l), d[2] * metric.r(l) * math.sin(th)
    for k in range(max_iter):
        s = math.sin(th)
        g22c, g33c = 1.0 / metric.r2(l), 1.0 / (metric.r2(l) * (s * s))
        dl, dth, dph = p_l * 1.0, p_th * g22c, p_ph * g33c
        b2 = p_th * p_th + (p_ph * p_ph) / (s * s)
        r = metric.r(l)
        dpl = b2 * metric.dr(l) / ((r * r) * r)
        dpth = (p_ph * p_ph) * (math.cos(th) / (metric.r2(l) * ((s * s) * s)))
        l, th, ph = l + dl * delta, th + dth * delta, ph + dph * delta
        p_l, p_th = p_l + dpl * delta, p_th + dpth * delta
        if l > R:
            return 1, k + 1, (l, th, ph, p_l, p_th, p_ph)
        if l < -R:
            return -1, k + 1, (l, th, ph, p_l, p_th, p_ph)
    return 0, max_iter, (l, th, ph, p_l, p_th, p_ph)


def texel(metric, state, W, H):
    l, th, ph, p_l, p_th, p_ph = state
    s = math.sin(th)
    d = [p_l, (p_th * (1.0 / metric.r2(l))) * metric.r(l), (p_ph * (1.0 / (metric.r2(l) * (s * s)))) * metric.r(l)]
    rn = math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    theta, phi = math.acos(d[2] / rn), math.atan2(d[1], d[0])
    phi = math.fmod(phi, 2.0 * PI)
    if phi < 0.0:
        phi = phi + 2.0 * PI
    phi = math.fmod(phi, 2.0 * PI)                       # second normalisation (images.rs:116); phi >= 0 here
    u = math.fmod(0.5 - phi / (2.0 * PI), 1.0)
    if u < 0.0:
        u = u + 1.0
    return int(u * W), int((theta / PI) * H)


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_python_restatement_equals_c_oracle_bitwise(oracle, kind):
    from curvis_b200 import scenes
    W, H, sim = 64, 36, (40000, 100.0, 0.05)
    metric = Ellis(1.0) if kind == "ellis" else Interstellar(0.1, 1e-4, 1.0)
    cam = oracle.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
    rot = np.array(cam.cam_to_world).reshape(3, 3).tolist()
    bp, bn = scenes.decodable_background(4096, 2048), scenes.decodable_background(4096, 2048, True)
    _, rec, _ = oracle.render_rows(oracle.metric(kind), cam, oracle.sim(*sim), bp, bn, threads=8)
    pixels = [(x, y) for y in range(0, H, 5) for x in range(0, W, 7)] + [(W // 2, H // 2), (0, 0), (W - 1, H - 1)]
    for px, py in pixels:
        d = pixel_ray(rot, cam.focal_length, cam.sensor_width, cam.sensor_height, W, H, px, py)
        side, steps, state = trace(metric, scenes.DEFAULT_CAMERA_POSITION, d, sim[2], sim[0], sim[1])
        r = rec[py, px]
        assert (side, steps) == (int(r["side"]), int(r["steps"])), (px, py)
        got = np.array(state, dtype=np.float64)
        want = np.array([r["l"], r["theta"], r["phi"], r["p_l"], r["p_theta"], r["p_phi"]], dtype=np.float64)
        assert got.tobytes() == want.tobytes(), (px, py, got, want)
        assert texel(metric, state, 4096, 2048) == (int(r["texel_x"]), int(r["texel_y"])), (px, py)


def test_negative_delta_and_odd_parameters(oracle):
    """update_relativistic_object accepts a negative delta ("evolving the object back in time",
    metrics.rs:279-280): the restatements agree there too."""
    from curvis_b200 import scenes
    cam = oracle.camera((0.0, -6.0, 1.1, 0.4), (1.0, 0.1, 0.2), (0.0, 0.0, 1.0), 20.0, 43.0, 24, 16)
    rot = np.array(cam.cam_to_world).reshape(3, 3).tolist()
    bp, bn = scenes.noise_background(64, 32, 1), scenes.noise_background(64, 32, 2)
    _, rec, _ = oracle.render_rows(oracle.metric("ellis", rho=2.0), cam, oracle.sim(3000, 40.0, -0.05), bp, bn)
    for px, py in [(0, 0), (5, 3), (12, 8), (23, 15)]:
        d = pixel_ray(rot, cam.focal_length, cam.sensor_width, cam.sensor_height, 24, 16, px, py)
        side, steps, state = trace(Ellis(2.0), (0.0, -6.0, 1.1, 0.4), d, -0.05, 3000, 40.0)
        r = rec[py, px]
        assert (side, steps) == (int(r["side"]), int(r["steps"]))
        assert np.array(state).tobytes() == np.array([r["l"], r["theta"], r["phi"], r["p_l"], r["p_theta"], r["p_phi"]]).tobytes()
