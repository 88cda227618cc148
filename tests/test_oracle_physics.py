"""The oracle against closed-form physics (CPU).  The reference holds no test for escape_photon / the Euler step (SURVEY.md
8c: "parity unpinned"), and its Rust cannot be run here; what CAN be checked independently of any restatement is that the
equations the oracle integrates (src/metrics.rs:223-270, :417-421, :461-485) are the null geodesics of the metric
ds^2 = -dt^2 + dl^2 + r(l)^2 dOmega^2.  For an equatorial photon with impact parameter b (= p_phi / p_t)

    (dl/dlambda)^2 + b^2 / r(l)^2 = 1,      dphi/dlambda = b / r(l)^2,

so the azimuth swept between two values of l is the quadrature  int b dl / (r^2 sqrt(1 - b^2/r^2))  (for Ellis it reduces to
elliptic integrals: the total sweep of a ray that stays on one side is 2 K(rho/b)).  Forward Euler converges to it linearly
in the step; Richardson extrapolation of two step sizes removes the first-order term.  This pins signs, factors and the
shape functions r(l), r'(l) of BOTH metrics to the geometry — it does not pin the Euler scheme's own rounding, which only
the reference binary could."""
import math

import numpy as np
import pytest
from scipy import integrate, special

PI = math.pi
pytestmark = pytest.mark.filterwarnings("ignore::scipy.integrate.IntegrationWarning")   # inverse-square-root end point at the turning radius


def _r(kind, l, rho=1.0, m=0.1, a=1e-4):
    if kind == "ellis":
        return math.sqrt(rho * rho + l * l)
    al = abs(l)
    if al <= a:
        return rho
    x = 2.0 * (al - a) / (PI * m)
    return rho + m * (x * math.atan(x) - math.log1p(x * x) / 2.0)


def _sweep_exact(kind, b, l_from, l_to):
    """Azimuth swept while l runs monotonically from l_from to l_to (no turning point in between)."""
    f = lambda l: b / (_r(kind, l) ** 2 * math.sqrt(max(1.0 - b * b / _r(kind, l) ** 2, 1e-300)))
    val, _ = integrate.quad(f, l_from, l_to, limit=400, epsabs=1e-13, epsrel=1e-13)
    return abs(val)


def _euler_phi(oracle, kind, l0, alpha, delta, R):
    """phi at escape of the equatorial photon fired at angle alpha from the radial direction, and the exit l, side, steps."""
    g = oracle.metric(kind)
    x, p = oracle.new_photon(g, (0.0, l0, PI / 2.0, 0.0), (math.cos(alpha), 0.0, math.sin(alpha)))
    side, steps, xf, pf = oracle.escape_photon(g, x, p, delta, 4_000_000, R)
    return side, steps, xf, pf, p


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_ray_through_the_throat_sweeps_the_geodesic_azimuth(oracle, kind):
    """b < r_min = rho: the photon crosses to the negative side; no turning point, l decreases monotonically."""
    l0, R = 5.0, 60.0
    alpha = PI - 0.12                                   # inward, slightly off the radial direction
    b = _r(kind, l0) * math.sin(alpha)                   # p_phi / p_t = d_phi r(l0) sin(pi/2)
    assert b < 1.0
    phis = []
    for delta in (0.004, 0.002):
        side, steps, xf, pf, p0 = _euler_phi(oracle, kind, l0, alpha, delta, R)
        assert side == -1 and p0[3] == pytest.approx(b, rel=1e-15)
        # exact sweep up to the l the Euler ray actually stopped at
        phis.append((xf[3], xf[1]))
    exact = [_sweep_exact(kind, b, l_end, l0) for _, l_end in phis]
    err = [abs(phi) - ex for (phi, _), ex in zip(phis, exact)]
    assert abs(err[0]) < 2e-2 and abs(err[1]) < abs(err[0]) * 0.6        # first-order convergence
    richardson = 2 * err[1] - err[0]
    assert abs(richardson) < 2e-4, (err, richardson)


@pytest.mark.parametrize("kind", ["ellis", "interstellar"])
def test_ray_that_turns_back_sweeps_the_geodesic_azimuth(oracle, kind):
    """b > rho: the photon reaches its turning point r(l_t) = b and leaves on the side it came from."""
    l0, R = 5.0, 60.0
    alpha = PI - 0.45
    b = _r(kind, l0) * math.sin(alpha)
    assert b > 1.2
    # turning point: r(l_t) = b
    from scipy.optimize import brentq
    l_t = brentq(lambda l: _r(kind, l) - b, 1e-3, l0)
    errs = []
    for delta in (0.004, 0.002):
        side, steps, xf, pf, _ = _euler_phi(oracle, kind, l0, alpha, delta, R)
        assert side == 1
        exact = _sweep_exact(kind, b, l_t, l0) + _sweep_exact(kind, b, l_t, xf[1])
        errs.append(abs(xf[3]) - exact)
    assert abs(errs[0]) < 3e-2 and abs(errs[1]) < abs(errs[0]) * 0.6
    assert abs(2 * errs[1] - errs[0]) < 5e-4, errs


def test_ellis_total_deflection_is_the_complete_elliptic_integral():
    """Sanity of the quadrature itself against the textbook result: a ray from infinity to infinity on one side of the
    Ellis wormhole sweeps 2 K(k), k = rho / b (deflection 2 K(k) - pi)."""
    rho, b = 1.0, 1.7
    l_t = math.sqrt(b * b - rho * rho)
    total = 2 * (_sweep_exact("ellis", b, l_t, 2000.0) + b / 2000.0)     # tail beyond l = 2000: int b dl / l^2
    assert total == pytest.approx(2 * special.ellipk((rho / b) ** 2), rel=1e-6)


def test_null_condition_and_conserved_quantities_along_an_oracle_trajectory(oracle):
    """p_t and p_phi never change (metrics.rs:259-264); H = -p_t^2 + p_l^2 + (p_theta^2 + p_phi^2/sin^2)/r^2 starts at 0 and
    its Euler drift shrinks linearly with the step."""
    g = oracle.metric("ellis")
    d = np.array([-0.9, 0.3, 0.2])
    drift = []
    for delta in (0.01, 0.005):
        tr = oracle.trajectory(g, (0.0, 5.0, 1.1, 0.4), d, delta, int(8.0 / delta))
        l, th, pl, pth, pph, pt = tr[:, 1], tr[:, 2], tr[:, 5], tr[:, 6], tr[:, 7], tr[:, 4]
        assert (pph == pph[0]).all() and (pt == 1.0).all()
        r2 = 1.0 + l * l
        H = -1.0 + pl ** 2 + (pth ** 2 + pph ** 2 / np.sin(th) ** 2) / r2
        drift.append(np.abs(H).max())
    assert drift[0] < 0.05 and drift[1] < 0.6 * drift[0]
