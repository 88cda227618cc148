"""Coefficients of the small-angle rotation used by CURVIS_PRECISION_F64_FAST
(curvis_b200/csrc/fast_f64.cuh: rotate_sincos).

For |x| < 2^-4, v = x*x:
    sin(x)     = x + x*v*S(v)
    cos(x) - 1 =     v*C(v)
    tan(x/2)   = x*(1/2 + v*T(v))      (the three-shear form of the rotation uses sin and tan-half)
S (degree 2) and C (degree 3) are polynomials in v (Remez, 60-digit arithmetic, rounded to double).  Prints
the coefficients and the worst approximation error (relative to x for sin, absolute for cos).
"""
import mpmath as mp
from gen_trig_coeffs import cheb_fit, remez_polish, show

mp.mp.dps = 60
X = mp.mpf(2) ** -4
V = X * X
DEG_S, DEG_C = 2, 3

def S(v):
    if v == 0: return -mp.mpf(1) / 6
    x = mp.sqrt(v); return (mp.sin(x) / x - 1) / v

def T(v):
    """tan(x/2) = x * (1/2 + v*T(v))"""
    if v == 0: return mp.mpf(1) / 24
    x = mp.sqrt(v); return (mp.tan(x / 2) / x - mp.mpf(1) / 2) / v

def C(v):
    if v == 0: return -mp.mpf(1) / 2
    x = mp.sqrt(v); return (mp.cos(x) - 1) / v

if __name__ == "__main__":
    sc = remez_polish(S, cheb_fit(S, DEG_S, 0, V), 0, V)
    cc = remez_polish(C, cheb_fit(C, DEG_C, 0, V), 0, V)
    show("S(v): sin(x) = x + x*v*S(v), ascending powers of v, |x| < 2^-4", sc)
    show("C(v): cos(x) = 1 + v*C(v), ascending powers of v, |x| < 2^-4", cc)
    grid = [V * k / 4000 for k in range(4001)]
    errS = max(abs(S(v) - sum(mp.mpf(float(c)) * v ** j for j, c in enumerate(sc))) * v for v in grid)
    errC = max(abs(C(v) - sum(mp.mpf(float(c)) * v ** j for j, c in enumerate(cc))) * v for v in grid)
    tc = remez_polish(T, cheb_fit(T, 2, 0, V), 0, V)
    show("T(v): tan(x/2) = x*(1/2 + v*T(v)), ascending powers of v, |x| < 2^-4", tc)
    errT = max(abs(T(v) - sum(mp.mpf(float(c)) * v ** j for j, c in enumerate(tc))) * v for v in grid)
    print("// max approximation error relative to x (tan half):", mp.nstr(errT, 5))
    print("// max approximation error relative to x (sin):", mp.nstr(errS, 5), " absolute (cos):", mp.nstr(errC, 5))
