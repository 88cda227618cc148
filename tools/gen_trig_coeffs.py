"""Generates the polynomial coefficients of curvis_b200/csrc/trig_f64.cuh.

sin(r) = r + r*u*S(u),  cos(r) = 1 - u/2 + u*u*C(u),  u = r*r, |r| <= pi/4 (+ slack).
S and C are degree-5 polynomials in u obtained by Chebyshev-node interpolation (near-minimax)
in 60-digit arithmetic, then rounded to double.  Also prints the 3-term Cody-Waite split of pi/2.
"""
import mpmath as mp
mp.mp.dps = 60

R = mp.pi / 4 * mp.mpf("1.02")       # a little slack past pi/4 (rounding of the quadrant index)
U = R * R

def S(u):
    if u == 0: return -mp.mpf(1) / 6
    r = mp.sqrt(u); return (mp.sin(r) / r - 1) / u

def C(u):
    if u == 0: return mp.mpf(1) / 24
    r = mp.sqrt(u); return (mp.cos(r) - 1 + u / 2) / (u * u)

def cheb_fit(f, deg, a, b):
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n); y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n): A[i, j] = x ** j
        y[i] = f(x)
    return list(mp.lu_solve(A, y))

def remez_polish(f, coef, a, b, iters=6):
    """A few Remez exchange steps starting from the Chebyshev interpolant (equioscillation)."""
    n = len(coef) + 1
    pts = [(a + b) / 2 - (b - a) / 2 * mp.cos(mp.pi * k / (n - 1)) for k in range(n)]
    for _ in range(iters):
        A = mp.matrix(n, n); y = mp.matrix(n, 1)
        for i, x in enumerate(pts):
            for j in range(n - 1): A[i, j] = x ** j
            A[i, n - 1] = (-1) ** i
            y[i] = f(x)
        sol = mp.lu_solve(A, y); coef = list(sol)[: n - 1]
        err = lambda x: f(x) - sum(c * x ** j for j, c in enumerate(coef))
        # new extrema: dense scan between sign changes
        grid = [a + (b - a) * k / 4000 for k in range(4001)]
        vals = [err(x) for x in grid]
        ext = []
        for i in range(1, 4000):
            if (vals[i] - vals[i - 1]) * (vals[i + 1] - vals[i]) <= 0: ext.append(grid[i])
        cand = [a] + ext + [b]
        if len(cand) < n: break
        # keep the n largest alternating
        cand.sort()
        while len(cand) > n:
            mags = [abs(err(x)) for x in cand]
            cand.pop(mags.index(min(mags)))
        pts = cand
    return coef

def show(name, coef):
    print(f"// {name}")
    for c in coef:
        d = float(c)
        print(f"    {d!r},   // {d.hex()}")

if __name__ == "__main__":
    sc = remez_polish(S, cheb_fit(S, 5, 0, U), 0, U)
    cc = remez_polish(C, cheb_fit(C, 5, 0, U), 0, U)
    show("S(u): sin(r) = r + r*u*S(u), ascending powers of u", sc)
    show("C(u): cos(r) = 1 - u/2 + u*u*C(u), ascending powers of u", cc)
    errS = max(abs(S(U * k / 2000) - sum(mp.mpf(float(c)) * (U * k / 2000) ** j for j, c in enumerate(sc))) * (U * k / 2000) for k in range(2001))
    errC = max(abs(C(U * k / 2000) - sum(mp.mpf(float(c)) * (U * k / 2000) ** j for j, c in enumerate(cc))) * (U * k / 2000) ** 2 for k in range(2001))
    print("// max approximation error relative to r (sin):", mp.nstr(errS, 5), " absolute (cos):", mp.nstr(errC, 5))
    p = mp.pi / 2
    c1 = float(p); c2 = float(p - mp.mpf(c1)); c3 = float(p - mp.mpf(c1) - mp.mpf(c2))
    print("// pi/2 split:", repr(c1), repr(c2), repr(c3))
    print("// 2/pi:", repr(float(2 / mp.pi)))
