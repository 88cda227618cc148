"""fp32 sin/cos polynomial coefficients for curvis_b200/csrc/render_f32.cu (Remez in 50 digits,
rounded to float).  sin(r) = r + r*u*S(u) (deg 3 in u), cos(r) = 1 - u/2 + u*u*C(u) (deg 2 in u)."""
import mpmath as mp
import numpy as np
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_trig_coeffs import S, C, cheb_fit, remez_polish
mp.mp.dps = 50
R = mp.pi / 4 * mp.mpf("1.02"); U = R * R
sc = remez_polish(S, cheb_fit(S, 3, 0, U), 0, U)
cc = remez_polish(C, cheb_fit(C, 2, 0, U), 0, U)
for name, co in (("S", sc), ("C", cc)):
    print(name, ", ".join(repr(float(np.float32(float(c)))) + "f" for c in co))
errS = max(abs(S(U * k / 2000) - sum(mp.mpf(float(np.float32(float(c)))) * (U * k / 2000) ** j for j, c in enumerate(sc))) * (U * k / 2000) for k in range(2001))
errC = max(abs(C(U * k / 2000) - sum(mp.mpf(float(np.float32(float(c)))) * (U * k / 2000) ** j for j, c in enumerate(cc))) * (U * k / 2000) ** 2 for k in range(2001))
print("max err sin (rel to r):", mp.nstr(errS, 4), " cos (abs):", mp.nstr(errC, 4), " fp32 eps/2 = 5.96e-8")
p = mp.pi / 2
c1 = float(np.float32(float(p))); c2 = float(np.float32(float(p - mp.mpf(c1)))); c3 = float(np.float32(float(p - mp.mpf(c1) - mp.mpf(c2))))
print("pi/2 split f32:", repr(c1), repr(c2), repr(c3))
