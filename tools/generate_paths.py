"""Writes curvis_b200/paths/path_through.csv: a 1000-way-point camera path that crosses the
wormhole from l = -4 to l = +4 in 20 time units on the equator, looking along the direction of
travel with an impact-parameter bump b(l) = 3 exp(-10 (l/4)^2) near the throat — the same path
family as the reference's paths/generate_path_through.py (its sample input for `curvis video`).
Columns: t, l, theta, phi, fx, fy, fz, upx, upy, upz."""
import os
import numpy as np

l0, l1, T, b0, n = -4.0, 4.0, 20.0, 3.0, 1000
ls, ts = np.linspace(l0, l1, n), np.linspace(0.0, T, n)
alpha = np.pi - np.arctan(b0 * np.exp(-10.0 * (ls / l0) ** 2) / ls)
fx, fy = np.sign(ls) * np.cos(alpha), np.sign(ls) * np.sin(alpha)
rows = ["t,l,theta,phi,fx,fy,fz,upx,upy,upz"]
for i in range(n):
    rows.append(",".join(str(float(v)) for v in (ts[i], ls[i], np.pi / 2, 0.0, fx[i], fy[i], 0.0, 0.0, 0.0, 1.0)))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curvis_b200", "paths", "path_through.csv")
open(out, "w").write("\n".join(rows))
print(out, len(rows) - 1, "way-points")
