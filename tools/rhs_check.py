"""kernel_variant 4 (shared reciprocals) against the plain IEEE operators on the GPU: python tools/rhs_check.py [log2 samples]."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvis_b200 as cv

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 31)
ctx = cv.Context([0])
out = {}
for name, metric in (("ellis", cv.EllisMetric(1.0)), ("ellis_rho2", cv.EllisMetric(2.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0)),
                     ("flat", cv.FlatSphericalMetric())):
    t0 = time.time()
    bad = ctx.debug_rhs_check(metric, n, seed=20261017)
    out[name] = {"samples": n, "mismatching_outputs": dict(zip(("dtheta", "dphi", "dp_l", "dp_theta"), bad)), "seconds": round(time.time() - t0, 2)}
print(json.dumps(out))
