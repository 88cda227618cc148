"""The guard band's budget re-measured for the Interstellar kernel only (tools/guard_study.py's scenes with that metric plus
two more cameras): largest deviation of the raw regrouped kernel from the operation-for-operation kernel among the rays with
stiffness < 1, and the frames with the guard on.   python tools/guard_study_interstellar.py [out.json]   (GPU box)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import guard_study as gs
import curvis_b200 as cv
from curvis_b200 import scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
dflt = (scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP)
cases = [
    ("interstellar_defaults_1080p", cv.InterstellarMetric(0.1, 1e-4, 1.0), dflt, (40000, 100.0, 0.05), 1920, 1080, None),
    ("interstellar_c3_4k_rows", cv.InterstellarMetric(0.1, 1e-4, 1.0), dflt, (2000, 45.0, 0.05), 3840, 2160, (900, 1260)),
    ("interstellar_tilted_720p", cv.InterstellarMetric(0.3, 0.5, 1.5), ((0.0, -7.0, 1.0, 2.5), (0.8, 0.5, -0.3), (0.1, 0.1, 1.0)), (40000, 150.0, 0.05), 1280, 720, None),
    ("interstellar_small_m_720p", cv.InterstellarMetric(0.02, 1e-3, 0.8), ((0.0, 3.0, 1.8, 0.4), (-0.9, 0.2, 0.3), (0.0, 0.3, 1.0)), (40000, 60.0, 0.03), 1280, 720, None),
]
out = []
for name, metric, cam, sim, W, H, rows in cases:
    r = gs.study(name, metric, cam, sim, W, H, ctx, bp, bn, rows)
    low = [b for b in r["by_stiffness"] if b["kappa"][1] is not None and b["kappa"][1] <= 1.0 and "eps_dir_max" in b]
    summary = {"scene": name, "rays": r["rays"], "stiffness_below_1": {"rays": sum(b["rays"] for b in low),
               "eps_dir_max": max(b["eps_dir_max"] for b in low), "dev_l_max": max(b["dev_l_max"] for b in low),
               "steps_differ": sum(b["steps_differ"] for b in low), "texel_differ": sum(b["texel_differ"] for b in low)},
               "raw": r["raw"], "guarded": r["guarded"]}
    out.append(summary)
    print(json.dumps(summary), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
