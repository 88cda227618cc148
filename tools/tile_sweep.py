"""One rank's tile of a frame split over 8 ranks, timed on ONE GPU: kernel time of the F64_FAST launch against resident CTAs per SM
(fewer warps per scheduler = a shorter latency per step for the tile's longest rays, at some cost in throughput), the guard band
and the longest-first refill.  Tiles: rows g::8 (interleaved, g = 7 holds row 1079, the pole-grazing neighbour of the central row)
and rows 945..1215 (the central contiguous eighth).   python tools/tile_sweep.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import curvis_b200 as cv
from curvis_b200 import _abi, scenes

ctx = cv.Context([0])
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
stream = torch.cuda.current_stream()
out = {}
for W, H in ((3840, 2160), (7680, 4320)):
    frame = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda:0")
    for mname, metric in (("ellis", cv.EllisMetric(1.0)), ("interstellar", cv.InterstellarMetric(0.1, 1e-4, 1.0))):
        sim = (40000, 100.0, 0.05)
        cam = cv.Camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, W, H)
        system = cv.RelativisticSystem(metric, cv.SphericalImage(bp), cv.SphericalImage(bn), cam, context=ctx)

        def run(tile, reps=4, **opts):
            for k, v in opts.items():
                ctx.set_option(k, v)
            ms = []
            for _ in range(reps):
                if tile == "central":
                    st = system.render_rows_device(*sim, H * 7 // 16, H * 9 // 16, frame.data_ptr(), stream.cuda_stream, want_stats=True, precision=_abi.PRECISION_F64_FAST)
                else:
                    st = system.render_frames_peers([cam], *sim, tile, H, [frame.data_ptr()], stream.cuda_stream, want_stats=True, row_stride=8, precision=_abi.PRECISION_F64_FAST)
                ms.append(st["kernel_ms"])
            return round(min(ms[1:]), 3), int(st["total_steps"])
        res = {}
        for tile in (7, 0, "central"):
            for guard in (0, 1):
                for lf in (0, 1):
                    row = {}
                    for bps in (5, 4, 3, 2):
                        row[f"bps{bps}"], steps = run(tile, blocks_per_sm=bps, guard=guard, longest_first=lf)
                    row["total_steps"] = steps
                    res[f"tile{tile}_guard{guard}_lf{lf}"] = row
        ctx.set_option("blocks_per_sm", 0); ctx.set_option("guard", 1); ctx.set_option("longest_first", 2)
        out[f"{mname}_{W}x{H}"] = res
        print(mname, W, H, flush=True)
        for k, v in res.items():
            print("  ", k, json.dumps(v), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
