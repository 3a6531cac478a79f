"""Single-GPU probe: cost of rendering 1/8 of the config-3 frame under different tile partitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pbrt_rust_b200 as pb
from pbrt_rust_b200 import multigpu
import bench

cfg = bench.make_cfg()
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8)
r.preprocess(cfg["scene"])
film = cfg["film"]; h, w = film.shape
d = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
ext = film.get_pixel_extent()

def run(name, tiles):
    for _ in range(3):
        r.render(cfg["scene"], tiles=tiles, out=d)
    acc = {}
    for _ in range(10):
        r.render(cfg["scene"], tiles=tiles, out=d)
        for k, v in r.last_stats.items():
            acc[k] = acc.get(k, 0) + v / 10
    print("%-28s cam rays %9d | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f total %.3f | trace Grays/s %.2f" % (
        name, acc["camera_rays"], acc["ms_raygen"], acc["ms_trace"], acc["ms_shade"], acc["ms_shadow"], acc["ms_film"],
        acc["ms_total"], acc["camera_rays"] / acc["ms_trace"] / 1e6))

import sys
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
def total(tiles):
    for _ in range(2):
        r.render(cfg["scene"], tiles=tiles, out=d)
    t = tr = 0.0
    for _ in range(6):
        r.render(cfg["scene"], tiles=tiles, out=d)
        t += r.last_stats["ms_total"] / 6
        tr += (r.last_stats["ms_trace"] + r.last_stats["ms_shadow"]) / 6
    return t, tr, r.last_stats["camera_rays"]
def report(name, parts):
    res = [total(p) for p in parts]
    print("%-34s per-rank ms_total %s | max %.3f mean %.3f" % (name, " ".join("%.2f" % x[0] for x in res), max(x[0] for x in res), sum(x[0] for x in res) / len(res)))
for t in (64,):
    report(f"cyclic {t}x{t}", [multigpu.partition_tiles(ext, k, W, tile=t) for k in range(W)])
def bands(nb):
    ys = [round(i * h / nb) for i in range(nb + 1)]
    return [(0, ys[i], w, ys[i + 1]) for i in range(nb)]
b = bands(W); report("row bands", [[b[k]] for k in range(W)])
b = bands(2 * W); report("mirrored band pairs", [[b[k], b[2 * W - 1 - k]] for k in range(W)])
b = bands(4 * W); report("cyclic bands x4", [[b[k + W * j] for j in range(4)] for k in range(W)])
b = bands(4 * W); report("mirrored cyclic bands x4", [[b[k], b[2 * W - 1 - k], b[2 * W + k], b[4 * W - 1 - k]] for k in range(W)])
gx, gy = (4, 2) if W == 8 else ((2, 2) if W == 4 else (W, 1))
report(f"blocks {gx}x{gy}", [[(round((k % gx) * w / gx), round((k // gx) * h / gy), round((k % gx + 1) * w / gx), round((k // gx + 1) * h / gy))] for k in range(W)])
for t in (256, 512):
    report(f"cyclic {t}x{t}", [multigpu.partition_tiles(ext, k, W, tile=t) for k in range(W)])
