"""Diagnostic: per-device stats of pbrtb200_group_render over a few frames (GPU box).
usage: python scripts/group_probe.py [n_gpus] [config] [frames]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pbrt_rust_b200 as pb
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
config = sys.argv[2] if len(sys.argv) > 2 else "c3"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = bench.make_cfg(config)
grp = pb.Group(list(range(n)))
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, ctx=grp)
r.preprocess(cfg["scene"])
h, w = cfg["film"].shape
host = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory().numpy().reshape(h, w, 4)
dev = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda:0")
summary = {}
for mode, out in (("host", host), ("device", dev)):
    walls = []
    for f in range(frames):
        t0 = time.perf_counter(); r.render(cfg["scene"], out=out); dt = (time.perf_counter() - t0) * 1e3
        walls.append(dt)
        b, ms = grp.bands()
        ds = grp.device_stats()
        print(mode, f, "wall %.2f ms" % dt, "bands", b, " | ".join("dev%d rays %d total %.2f trace %.2f shadow %.2f shade %.2f film %.2f" % (i, d["camera_rays"], d["ms_total"], d["ms_trace"], d["ms_shadow"], d["ms_shade"], d["ms_film"]) for i, d in enumerate(ds)))
    summary[mode] = float(np.median(walls[-max(3, frames // 3):]))
print("SUMMARY n=%d %s host_film_stores=%s" % (n, config, os.environ.get("PBRTB200_HOST_FILM_STORES", "default")),
      " ".join("%s %.3f ms" % kv for kv in summary.items()), "bands", grp.bands()[0])
