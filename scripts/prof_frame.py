"""Renders N frames of a bench workload (default config 3) — the short command wrapped by ncu.
usage: python scripts/prof_frame.py [frames] [config]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pbrt_rust_b200 as pb
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
config = sys.argv[2] if len(sys.argv) > 2 else "c3"
cfg = bench.make_cfg(config)
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8)
for i in range(n):
    film = r.render(cfg["scene"])
print(r.last_stats)
