"""One-shot latency: scene preprocessing, first frame (allocations, pixel list, lazy kernel loading) and
steady-state frames of a bench config.  usage: python scripts/first_frame_probe.py [config ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pbrt_rust_b200 as pb
import bench

for config in (sys.argv[1:] or ["c3"]):
    cfg = bench.make_cfg(config)
    t0 = time.perf_counter()
    r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8)
    r.preprocess(cfg["scene"])
    t_pre = time.perf_counter() - t0
    h, w = cfg["film"].shape
    host = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory().numpy().reshape(h, w, 4)
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        r.render(cfg["scene"], out=host)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(config, "preprocess %.1f ms | frames (host film) %s ms | device ms_total of the last %.2f" %
          (t_pre * 1e3, " ".join("%.1f" % t for t in ts), r.last_stats["ms_total"]), flush=True)
