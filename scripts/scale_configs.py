"""One process, one box: BASELINE configs at named size on 1 / 2 / 4 / 8 GPUs through pbrtb200_group_render
(the scene is built once on the host and uploaded to every device of each group).
usage: python scripts/scale_configs.py c5 [c4 ...] [--frames K] [--max-gpus N]
writes gpurun_out/scale_<config>.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pbrt_rust_b200 as pb
import bench

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="+")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--max-gpus", type=int, default=8)
args = ap.parse_args()
navail = torch.cuda.device_count()
os.makedirs("gpurun_out", exist_ok=True)
for config in args.configs:
    t0 = time.perf_counter()
    cfg = bench.make_cfg(config)
    host = pb.HostScene(cfg["scene"])
    t_host = time.perf_counter() - t0
    film = cfg["film"]
    h, w = film.shape
    e = cfg["sampler"].ext
    frame_cam = (e[1] - e[0]) * (e[3] - e[2]) * cfg["sampler"].samples_per_pixel()
    rows, ref_film = [], None
    for n in (1, 2, 4, 8):
        if n > min(navail, args.max_gpus):
            break
        ctx = pb.Group(list(range(n))) if n > 1 else pb.Context(0)
        t0 = time.perf_counter()
        ctx.upload(host, scene_key=cfg["scene"])
        t_up = time.perf_counter() - t0
        r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, ctx=ctx)
        d_film = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda:0")
        h_film = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory().numpy().reshape(h, w, 4)
        warm = 15 if n > 1 else 1
        ts = []
        for i in range(warm + args.frames):
            t0 = time.perf_counter()
            r.render(cfg["scene"], out=d_film)
            ts.append(time.perf_counter() - t0)
        st = dict(r.last_stats)
        ms = 1e3 * float(np.median(ts[warm:]))
        te = []
        for i in range(1 + args.frames):
            t0 = time.perf_counter()
            r.render(cfg["scene"], out=h_film)
            te.append(time.perf_counter() - t0)
        ms_e2e = 1e3 * float(np.median(te[1:]))
        if ref_film is None:
            ref_film = h_film.copy()
        same = bool(np.array_equal(h_film.view(np.uint32), ref_film.view(np.uint32)))
        rays = (st["camera_rays"] + st["shadow_rays"]) * frame_cam / max(1, st["camera_rays"])
        row = dict(n_gpus=n, ms_per_frame=ms, ms_per_frame_e2e=ms_e2e, mrays_per_s=rays / ms / 1e3, mrays_per_s_e2e=rays / ms_e2e / 1e3,
                   first_frame_ms=1e3 * ts[0], upload_s=t_up, film_bit_identical_to_1_gpu=same,
                   bands=ctx.bands() if n > 1 else None,
                   device_memory_in_use_gb=[round((torch.cuda.mem_get_info(i)[1] - torch.cuda.mem_get_info(i)[0]) / 2**30, 2) for i in range(n)])
        rows.append(row)
        print(config, json.dumps(row), flush=True)
        del r, ctx, d_film
        torch.cuda.empty_cache()
    f = host.flat.contents
    out = dict(config=config, workload=bench.WORKLOADS[config], n_prims=int(f.n_prims), host_scene_build_s=t_host, rays_per_frame=rays, rows=rows)
    json.dump(out, open(f"gpurun_out/scale_{config}.json", "w"), indent=1)
