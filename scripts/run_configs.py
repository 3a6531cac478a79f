"""BASELINE.json's configs at their named sizes on the GPU(s) of this box: throughput through the C ABI
(one context, or pbrtb200_group_render over N GPUs in this one process) and parity against the CPU
oracle — primary-hit ids over the WHOLE frame at 1 sample per pixel, and the image on a cropped film of
the full-size scene (a 4K x 64 spp / 1080p x 256 spp CPU frame takes minutes to hours).

usage: python scripts/run_configs.py c2|c3|c4|c5 [--gpus N] [--frames K] [--no-oracle] [--small]
writes gpurun_out/config_<c>_n<N>.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pbrt_rust_b200 as pb
from pbrt_rust_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--no-oracle", action="store_true")
ap.add_argument("--small", action="store_true", help="reduced geometry (debug)")
args = ap.parse_args()
N = args.gpus


def mk(crop=(0, 1, 0, 1), spp=None):
    kw = {} if spp is None else dict(xs=spp[0], ys=spp[1])
    if args.config == "c2":
        return scenes.config2(crop=crop)
    if args.config == "c3":
        return scenes.config3(crop=crop, **kw)
    if args.config == "c4":
        return scenes.config4(n_ground=(100, 50), n_spheres=500, crop=crop, **kw) if args.small else scenes.config4(crop=crop, **kw)
    if args.config == "c5":
        return scenes.config5(nx=500, nz=500, crop=crop, **kw) if args.small else scenes.config5(crop=crop, **kw)
    raise SystemExit("unknown config")


t0 = time.perf_counter()
cfg = mk()
t_gen = time.perf_counter() - t0
ctx = pb.Group(list(range(N))) if N > 1 else pb.Context(0)
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, ctx=ctx)
t0 = time.perf_counter()
r.preprocess(cfg["scene"])
t_build = time.perf_counter() - t0
f = r.host_scene.flat.contents
info = dict(config=args.config, n_gpus=N, n_prims=int(f.n_prims), n_nodes=int(f.n_nodes), n_tris=int(f.n_tris),
            n_spheres=int(f.n_spheres), scene_gen_s=t_gen, bvh_build_flatten_upload_s=t_build)
film = cfg["film"]
h, w = film.shape
d_film = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda:0")
h_film = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory().numpy().reshape(h, w, 4)
times, times_e2e, stats = [], [], None
torch.cuda.reset_peak_memory_stats()
free0 = [torch.cuda.mem_get_info(i)[0] for i in range(N)]
if args.config == "c2":
    for i in range(args.frames + 1):
        t0 = time.perf_counter()
        hits, _, _ = r.primary_hits(cfg["scene"])
        times.append(time.perf_counter() - t0)
        stats = dict(r.last_stats)
    times_e2e = times
else:
    warm = 15 if N > 1 else 1
    for i in range(args.frames + warm):
        t0 = time.perf_counter()
        r.render(cfg["scene"], out=d_film)
        times.append(time.perf_counter() - t0)
        stats = dict(r.last_stats)
    times = times[warm:]
    for i in range(args.frames + 1):
        t0 = time.perf_counter()
        r.render(cfg["scene"], out=h_film)
        times_e2e.append(time.perf_counter() - t0)
    times_e2e = times_e2e[1:]
ms = 1e3 * float(np.median(times))
ms_e2e = 1e3 * float(np.median(times_e2e))
e = cfg["sampler"].ext
frame_cam = (e[1] - e[0]) * (e[3] - e[2]) * cfg["sampler"].samples_per_pixel()
rays = (stats["camera_rays"] + stats["shadow_rays"]) * frame_cam / max(1, stats["camera_rays"])
info.update(ms_per_frame=ms, ms_per_frame_e2e=ms_e2e, rays_per_frame=rays, mrays_per_s=rays / ms / 1e3,
            mrays_per_s_e2e=rays / ms_e2e / 1e3, device_stage_ms={k: v for k, v in stats.items() if k.startswith("ms_")},
            kernel_launches=stats["kernel_launches"],
            device_memory_used_gb=[round((free0[i] - torch.cuda.mem_get_info(i)[0]) / 2**30 + 0.0, 2) for i in range(N)],
            device_memory_in_use_total_gb=[round((torch.cuda.mem_get_info(i)[1] - torch.cuda.mem_get_info(i)[0]) / 2**30, 2) for i in range(N)])
if N > 1:
    info["band_rows"], info["per_device_ms"] = ctx.bands()
if not args.no_oracle:
    from oracle import orc
    t0 = time.perf_counter()
    osc = orc.OracleScene(cfg["scene"])
    info["oracle_bvh_build_s"] = time.perf_counter() - t0
    nt = os.cpu_count() or 8
    if args.config == "c2":
        ref = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0, primary_only=True, n_threads=nt), want_hits=True)
        info["hit_id_agreement"] = float(np.mean(hits["prim"] == ref["hit_ids"]))
        info["hit_id_rays"] = int(hits.shape[0])
    else:
        # (1) primary-hit ids over the whole frame, one sample per pixel, same uploaded scene
        c1 = mk(spp=(1, 1))
        r1 = pb.GpuRenderer(c1["sampler"], c1["camera"], c1["integrator"], num_cpus=8, ctx=pb.Context(0) if N > 1 else ctx)
        if N > 1:
            r1.ctx.upload(r.host_scene, scene_key=c1["scene"])
        else:
            ctx.scene_key = c1["scene"]   # same geometry, already on the device
        hits1, _, _ = r1.primary_hits(c1["scene"])
        t0 = time.perf_counter()
        ref1 = orc.render(osc, orc.render_config(c1["camera"], c1["sampler"], num_cpus=8, mode=0, primary_only=True, n_threads=nt), want_hits=True)
        info["oracle_primary_s"] = time.perf_counter() - t0
        same = hits1["prim"] == ref1["hit_ids"]
        bad = np.flatnonzero(~same)
        both = bad[(hits1["prim"][bad] != 0xFFFFFFFF) & (ref1["hit_ids"][bad] != 0xFFFFFFFF)]
        info.update(hit_id_rays=int(hits1.shape[0]), hit_id_agreement=float(same.mean()), hit_id_mismatches=int(bad.size),
                    mismatch_max_rel_dt=float(np.max(np.abs(hits1["t"][both] - ref1["hit_ts"][both]) / np.maximum(ref1["hit_ts"][both], 1e-9))) if both.size else 0.0,
                    hit_fraction=float((ref1["hit_ids"] != 0xFFFFFFFF).mean()))
        # (2) the image on a cropped film of the full-size scene (the reference's own crop-window semantics)
        xr, yr = film.x_res, film.y_res
        crop = (0.45, 0.45 + 64.0 / xr, 0.55, 0.55 + 48.0 / yr)
        ccfg = mk(crop)
        cr = pb.GpuRenderer(ccfg["sampler"], ccfg["camera"], ccfg["integrator"], num_cpus=8, ctx=r1.ctx)
        r1.ctx.scene_key = ccfg["scene"]
        cfilm = cr.render(ccfg["scene"])
        t0 = time.perf_counter()
        ref = orc.render(osc, orc.render_config(ccfg["camera"], ccfg["sampler"], num_cpus=8, mode=0, n_threads=nt))
        info["oracle_crop_s"] = time.perf_counter() - t0
        rgb, rgb_ref = pb.film_to_rgb(cfilm), ref["rgb"]
        rel = np.abs(rgb - rgb_ref) / np.maximum(np.abs(rgb_ref), 1e-3)
        info.update(crop_pixels=list(cfilm.shape[:2]), crop_weights_equal=bool(np.array_equal(cfilm[..., 3], ref["film"][..., 3])),
                    image_rmse=float(np.sqrt(np.mean((rgb - rgb_ref) ** 2))), frac_pixels_rel_err_le_1e4=float((rel.max(-1) <= 1e-4).mean()),
                    max_abs_err=float(np.abs(rgb - rgb_ref).max()), mean_rgb=float(rgb_ref.mean()))
print(json.dumps(info))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(info, open(f"gpurun_out/config_{args.config}_n{N}.json", "w"), indent=1)
