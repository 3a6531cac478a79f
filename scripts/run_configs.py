"""Runs BASELINE.json's other configs at their named sizes on the GPU(s) of this box and records
throughput + a parity check against the CPU oracle on a cropped film of the same scene.
usage: python scripts/run_configs.py c2|c4|c5 [--frames N] [--no-oracle]
(torchrun for several GPUs: tiles are partitioned like bench.py)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pbrt_rust_b200 as pb
from pbrt_rust_b200 import scenes, multigpu

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--no-oracle", action="store_true")
ap.add_argument("--small", action="store_true", help="reduced geometry (debug)")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
import torch
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def mk(crop=(0, 1, 0, 1)):
    if args.config == "c2":
        return scenes.config2(crop=crop)
    if args.config == "c4":
        return scenes.config4(n_ground=(100, 50), n_spheres=500, crop=crop) if args.small else scenes.config4(crop=crop)
    if args.config == "c5":
        return scenes.config5(nx=500, nz=500, crop=crop) if args.small else scenes.config5(crop=crop)
    raise SystemExit("unknown config")

t0 = time.perf_counter(); cfg = mk(); t_gen = time.perf_counter() - t0
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, device=local)
t0 = time.perf_counter(); r.preprocess(cfg["scene"]); t_build = time.perf_counter() - t0
f = r.host_scene.flat.contents
info = dict(config=args.config, n_gpus=world, n_prims=int(f.n_prims), n_nodes=int(f.n_nodes), n_tris=int(f.n_tris),
            n_spheres=int(f.n_spheres), scene_gen_s=t_gen, bvh_build_flatten_upload_s=t_build)
film = cfg["film"]; h, w = film.shape
d_film = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
tiles = None
if world > 1 and args.config != "c2":
    # time-balanced row bands (multigpu.BandBalancer), settled on a few untimed frames
    bal = multigpu.BandBalancer(film.get_pixel_extent(), world)
    tiles = bal.tiles_for(rank)
    for _ in range(8):
        r.render(cfg["scene"], tiles=tiles, out=d_film)
        t = torch.zeros(world, dtype=torch.float64, device="cuda"); t[rank] = r.last_stats["ms_total"]
        dist.all_reduce(t)
        tl = [float(x) for x in t.cpu()]
        if bal.imbalance(tl) < 1.02 or not bal.update(tl):
            break
        tiles = bal.tiles_for(rank)
    info["band_rows"] = bal.b
times, stats = [], None
if args.config == "c2":
    for i in range(args.frames + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hits, _, _ = r.primary_hits(cfg["scene"])
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0); stats = dict(r.last_stats)
else:
    for i in range(args.frames + 1):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r.render(cfg["scene"], tiles=tiles, out=d_film)
        if world > 1: dist.reduce(d_film, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0); stats = dict(r.last_stats)
ms = 1e3 * float(np.median(times[1:]))
rays = stats["camera_rays"] + stats["shadow_rays"]
if world > 1:
    t = torch.tensor([float(stats["camera_rays"]), float(stats["shadow_rays"]), ms], dtype=torch.float64, device="cuda")
    s = t.clone(); dist.all_reduce(s); m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MAX)
    e = cfg["sampler"].ext
    frame_cam = (e[1] - e[0]) * (e[3] - e[2]) * cfg["sampler"].samples_per_pixel()
    rays = (float(s[0]) + float(s[1])) * frame_cam / float(s[0]); ms = float(m[2])
info.update(ms_per_frame=ms, rays_per_frame=rays, mrays_per_s=rays / ms / 1e3, device_stage_ms={k: v for k, v in stats.items() if k.startswith("ms_")})
if rank == 0 and not args.no_oracle:
    from oracle import orc
    t0 = time.perf_counter(); osc = orc.OracleScene(cfg["scene"]); info["oracle_bvh_build_s"] = time.perf_counter() - t0
    if args.config == "c2":
        ref = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0, primary_only=True, n_threads=16), want_hits=True)
        info["hit_id_agreement"] = float(np.mean(hits["prim"] == ref["hit_ids"]))
    else:
        # parity on a cropped film of the same scene (the reference's own crop-window semantics)
        xr, yr = film.x_res, film.y_res
        crop = (0.45, 0.45 + 48.0 / xr, 0.55, 0.55 + 32.0 / yr)
        ccfg = mk(crop)
        cr = pb.GpuRenderer(ccfg["sampler"], ccfg["camera"], ccfg["integrator"], num_cpus=8, device=local, ctx=r.ctx)
        cr.host_scene, cr._scene_key = r.host_scene, ccfg["scene"]      # same uploaded scene
        cfilm = cr.render(ccfg["scene"])
        chits, _, _ = cr.primary_hits(ccfg["scene"])
        t0 = time.perf_counter()
        ref = orc.render(osc, orc.render_config(ccfg["camera"], ccfg["sampler"], num_cpus=8, mode=0, n_threads=16), want_hits=True)
        info["oracle_crop_s"] = time.perf_counter() - t0
        rgb, rgb_ref = pb.film_to_rgb(cfilm), ref["rgb"]
        rel = np.abs(rgb - rgb_ref) / np.maximum(np.abs(rgb_ref), 1e-3)
        info.update(crop_pixels=list(cfilm.shape[:2]), hit_id_agreement=float(np.mean(chits["prim"] == ref["hit_ids"])),
                    image_rmse=float(np.sqrt(np.mean((rgb - rgb_ref) ** 2))), frac_pixels_rel_err_le_1e4=float((rel.max(-1) <= 1e-4).mean()),
                    max_abs_err=float(np.abs(rgb - rgb_ref).max()), mean_rgb=float(rgb_ref.mean()))
if rank == 0:
    print(json.dumps(info))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(info, open(f"gpurun_out/config_{args.config}_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
