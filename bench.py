#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 rendering back end (contract in the task prompt).

A "step" is one full frame of the hot path: BASELINE.json config 3 — a procedural 1 M-triangle
heightfield + quad area light, Whitted/direct lighting, 1920x1080, Stratified 4x4 = 16 spp,
box filter — i.e. 33.2 M camera rays + their shadow rays per frame.

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

metric  : Mrays/s (primary + shadow), whole job.  `value` = scene resident and film left in HBM;
          `e2e` = same call through the C ABI with a pinned HOST film buffer (D2H inside the timed
          region, descriptors H2D).  ms_per_step is the frame time.
roofline: dominant kernel = k_trace (closest hit).  achieved = algorithmic bytes / device time of
          its launches in the timed region; algorithmic bytes/ray = 32 N_nodes + 48 N_tri +
          80 N_sph + 48 with N_* counted by the CPU oracle on a bounded sample
          (same scene and camera at 1/16 of the pixels) and scaled by the ray count.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "config3: 1M-tri heightfield + quad area light, direct lighting, 1920x1080, 16 spp"


def make_cfg(scale=1):
    from pbrt_rust_b200 import scenes
    return scenes.config3(nx=1000, nz=500, xres=1920 // scale, yres=1080 // scale, xs=4, ys=4)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_sample(cfg_small, n_threads, mode, count):
    """Oracle run on the bounded sample; returns (stats dict, seconds)."""
    from oracle import orc
    osc = orc.OracleScene(cfg_small["scene"])
    oc = orc.render_config(cfg_small["camera"], cfg_small["sampler"], num_cpus=n_threads, mode=mode,
                           n_threads=n_threads, count_traversal=count)
    t0 = time.perf_counter()
    res = orc.render(osc, oc)
    return res["stats"], time.perf_counter() - t0, osc, oc


def bytes_per_ray(st):
    """B_ray = 32 N_nodes + 48 N_tri + 80 N_sph + 48, averaged over the sample."""
    prim = (32 * st["nodes_visited"] + 48 * st["tris_tested"] + 80 * st["spheres_tested"]) / max(1, st["camera_rays"]) + 48
    sh = (32 * st["sh_nodes_visited"] + 48 * st["sh_tris_tested"] + 80 * st["sh_spheres_tested"]) / max(1, st["shadow_rays"]) + 48
    return prim, sh


def run_reference(args):
    """CPU arm: the oracle in reference-faithful (strict) mode — the reference's own task split,
    per-task StdRng and sub-films, one thread per host core — on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build()
    from oracle import orc
    cores = os.cpu_count() or 1
    cfg_s = make_cfg(scale=2)
    osc = orc.OracleScene(cfg_s["scene"])
    oc = orc.render_config(cfg_s["camera"], cfg_s["sampler"], num_cpus=cores, mode=1, n_threads=cores)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = orc.render(osc, oc)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            # shadow rays are not counted in the plain timing run; count once below
    oc_c = orc.render_config(cfg_s["camera"], cfg_s["sampler"], num_cpus=cores, mode=1, n_threads=cores,
                             count_traversal=True)
    st = orc.render(osc, oc_c)["stats"]
    rays = st["camera_rays"] + st["shadow_rays"]
    ms = 1e3 * float(np.mean(times))
    v = rays / (ms * 1e-3) / 1e6
    sample = "same scene+camera at 960x540x16spp (1/4 of the frame's pixels) per step, strict task mode"
    line = {
        "impl": "reference", "metric": "Mrays/s (primary+shadow)", "value": v, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of pbrt_rust's algorithm (oracle); the Rust crate cannot be built here",
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / oracle-count leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0:
        g.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    torch.cuda.set_device(local)
    import pbrt_rust_b200 as pb

    cfg = make_cfg()
    film = cfg["film"]
    r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, device=local)
    stream = torch.cuda.current_stream()
    r.ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    r.preprocess(cfg["scene"])  # host BVH build + flatten + upload (once, like scene creation)
    t_scene = time.perf_counter() - t0
    h, w = film.shape
    from pbrt_rust_b200 import multigpu
    # N > 1 partition.  "bands" (default): contiguous row bands whose boundaries follow the ranks'
    # measured frame times during the warm-up frames (multigpu.BandBalancer) and are frozen for the
    # timed frames; PBRTB200_PARTITION=cyclic: 64x64 tiles, tile -> rank = id mod N.
    partition = os.environ.get("PBRTB200_PARTITION", "bands") if world > 1 else "whole"
    balancer = multigpu.BandBalancer(film.get_pixel_extent(), world) if partition == "bands" else None
    tiles = None
    if world > 1:
        tiles = balancer.tiles_for(rank) if balancer else multigpu.partition_tiles(film.get_pixel_extent(), rank, world)
    d_film = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
    h_film_t = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory()
    h_film = h_film_t.numpy().reshape(h, w, 4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1 film gather.  Default "p2p": rank 0 shares its film buffer over CUDA IPC and every
    # rank's film kernel stores its owned tiles straight into it over NVLink (two buffers,
    # alternating, so rank 0 can still read frame k while frame k+1 is written); the only other
    # per-frame communication is a barrier.  PBRTB200_GATHER=nccl: reduce(SUM) of zero-padded films.
    gather = os.environ.get("PBRTB200_GATHER", "p2p") if world > 1 else "none"
    peers, frame_no = [], [0]
    if gather == "p2p":
        peers = [multigpu.PeerFilm(r.ctx, h * w, dist, torch.device("cuda", local)) for _ in range(2)]
        if not all(p.ok for p in peers):  # no CUDA IPC / peer access between these GPUs (same on every rank)
            if rank == 0:
                print("bench: CUDA IPC film sharing unavailable; using the NCCL gather", file=sys.stderr)
            gather, peers = "nccl", []
    peer_views = [p.tensor() for p in peers] if (gather == "p2p" and rank == 0) else []
    fence = torch.zeros(1, dtype=torch.int32, device="cuda")

    def frame(resident):
        if world == 1:
            # resident: film left in HBM; e2e: C ABI with a host film buffer (D2H inside the call)
            r.render(cfg["scene"], tiles=tiles, out=d_film if resident else h_film)
            return r.last_stats
        k = frame_no[0] & 1
        frame_no[0] += 1
        if gather == "p2p":
            r.render(cfg["scene"], tiles=tiles, out=peers[k].ptr, keep_others=True)
            # frame fence: a 4-byte all-reduce ordered on the stream after this rank's film kernel;
            # when it completes, every rank's stores have landed in rank 0's HBM (no host block)
            dist.all_reduce(fence, op=dist.ReduceOp.SUM)
            src = peer_views[k] if rank == 0 else None
        else:
            r.render(cfg["scene"], tiles=tiles, out=d_film)
            dist.reduce(d_film, dst=0, op=dist.ReduceOp.SUM)
            src = d_film
        if not resident:
            # what a user of N GPUs receives: the finished frame in (pinned) host memory on rank 0
            if rank == 0:
                h_film_t.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
        return r.last_stats

    def rank_times(ms):
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = ms
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]

    def rebalance(frames=12):
        """Warm-up only: move the band boundaries until the slowest rank is within 2 % of the mean."""
        nonlocal tiles
        for _ in range(frames):
            times = rank_times(frame(True)["ms_total"])
            if balancer.imbalance(times) < 1.02 or not balancer.update(times):
                break
            tiles = balancer.tiles_for(rank)

    def timed(resident, steps, warmup):
        if balancer and resident:
            rebalance()
        for _ in range(warmup):
            frame(resident)
        acc = {"ms_trace": 0.0, "ms_shadow": 0.0, "ms_total": 0.0, "ms_raygen": 0.0, "ms_shade": 0.0,
               "ms_film": 0.0, "kernel_launches": 0, "rays": 0, "camera_rays": 0, "shadow_rays": 0}
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            st = frame(resident)
            for k in ("ms_trace", "ms_shadow", "ms_total", "ms_raygen", "ms_shade", "ms_film", "kernel_launches",
                      "camera_rays", "shadow_rays"):
                acc[k] += st[k]
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms, float(acc["camera_rays"]), float(acc["shadow_rays"])], dtype=torch.float64,
                         device="cuda")
        if world > 1:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, cam, sh = float(mx[0]), float(sm[1]), float(sm[2])
        else:
            cam, sh = float(t[1]), float(t[2])
        # Tile halos make ranks re-evaluate a few boundary samples; count every ray of the frame once
        # (scale to the frame's own camera-sample count) so that N-GPU values stay comparable.
        e = cfg["sampler"].ext
        frame_cam = float((e[1] - e[0]) * (e[3] - e[2]) * cfg["sampler"].samples_per_pixel()) * steps
        rays = (cam + sh) * (frame_cam / cam)
        return ms, rays / steps, acc

    clk = ClockSampler(local)
    clk.start()
    t_wait = time.perf_counter()
    while clk.proc and not clk.rows and time.perf_counter() - t_wait < 3.0:
        time.sleep(0.02)  # nvidia-smi needs a moment before its first sample
    ms, rays_per_frame, acc = timed(True, args.steps, args.warmup)
    clocks = clk.stop()
    per_rank = rank_times(acc["ms_total"] / args.steps) if world > 1 else None
    ms_e2e, rays_e2e, _ = timed(False, args.steps, 1)

    # full-frame sanity: the last e2e film must be a plausible image
    if rank == 0:
        wsum = h_film[..., 3]
        assert np.isfinite(h_film).all() and wsum.min() > 0 and h_film[..., :3].max() > 0

    value = rays_per_frame / (ms * 1e-3) / 1e6
    e2e_v = rays_e2e / (ms_e2e * 1e-3) / 1e6
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    roofline, cpu_baseline = None, None
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        # bounded sample: the full frame costs ~7 s on 16 host cores; halve the resolution on
        # smaller hosts so the CPU leg stays within ~10-30 s
        scale = 1 if cores >= 12 else 2
        cfg_s = make_cfg(scale=scale)
        st_c, _, osc, _ = oracle_sample(cfg_s, cores, 0, True)       # counts (default mode)
        b_prim, b_sh = bytes_per_ray(st_c)
        from oracle import orc
        oc = orc.render_config(cfg_s["camera"], cfg_s["sampler"], num_cpus=cores, mode=1, n_threads=cores)
        t0 = time.perf_counter()
        orc.render(osc, oc)
        dt = time.perf_counter() - t0
        sample = ("the full 1920x1080x16spp frame, strict task mode" if scale == 1 else
                  "same scene+camera at 960x540x16spp (1/4 of the frame's pixels), strict task mode")
        cpu_baseline = {"value": (st_c["camera_rays"] + st_c["shadow_rays"]) / dt / 1e6, "unit": "Mrays/s",
                        "cores": cores, "kind": "port", "sample": sample, "seconds": dt}
        n_cam = acc["camera_rays"] / args.steps
        alg_bytes = b_prim * n_cam                                     # per frame, closest-hit kernel
        t_trace = acc["ms_trace"] / args.steps * 1e-3
        achieved = alg_bytes / t_trace / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_trace_closest_bytes_per_frame")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_trace<closest> (all launches of one frame)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_primary_ray": b_prim, "alg_bytes_per_shadow_ray": b_sh,
                    "kernel_ms_per_frame": t_trace * 1e3,
                    "shadow_kernel_ms_per_frame": acc["ms_shadow"] / args.steps,
                    "shadow_achieved": b_sh * (acc["shadow_rays"] / args.steps) / max(1e-9, acc["ms_shadow"] / args.steps * 1e-3) / 1e9}

    line = {
        "metric": "Mrays/s (primary+shadow)", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_frame": rays_per_frame, "tiles": partition,
                   "l2": "per-frame working set (112 MB scene + >2 GB wavefront buffers) exceeds the 126 MB L2; no explicit flush",
                   "scene_build_upload_s": t_scene,
                   "partition": {"whole": "whole film", "cyclic": "64x64 tiles, cyclic",
                                 "bands": "row bands balanced on measured rank times during warm-up"}[partition],
                   "band_rows": balancer.b if balancer else None,
                   "per_rank_device_ms": per_rank,
                   "film_gather": {"none": "single GPU", "p2p": "film kernels store owned tiles into rank 0's HBM over NVLink (CUDA IPC) + barrier",
                                   "nccl": "reduce(SUM) of zero-padded films (NCCL)"}[gather]},
        "clocks": clocks,
        "e2e": {"value": e2e_v, "unit": "Mrays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 1024 + 64 + 44 + 12 + 128 + (16 * len(tiles) if tiles else 0),
                "d2h_bytes_per_step": h * w * 16},
        "gpu_launches": int(acc["kernel_launches"]),
        "stage_ms_per_frame": {k: acc[k] / args.steps for k in ("ms_raygen", "ms_trace", "ms_shade", "ms_shadow", "ms_film", "ms_total")},
    }
    if roofline:
        line["roofline"] = roofline
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
