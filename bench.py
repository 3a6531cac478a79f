#!/usr/bin/env python
"""bench.py — benchmark of the B200 rendering back end (contract in the task prompt).

A "step" is one full frame of the hot path.  Default workload = BASELINE.json config 3: a procedural
1 M-triangle heightfield + quad area light, Whitted / direct lighting, 1920x1080, Stratified 4x4 =
16 spp, box filter — 33.2 M camera rays + their shadow rays per frame.  `--config c1|c2|c4|c5` runs
the other BASELINE configs at their named sizes through the same code.

  python bench.py --gpus N --steps K --warmup W [--config c3]   # our arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...        # CPU restatement of the reference

metric   Mrays/s (primary + shadow), whole job.  `value`: scene resident, film left in HBM.  `e2e`:
         the same frame through the C ABI into a pinned HOST film (D2H inside the timed region).
N > 1    rank 0 drives all N GPUs through pbrtb200_group_render (one process, one host thread per
         device; include/pbrtb200.h); the other ranks only join the barriers.  Every device renders a
         row band and copies its own rows to the host film over its own PCIe link.
roofline the traversal kernels are bound by instruction issue, not by HBM (the 112 MB scene lives in
         L2 / L1): achieved = warp instructions per second of k_trace<closest> (instructions per ray
         from the committed ncu capture profiles/r02_counters.json x rays of the timed frames / its
         measured device time), peak = 148 SMs x 4 schedulers x the SM clock sampled during the run.
         The measured DRAM traffic and the HBM fraction it implies are reported next to it.
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

WORKLOADS = {
    "c1": "config1: 8 spheres + point light (aggregates/mod.rs:82-91 fixture), Whitted, 640x480, 4 spp",
    "c2": "config2: BVH over 100K random triangles, primary closest hit only, 1920x1080, 1 spp",
    "c3": "config3: 1M-tri heightfield + quad area light, direct lighting, 1920x1080, 16 spp",
    "c4": "config4: 200K-tri ground + 20K spheres, textured matte/plastic, point + area light, 3840x2160, 64 spp",
    "c5": "config5: 50M-tri heightfield, 4 area lights, 1920x1080, 256 spp",
}


def make_cfg(config="c3", scale=1):
    from pbrt_rust_b200 import scenes
    if config == "c1":
        return scenes.config1(xres=640 // scale, yres=480 // scale)
    if config == "c2":
        return scenes.config2(xres=1920 // scale, yres=1080 // scale)
    if config == "c4":
        return scenes.config4(xres=3840 // scale, yres=2160 // scale)
    if config == "c5":
        return scenes.config5(xres=1920 // scale, yres=1080 // scale)
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


def bytes_per_ray(st):
    """B_ray = 32 N_nodes + 48 N_tri + 80 N_sph + 48, averaged over the sample."""
    prim = (32 * st["nodes_visited"] + 48 * st["tris_tested"] + 80 * st["spheres_tested"]) / max(1, st["camera_rays"]) + 48
    sh = (32 * st["sh_nodes_visited"] + 48 * st["sh_tris_tested"] + 80 * st["sh_spheres_tested"]) / max(1, st["shadow_rays"]) + 48
    return prim, sh


def cpu_scene(config, scale, cores, mode, count=False):
    """(oracle scene, RenderConfig, shares_product_code).  Config 3 is built from oracle/refscene.py,
    which needs liborc.so and numpy only; the other configs describe their scene through the api.py
    data classes, whose constructors call the host mirror inside libpbrtb200.so (transforms, film and
    camera descriptors — no rendering code)."""
    from oracle import orc
    if config == "c3":
        from oracle import refscene
        h, c = refscene.config3(nx=1000, nz=500, xres=1920 // scale, yres=1080 // scale, xs=4, ys=4, num_cpus=cores,
                                mode=mode, n_threads=cores, count_traversal=count)
        return h, c, False
    cfg = make_cfg(config, scale)
    c = orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=cores, mode=mode, n_threads=cores, count_traversal=count,
                          primary_only=(config == "c2"))
    return orc.OracleScene(cfg["scene"]), c, True


# bounded CPU sample per config: (scale of the frame's resolution) so that one oracle frame stays in the
# ~10-30 s range on 16 cores; c4 / c5 at full size would take minutes
CPU_SCALE = {"c1": 1, "c2": 1, "c3": 1, "c4": 8, "c5": 8}


def run_reference(args):
    """CPU arm: the C++ restatement of the reference (oracle) in reference-faithful (strict) mode — the
    reference's own task split, per-task StdRng and sub-films, num_tasks worker tasks on the host's
    cores — one frame of the workload (or a bounded sample of it) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    cores = os.cpu_count() or 1
    scale = CPU_SCALE[args.config] * (2 if (cores < 12 and args.config == "c3") else 1)
    scene, rc, shares = cpu_scene(args.config, scale, cores, 1)
    lay = orc.layout(rc)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.render(scene, rc)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    rc.count_traversal = 1  # one counted frame: shadow rays are only counted with the traversal counters on
    st = orc.render(scene, rc)["stats"]
    rays = st["camera_rays"] + st["shadow_rays"]
    ms = 1e3 * float(np.mean(times))
    v = rays / (ms * 1e-3) / 1e6
    sample = ("the full frame" if scale == 1 else f"same scene and camera at 1/{scale} of the resolution "
              f"(1/{scale * scale} of the pixels)") + f" per step, strict task mode ({lay['num_tasks']} tasks)"
    line = {
        "impl": "reference", "metric": "Mrays/s (primary+shadow)", "value": v, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOADS[args.config], "sample": sample, "same_config": scale == 1,
                                        "shares_code_with_gpu_arm": shares},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "busy_threads": min(cores, lay["num_tasks"]),
                         "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of pbrt_rust's algorithm (oracle); the Rust crate cannot be built here. Strict mode keeps the "
                "reference's task count (ceil(log2(..)), sampler_renderer.rs:41-44): only that many threads are ever busy.",
    }
    print(json.dumps(line))


def cpu_baseline_leg(config, n_cam_per_frame):
    """The CPU oracle timed on this box's host cores on a bounded sample of the workload: strict mode
    (the reference's own task count) and an all-cores run (row bands over every core), plus the
    traversal counts of SURVEY §8d."""
    from oracle import orc
    cores = os.cpu_count() or 1
    scale = CPU_SCALE[config] * (2 if (cores < 12 and config == "c3") else 1)
    scene, rc, _ = cpu_scene(config, scale, cores, 0, count=True)
    t0 = time.perf_counter()
    st_c = orc.render(scene, rc)["stats"]                 # counts, default mode (all cores)
    b_prim, b_sh = bytes_per_ray(st_c)
    rays = st_c["camera_rays"] + st_c["shadow_rays"]
    rc.count_traversal = 0
    rc.mode = 1
    lay = orc.layout(rc)
    t0 = time.perf_counter()
    orc.render(scene, rc)
    dt_strict = time.perf_counter() - t0
    rc.mode = 0
    t0 = time.perf_counter()
    orc.render(scene, rc)
    dt_all = time.perf_counter() - t0
    sample = "the full frame" if scale == 1 else f"same scene and camera at 1/{scale} of the resolution (1/{scale * scale} of the pixels)"
    return {"value": rays / dt_strict / 1e6, "unit": "Mrays/s", "cores": cores, "busy_threads": min(cores, lay["num_tasks"]),
            "kind": "port", "sample": sample + f", strict task mode ({lay['num_tasks']} tasks as the reference computes them)",
            "seconds": dt_strict,
            "all_cores": {"value": rays / dt_all / 1e6, "unit": "Mrays/s", "busy_threads": cores, "seconds": dt_all,
                          "mode": "same frame, row bands over every host core (not the reference's task split)"}}, b_prim, b_sh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
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
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # one NCCL collective proves the N ranks see each other over NVLink; after that the ranks meet
        # on a gloo (CPU) group: an NCCL barrier is a kernel that SPINS on its GPU until every rank
        # arrives, and ranks 1..N-1 would spin on the very GPUs rank 0 is rendering on
        t = torch.ones(1, device=torch.device("cuda", local))
        dist.all_reduce(t)
        assert int(t.item()) == world
        cpu_group = dist.new_group(backend="gloo")
    torch.cuda.set_device(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank != 0:
        # Ranks 1..N-1 hold their GPU for the job and join the barriers around the timed regions; rank 0
        # drives every GPU through the C ABI's group call (one process, one host thread per device).
        for _ in range(4):
            barrier()
        dist.destroy_process_group()
        return

    import pbrt_rust_b200 as pb
    cfg = make_cfg(args.config)
    film = cfg["film"]
    devices = list(range(world))
    r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, devices=devices)
    t0 = time.perf_counter()
    r.preprocess(cfg["scene"])  # host BVH build + flatten + upload to every device (once, like scene creation)
    t_scene = time.perf_counter() - t0
    h, w = film.shape
    primary_only = args.config == "c2"
    d_film = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda:0")
    h_film_t = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory()
    h_film = h_film_t.numpy().reshape(h, w, 4)
    spp = cfg["sampler"].samples_per_pixel()
    e = cfg["sampler"].ext
    n_samples = (e[1] - e[0]) * (e[3] - e[2]) * spp
    hits_host = np.zeros(n_samples, dtype=pb.HIT_DTYPE) if primary_only else None

    def frame(resident):
        if primary_only:  # config 2: camera samples -> closest hits (hit records to the host in the e2e leg)
            r.primary_hits(cfg["scene"])
            return r.last_stats
        r.render(cfg["scene"], out=d_film if resident else h_film)
        return r.last_stats

    def timed(resident, steps, warmup):
        for _ in range(warmup):
            frame(resident)
        acc = {"ms_trace": 0.0, "ms_shadow": 0.0, "ms_total": 0.0, "ms_raygen": 0.0, "ms_shade": 0.0,
               "ms_film": 0.0, "kernel_launches": 0, "camera_rays": 0, "shadow_rays": 0}
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            st = frame(resident)
            for k in acc:
                acc[k] += st[k]
        torch.cuda.synchronize()   # (every frame call is synchronous: the film is complete when it returns)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        barrier()
        # Band halos make devices re-evaluate a few boundary samples; count every ray of the frame once
        # (scale to the frame's own camera-sample count) so that N-GPU values stay comparable.
        cam = float(acc["camera_rays"])
        rays = (cam + float(acc["shadow_rays"])) * (float(n_samples) * steps / max(cam, 1.0))
        return ms, rays / steps, acc

    clk = ClockSampler(0)
    clk.start()
    t_wait = time.perf_counter()
    while clk.proc and not clk.rows and time.perf_counter() - t_wait < 3.0:
        time.sleep(0.02)  # nvidia-smi needs a moment before its first sample
    # warm-up frames also let the group settle its row bands on measured device times
    ms, rays_per_frame, acc = timed(True, args.steps, max(args.warmup, 16) if world > 1 else args.warmup)
    clocks = clk.stop()
    bands = r.ctx.bands() if world > 1 else None
    # (N > 1: the host-film frames have their own per-device times — each film kernel writes over its own
    # PCIe link — so the bands settle again before the timed region)
    ms_e2e, rays_e2e, acc_e = timed(False, args.steps, 16 if world > 1 else 2)

    if not primary_only:  # full-frame sanity: the last e2e film must be a plausible image
        wsum = h_film[..., 3]
        assert np.isfinite(h_film).all() and wsum.min() > 0 and h_film[..., :3].max() > 0

    value = rays_per_frame / (ms * 1e-3) / 1e6
    e2e_v = rays_e2e / (ms_e2e * 1e-3) / 1e6

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    counters = {}
    try:
        counters = json.load(open(os.path.join(ROOT, "profiles", "r02_counters.json"))).get(args.config, {})
    except Exception:
        pass

    n_cam = acc["camera_rays"] / args.steps
    t_trace = acc["ms_trace"] / args.steps * 1e-3
    sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions / s
    ipr = counters.get("closest_warp_inst_per_ray")
    traffic = counters.get("closest_dram_bytes_per_launch")
    achieved = (ipr * n_cam / t_trace / 1e9) if (ipr and t_trace > 0 and world == 1) else None
    roofline = {
        "bound": "issue", "kernel": "k_trace<closest> (all launches of one frame)",
        "achieved": achieved, "peak": issue_peak, "unit": "G warp-inst/s",
        "frac": (achieved / issue_peak) if achieved else None,
        "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
        "warp_inst_per_ray": ipr, "counters_source": "profiles/r02_counters.json (ncu smsp__inst_executed.sum of the same command)" if ipr else None,
        "traffic": traffic,
        "hbm": {"achieved": (traffic / t_trace / 1e9) if (traffic and t_trace > 0 and world == 1) else None, "peak": hbm_peak,
                "unit": "GB/s", "peak_source": peak_src,
                "frac": (traffic / t_trace / 1e9 / hbm_peak) if (traffic and t_trace > 0 and world == 1) else None},
        "kernel_ms_per_frame": t_trace * 1e3, "shadow_kernel_ms_per_frame": acc["ms_shadow"] / args.steps,
        "shadow_warp_inst_per_ray": counters.get("any_warp_inst_per_ray"),
    }
    cpu_baseline = None
    if not args.no_cpu and world == 1:
        cpu_baseline, b_prim, b_sh = cpu_baseline_leg(args.config, n_cam)
        roofline["algorithmic_bytes_per_primary_ray"] = b_prim   # SURVEY §8d; far above the DRAM traffic because
        roofline["algorithmic_bytes_per_shadow_ray"] = b_sh      # the scene is L2 / L1-resident
        roofline["algorithmic_gbps"] = b_prim * n_cam / t_trace / 1e9 if t_trace > 0 else None

    line = {
        "metric": "Mrays/s (primary+shadow)", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "rays_per_frame": rays_per_frame,
                   "l2": "per-frame working set (scene + >1 GB of wavefront buffers) exceeds the 126 MB L2; no explicit flush",
                   "scene_build_upload_s": t_scene,
                   "partition": "whole film" if world == 1 else "row bands of equal cost (cost probe, then measured device times), one per GPU",
                   "band_rows": bands[0] if bands else None, "per_device_ms": bands[1] if bands else None,
                   "film_gather": "single GPU" if world == 1 else
                                  "value: film kernels store their rows into GPU 0's film over NVLink (peer access); "
                                  "e2e: every GPU's film kernel stores its own rows into the pinned host film over its own PCIe link",
                   "driver": "one process drives all GPUs (pbrtb200_group_render); ranks > 0 only join the barriers" if world > 1 else "single context"},
        "clocks": clocks,
        "e2e": {"value": e2e_v, "unit": "Mrays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": (1024 + 64 + 44 + 12 + 128) * world,
                "d2h_bytes_per_step": int(n_samples * 16) if primary_only else h * w * 16},
        "gpu_launches": int(acc["kernel_launches"]),
        "stage_ms_per_frame": {k: acc[k] / args.steps for k in ("ms_raygen", "ms_trace", "ms_shade", "ms_shadow", "ms_film", "ms_total")},
        "roofline": roofline,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
