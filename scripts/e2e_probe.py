"""Where does the host-film (e2e) frame spend its extra time?  Wall clock vs device stage times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pbrt_rust_b200 as pb
import bench

cfg = bench.make_cfg()
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8)
stream = torch.cuda.current_stream()
r.ctx.set_stream(stream.cuda_stream)
r.preprocess(cfg["scene"])
h, w = cfg["film"].shape
d_film = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
h_t = torch.zeros(h * w * 4, dtype=torch.float32).pin_memory()
h_film = h_t.numpy().reshape(h, w, 4)
p_film = np.zeros((h, w, 4), np.float32)
for name, out in (("device", d_film), ("pinned", h_film), ("pageable", p_film)):
    for _ in range(3):
        r.render(cfg["scene"], out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = 0.0
    for _ in range(10):
        st = r.render(cfg["scene"], out=out) is None or r.last_stats
        tot += st["ms_total"]
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 100
    print("%-9s wall %.3f ms/frame   device ms_total %.3f  | last frame: raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f launches %d" % (
        name, wall, tot / 10, st["ms_raygen"], st["ms_trace"], st["ms_shade"], st["ms_shadow"], st["ms_film"], st["kernel_launches"]))
