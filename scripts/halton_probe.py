"""Config 3 at full size with the HaltonSampler (16 samples per pixel on average): stage times."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pbrt_rust_b200 as pb
from pbrt_rust_b200 import scenes

cfg = scenes.config3()
e = cfg["sampler"].ext
out = {}
for name, smp in (("stratified 4x4", cfg["sampler"]), ("halton 16", pb.Sampler.halton(e[0], e[1], e[2], e[3], 16, 0.0, 0.0))):
    r = pb.GpuRenderer(smp, cfg["camera"], cfg["integrator"], num_cpus=8)
    for _ in range(3):
        film = r.render(cfg["scene"])
    st = r.last_stats
    out[name] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}
    if smp.kind == 2:
        out[name]["cap_and_real_samples"] = r.halton_layout()
    out[name]["mean_rgb"] = float(pb.film_to_rgb(film).mean())
print(json.dumps(out, indent=1))
