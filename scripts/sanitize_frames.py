"""Small frames that cover every kernel of the render path, for compute-sanitizer (scripts/gpu_sanitize.sh):
the smoke frame, a wide-filter frame, a multi-light / multi-sample frame (k_fold), a mixed sphere /
triangle textured frame, a frame forced through the sample ring and banded film, a Halton frame, the
trace hooks, and frames through a (single-device) group: staged host film, then a host film pinned with
pbrtb200_group_pin_host_film, which the film kernel writes directly (PBRTB200_HOST_FILM_STORES)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pbrt_rust_b200 as pb
from pbrt_rust_b200 import scenes


def render(cfg, **kw):
    r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, **kw)
    film = r.render(cfg["scene"])
    assert np.isfinite(film).all()
    return r, film


r, _ = render(scenes.config3(nx=60, nz=30, xres=96, yres=64, xs=2, ys=2))
r.primary_hits(scenes.config3(nx=60, nz=30, xres=96, yres=64, xs=2, ys=2)["scene"])
render(scenes.config1(xres=64, yres=48, filt=pb.Filter.gaussian(2.0, 2.0, 2.0)))
render(scenes.config3(nx=40, nz=20, xres=64, yres=48, xs=2, ys=2, n_lights=3, light_samples=2))
render(scenes.config4(n_ground=(30, 15), n_spheres=60, xres=64, yres=36, xs=2, ys=2))
render(scenes.config1(xres=48, yres=32, sampler="halton"))
os.environ["PBRTB200_FRAME_BUDGET_MB"] = "1"
os.environ["PBRTB200_CHUNK_LOG2"] = "11"
render(scenes.config3(nx=40, nz=20, xres=64, yres=48, xs=2, ys=2, n_lights=2, light_samples=1))
del os.environ["PBRTB200_FRAME_BUDGET_MB"], os.environ["PBRTB200_CHUNK_LOG2"]
cfg = scenes.config2(n=2000, xres=32, yres=32)
rr = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8)
rays = np.zeros((4096, 8), np.float32)
rng = np.random.default_rng(1)
rays[:, 0:3] = rng.uniform(-12, 12, (4096, 3))
rays[:, 4:7] = rng.uniform(-1, 1, (4096, 3))
rays[:, 7] = 1e30
rr.intersect(cfg["scene"], rays)
rr.intersect_p(cfg["scene"], rays)
grp = pb.Group([0])
cfg = scenes.config3(nx=40, nz=20, xres=64, yres=48, xs=2, ys=2)
r, film = render(cfg, ctx=grp)
grp.pin_host_film(film)             # page-locked + mapped: k_film stores straight into host memory
r.render(cfg["scene"], out=film)
assert np.isfinite(film).all()
grp.unpin_host_film()
print("sanitize_frames: done")
