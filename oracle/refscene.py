"""Config 3 (the benchmark frame) built for the CPU oracle WITHOUT the product: only liborc.so and the
pure-numpy generators of pbrt_rust_b200/procgen.py, loaded by file path (importing the package would
load libpbrtb200.so).  Used by bench.py's `--impl reference` arm so that the two arms share no code.
TEST / BENCH INFRASTRUCTURE ONLY — mirrors pbrt_rust_b200.scenes.config3 + oracle.orc.render_config;
tests/test_oracle_kat.py checks that both descriptions render the same film."""
import ctypes as C
import importlib.util
import os

import numpy as np

from . import orc

_HERE = os.path.dirname(os.path.abspath(__file__))


def _procgen():
    spec = importlib.util.spec_from_file_location("_pbrt_procgen", os.path.join(_HERE, "..", "pbrt_rust_b200", "procgen.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _Handle:
    """Owns an OrcScene* (duck-types oracle.orc.OracleScene for orc.render)."""

    def __init__(self, h):
        self.h = h

    def set_strict_flags(self, on):
        orc.lib().orc_set_strict_flags(self.h, int(on))

    def __del__(self):
        try:
            if self.h:
                orc.lib().orc_scene_free(self.h)
                self.h = None
        except Exception:
            pass


def config3(nx=1000, nz=500, xres=1920, yres=1080, xs=4, ys=4, num_cpus=8, mode=1, n_threads=8, count_traversal=False):
    """(scene handle, RenderConfig) of scenes.config3: 2*nx*nz-triangle heightfield, matte Kd 0.5, one 4x4
    quad area light at y = 8 (L = 15, 1 sample), BVH sah/4, camera look_at((0,12,-13),(0,0,-2),(0,1,0)),
    fov 36, box filter 0.5, Stratified xs*ys jittered."""
    L = orc.lib()
    pg = _procgen()
    f, p = orc._f, orc._p
    ident = f(np.eye(4))
    h = C.c_void_p(L.orc_scene_new())
    kd = L.orc_add_texture(h, 0, p(f([0.5] * 3 + [0] * 9)), 0, p(f([0] * 12)), 0, 0, 0, 0)
    sg = L.orc_add_texture(h, 0, p(f([0.0] * 12)), 0, p(f([0] * 12)), 0, 0, 0, 0)
    mat = L.orc_add_material(h, 0, kd, sg, 0, 0, -1)
    light = L.orc_add_area_light(h, p(f([15.0, 15.0, 15.0])), 1)
    vi, P = pg.heightfield(nx, nz)
    orc._ck(L.orc_add_mesh(h, p(ident), p(ident), 0, p(vi), orc.u64(vi.size), p(P), orc.u64(P.shape[0]), None, None, None,
                           orc.u32(mat), orc.i32(-1)))
    lvi, lP = pg.quad_light()
    orc._ck(L.orc_add_mesh(h, p(ident), p(ident), 0, p(lvi), orc.u64(lvi.size), p(lP), orc.u64(lP.shape[0]), None, None, None,
                           orc.u32(mat), orc.i32(light)))
    orc._ck(L.orc_build(h, orc.u32(4), 2))
    # camera_to_world = look_at(...).inverse(): look_at returns (world->camera, its inverse)
    w2c, c2w = np.zeros(16, np.float32), np.zeros(16, np.float32)
    L.orc_look_at(p(f([0, 12, -13])), p(f([0, 0, -2])), p(f([0, 1, 0])), p(w2c), p(c2w))
    c = orc.RenderConfig()
    c.cam_to_world[:] = c2w.tolist()
    c.cam_to_world_inv[:] = w2c.tolist()
    aspect = xres / yres
    c.screen_window[:] = [-aspect, aspect, -1.0, 1.0] if aspect > 1 else [-1.0, 1.0, -1.0 / aspect, 1.0 / aspect]
    c.sopen, c.sclose, c.lensr, c.focald, c.fov = 0.0, 0.0, 0.0, 1e6, 36.0
    c.x_res, c.y_res = xres, yres
    c.crop[:] = [0.0, 1.0, 0.0, 1.0]
    c.filter_type, c.filter_xw, c.filter_yw, c.filter_p0, c.filter_p1 = 0, 0.5, 0.5, 0.0, 0.0
    c.sampler_kind, c.xs, c.ys, c.jitter = 0, xs, ys, 1
    c.num_tasks, c.num_cpus, c.mode, c.n_threads = 0, num_cpus, mode, n_threads
    c.count_traversal, c.primary_only = int(count_traversal), 0
    return _Handle(h), c
