"""ctypes wrapper of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It consumes the same pbrt_rust_b200.api description objects as the product
(pure data) but shares no code path with it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborc.so")
_lib = None

f32, u32, i32, u64 = C.c_float, C.c_uint32, C.c_int32, C.c_uint64


class RenderConfig(C.Structure):
    _fields_ = [("cam_to_world", f32 * 16), ("cam_to_world_inv", f32 * 16), ("screen_window", f32 * 4),
                ("sopen", f32), ("sclose", f32), ("lensr", f32), ("focald", f32), ("fov", f32),
                ("x_res", i32), ("y_res", i32), ("crop", f32 * 4), ("filter_type", i32),
                ("filter_xw", f32), ("filter_yw", f32), ("filter_p0", f32), ("filter_p1", f32),
                ("sampler_kind", i32), ("xs", i32), ("ys", i32), ("jitter", i32), ("num_tasks", i32),
                ("num_cpus", i32), ("mode", i32), ("n_threads", i32), ("count_traversal", i32),
                ("primary_only", i32)]


class RenderStats(C.Structure):
    _fields_ = [("camera_rays", u64), ("camera_hits", u64), ("shadow_rays", u64),
                ("nodes_visited", u64), ("tris_tested", u64), ("spheres_tested", u64),
                ("sh_nodes_visited", u64), ("sh_tris_tested", u64), ("sh_spheres_tested", u64),
                ("nan_samples", u64), ("seconds", C.c_double), ("num_tasks", i32),
                ("sample_ext", i32 * 4), ("pixel_ext", i32 * 4)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_new.restype = C.c_void_p
        L.orc_num_nodes.restype = u64
        L.orc_num_prims.restype = u64
        L.orc_num_tasks_for.restype = u32
        L.orc_filter_eval.restype = f32
        L.orc_van_der_corput.restype = f32
        L.orc_sobol2.restype = f32
        L.orc_filter_eval.argtypes = [C.c_int, f32, f32, f32, f32, f32, f32]
        L.orc_filter_table.argtypes = [C.c_int, f32, f32, f32, f32, C.c_void_p]
        L.orc_quadratic.argtypes = [f32, f32, f32, C.c_void_p, C.c_void_p]
        L.orc_sphere_props.argtypes = [f32, f32, f32, f32, C.c_void_p]
        L.orc_sphere_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, f32, f32, f32, f32, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.orc_add_sphere.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, f32, f32, f32, f32, u32]
        L.orc_add_cylinder.argtypes = L.orc_add_sphere.argtypes
        L.orc_add_disk.argtypes = L.orc_add_sphere.argtypes
        L.orc_add_spot_light.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, f32, f32]
        L.orc_get_crop_window.argtypes = [u64, u64, f32, C.c_void_p]
        L.orc_film_extents.argtypes = [C.c_int, C.c_int, f32, f32, C.c_void_p, C.c_void_p, C.c_void_p]
        i32 = C.c_int32
        L.orc_add_image_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, u32, u32, C.c_int, C.c_int,
                                            f32, C.c_int, f32, f32]
        L.orc_mipmap_new.restype = C.c_void_p
        L.orc_mipmap_new.argtypes = [C.c_void_p, u32, u32, C.c_int, C.c_int, f32, C.c_int, f32, f32]
        L.orc_mipmap_free.argtypes = [C.c_void_p]
        L.orc_mipmap_levels.restype = u32
        L.orc_mipmap_levels.argtypes = [C.c_void_p]
        L.orc_mipmap_level_size.argtypes = [C.c_void_p, u32, C.c_void_p]
        L.orc_mipmap_level.argtypes = [C.c_void_p, u32, C.c_void_p]
        L.orc_mipmap_lookup.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_void_p]
        L.orc_image_texture_eval_planar.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_void_p]
        L.orc_radical_inverse.restype = C.c_double
        L.orc_radical_inverse.argtypes = [u64, u64]
        L.orc_halton_cap.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_halton_samples.argtypes = [C.c_void_p, C.c_int, u32, C.c_void_p, C.c_void_p]
        L.orc_noise.restype = f32
        L.orc_noise.argtypes = [f32, f32, f32]
        L.orc_fbm.restype = f32
        L.orc_fbm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, f32, C.c_int]
        L.orc_mapping3d_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sinc_1d.restype = f32
        L.orc_sinc_1d.argtypes = [f32, f32]
        L.orc_modulo.restype = i32
        L.orc_modulo.argtypes = [i32, i32]
        L.orc_rgb_to_bytes.argtypes = [C.c_void_p, u64, C.c_void_p]
        _lib = L
    return _lib


class OracleMIPMap:
    """texture/mipmap.rs MIPMap built the way TextureCache::get_texture does (imagemap.rs:96-173).
    texels: (h, w, 3) float32 = read_image's output, or None for an unreadable file."""

    def __init__(self, texels, spectrum=True, do_trilinear=False, max_aniso=8.0, wrap=0, scale=1.0, gamma=1.0):
        L = lib()
        if texels is None:
            self.h = C.c_void_p(L.orc_mipmap_new(None, 0, 0, int(spectrum), int(do_trilinear), max_aniso, wrap, scale, gamma))
        else:
            t = _f(texels)
            self.h = C.c_void_p(L.orc_mipmap_new(_p(t), t.shape[1], t.shape[0], int(spectrum), int(do_trilinear),
                                                 max_aniso, wrap, scale, gamma))

    def __del__(self):
        try:
            lib().orc_mipmap_free(self.h)
        except Exception:
            pass

    def levels(self):
        return int(lib().orc_mipmap_levels(self.h))

    def level(self, i):
        wh = np.zeros(2, np.uint32)
        lib().orc_mipmap_level_size(self.h, i, _p(wh))
        out = np.zeros((int(wh[1]), int(wh[0]), 3), np.float32)
        lib().orc_mipmap_level(self.h, i, _p(out))
        return out

    def lookup(self, st6):
        q = _f(st6).reshape(-1, 6)
        out = np.zeros((q.shape[0], 3), np.float32)
        lib().orc_mipmap_lookup(self.h, _p(q), q.shape[0], _p(out))
        return out

    def eval_planar(self, p, dpdx=(0, 0, 0), dpdy=(0, 0, 0)):
        """ImageTexture::eval with PlanarMapping2D::new() — the call in imagemap.rs's tests."""
        q = _f(np.concatenate([np.ravel(p), np.ravel(dpdx), np.ravel(dpdy)])).reshape(1, 9)
        out = np.zeros((1, 3), np.float32)
        lib().orc_image_texture_eval_planar(self.h, _p(q), 1, _p(out))
        return out[0]


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    # data_as keeps a reference to the array, so a temporary stays alive until the call returns
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    pass


def _ck(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


class OracleScene:
    """Builds the oracle's own scene + BVH from a pbrt_rust_b200.api.Scene description."""

    def __init__(self, scene):
        L = lib()
        self.h = C.c_void_p(L.orc_scene_new())
        tex_ids, mat_ids, light_ids = {}, {}, {}
        self.tex_ids = tex_ids  # id(api.Texture) -> oracle texture index

        def tex(t):
            if t is None:
                return 0
            if id(t) in tex_ids:
                return tex_ids[id(t)]
            kids = [tex(c) for c in t.children()] + [0, 0, 0]
            a, b, c3 = kids[0], kids[1], kids[2]
            mk = t.mapping.kind if t.mapping is not None else 0
            mp = _f(t.mapping.params if t.mapping is not None else [0] * 12)
            if t.kind == 3:
                im = t.image
                tx = None if im["texels"] is None else _f(im["texels"])
                i = L.orc_add_image_texture(self.h, mk, _p(mp), _p(tx), 0 if tx is None else tx.shape[1],
                                            0 if tx is None else tx.shape[0], int(im["spectrum"]),
                                            int(im["do_trilinear"]), im["max_aniso"], im["wrap"], im["scale"],
                                            im["gamma"])
                tex_ids[id(t)] = i
                return i
            i = L.orc_add_texture(self.h, t.kind, _p(_f(t.value12())), mk, _p(mp), a, b, c3, t.aa)
            tex_ids[id(t)] = i
            return i

        def mat(m):
            if id(m) in mat_ids:
                return mat_ids[id(m)]
            bm = -1 if getattr(m, "bump_map", None) is None else tex(m.bump_map)
            i = L.orc_add_material(self.h, m.kind, tex(m.kd), tex(m.sigma), tex(m.ks), tex(m.roughness), bm)
            mat_ids[id(m)] = i
            return i

        for lt in scene.all_lights():
            if lt.kind == "area":
                i = L.orc_add_area_light(self.h, _p(_f(lt.L)), lt.num_samples)
            elif lt.kind == "point":
                i = L.orc_add_point_light(self.h, _p(_f(lt.l2w.m)), _p(_f(lt.l2w.m_inv)), _p(_f(lt.I)))
            else:
                i = L.orc_add_spot_light(self.h, _p(_f(lt.l2w.m)), _p(_f(lt.l2w.m_inv)), _p(_f(lt.I)),
                                         lt.width, lt.fall)
            light_ids[id(lt)] = i
        agg = scene.aggregate
        for p in agg.prims:
            s = p.shape
            m = mat(p.material) if p.material is not None else 0
            if s.kind == "sphere":
                _ck(L.orc_add_sphere(self.h, _p(_f(s.o2w.m)), _p(_f(s.o2w.m_inv)), int(s.ro), s.rad, s.z0,
                                     s.z1, s.pm, m))
            elif s.kind == "cylinder":
                _ck(L.orc_add_cylinder(self.h, _p(_f(s.o2w.m)), _p(_f(s.o2w.m_inv)), int(s.ro), s.rad, s.z0,
                                       s.z1, s.pm, m))
            elif s.kind == "disk":
                _ck(L.orc_add_disk(self.h, _p(_f(s.o2w.m)), _p(_f(s.o2w.m_inv)), int(s.ro), s.height, s.rad,
                                   s.ri, s.pm, m))
            else:
                al = -1 if p.area_light is None else light_ids[id(p.area_light)]
                _ck(L.orc_add_mesh(self.h, _p(_f(s.o2w.m)), _p(_f(s.o2w.m_inv)), int(s.ro), _p(s.vi),
                                   u64(s.vi.size), _p(s.P), u64(s.P.shape[0]), _p(s.N), _p(s.S), _p(s.uv),
                                   u32(m), i32(al)))
        sm = {"middle": 0, "equal": 1, "sah": 2}.get(agg.sm, 2)
        _ck(L.orc_build(self.h, u32(agg.max_prims), sm))

    def __del__(self):
        try:
            if self.h:
                lib().orc_scene_free(self.h)
                self.h = None
        except Exception:
            pass

    def nodes(self):
        L = lib()
        n = L.orc_num_nodes(self.h)
        b = np.zeros((n, 6), np.float32)
        m = np.zeros((n, 3), np.uint32)
        L.orc_get_nodes(self.h, _p(b), _p(m))
        return b, m

    def prim_order(self):
        L = lib()
        n = L.orc_num_prims(self.h)
        out = np.zeros((n, 3), np.uint32)
        L.orc_get_prim_order(self.h, _p(out))
        return out

    def trace_closest(self, rays, counters=False, n_threads=8):
        rays = _f(rays).reshape(-1, 8)
        n = rays.shape[0]
        prim = np.zeros(n, np.uint32)
        tbb = np.zeros((n, 3), np.float32)
        cnt = np.zeros((n, 3), np.uint32) if counters else None
        _ck(lib().orc_trace_closest(self.h, _p(rays), u64(n), _p(prim), _p(tbb), _p(cnt), n_threads))
        return prim, tbb, cnt

    def trace_any(self, rays, counters=False, early_exit=True, n_threads=8):
        rays = _f(rays).reshape(-1, 8)
        n = rays.shape[0]
        occ = np.zeros(n, np.uint8)
        cnt = np.zeros((n, 3), np.uint32) if counters else None
        _ck(lib().orc_trace_any(self.h, _p(rays), u64(n), _p(occ), _p(cnt), int(early_exit), n_threads))
        return occ, cnt

    def set_strict_flags(self, on):
        lib().orc_set_strict_flags(self.h, int(on))


def render_config(camera, sampler, num_cpus=8, num_tasks=0, mode=0, n_threads=8, count_traversal=False,
                  primary_only=False):
    film = camera.film
    c = RenderConfig()
    c.cam_to_world[:] = _f(camera.cam2world.m).reshape(-1).tolist()
    c.cam_to_world_inv[:] = _f(camera.cam2world.m_inv).reshape(-1).tolist()
    c.screen_window[:] = list(camera.screen_window)
    c.sopen, c.sclose, c.lensr, c.focald, c.fov = camera.sopen, camera.sclose, camera.lensr, camera.focald, camera.fov
    c.x_res, c.y_res = film.x_res, film.y_res
    c.crop[:] = list(film.crop)
    c.filter_type, c.filter_xw, c.filter_yw = film.filter.ty, film.filter.xw, film.filter.yw
    c.filter_p0, c.filter_p1 = film.filter.p0, film.filter.p1
    c.sampler_kind, c.xs, c.ys, c.jitter = sampler.kind, sampler.xs, sampler.ys, int(sampler.jitter)
    c.num_tasks, c.num_cpus, c.mode, c.n_threads = num_tasks, num_cpus, mode, n_threads
    c.count_traversal, c.primary_only = int(count_traversal), int(primary_only)
    return c


def layout(cfg):
    st = RenderStats()
    _ck(lib().orc_render_layout(C.byref(cfg), C.byref(st)))
    return st.as_dict()


def render(oscene, cfg, want_hits=False, strict_flags=False):
    """Returns dict(film=(H,W,4), rgb=(H,W,3), stats, hit_ids, hit_ts)."""
    lay = layout(cfg)
    pe, se = lay["pixel_ext"], lay["sample_ext"]
    h, w = pe[3] - pe[2], pe[1] - pe[0]
    film = np.zeros((h, w, 4), np.float32)
    rgb = np.zeros((h, w, 3), np.float32)
    spp = cfg.xs * cfg.ys if cfg.sampler_kind == 0 else 1 << max(0, (cfg.xs - 1).bit_length())
    if cfg.sampler_kind == 2:  # HaltonSampler: padded layout, cap slots per pixel
        spp = halton_cap(cfg)[0]
    ns = (se[1] - se[0]) * (se[3] - se[2]) * spp
    hit_ids = np.zeros(ns, np.uint32) if want_hits else None
    hit_ts = np.zeros(ns, np.float32) if want_hits else None
    st = RenderStats()
    oscene.set_strict_flags(strict_flags)
    L = lib()
    if strict_flags:
        # orc_render resets the flag; use the dedicated switch inside the call
        pass
    rc = L.orc_render(oscene.h, C.byref(cfg), _p(film), _p(rgb), _p(hit_ids), _p(hit_ts), C.byref(st))
    _ck(rc)
    return dict(film=film, rgb=rgb, stats=st.as_dict(), hit_ids=hit_ids, hit_ts=hit_ts)


def halton_cap(cfg):
    """(cap, per-pixel counts over the full sampler extent) of a HaltonSampler frame."""
    lay = layout(cfg)
    se = lay["sample_ext"]
    counts = np.zeros((se[3] - se[2], se[1] - se[0]), np.uint32)
    cap = C.c_uint32(0)
    _ck(lib().orc_halton_cap(C.byref(cfg), 0, C.byref(cap), _p(counts)))
    return int(cap.value), counts


def halton_samples(cfg, light_pairs=0):
    """Camera samples of a HaltonSampler frame in the padded layout: (H, W, cap, 5) with NaN image
    coordinates in unused slots, and the light-sample floats (H, W, cap, 2 * light_pairs)."""
    cap, counts = halton_cap(cfg)
    h, w = counts.shape
    cs = np.zeros((h, w, cap, 5), np.float32)
    lu = np.zeros((h, w, cap, max(1, 2 * light_pairs)), np.float32)
    _ck(lib().orc_halton_samples(C.byref(cfg), light_pairs, cap, _p(cs), _p(lu) if light_pairs else None))
    return cs, lu, counts


def camera_samples(cfg, light_pairs, x0, x1, y0, y1, spp):
    n = (x1 - x0) * (y1 - y0) * spp
    cs = np.zeros((n, 5), np.float32)
    rays = np.zeros((n, 8), np.float32)
    diff = np.zeros((n, 12), np.float32)
    lu = np.zeros((n, max(1, 2 * light_pairs)), np.float32)
    _ck(lib().orc_camera_samples(C.byref(cfg), light_pairs, x0, x1, y0, y1, _p(cs), _p(rays), _p(diff), _p(lu)))
    return cs, rays, diff, lu
