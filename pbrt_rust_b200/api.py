"""Python mirror of the pbrt_rust objects that sit above the drop-in boundary.

The classes are thin, data-only descriptions with the reference's constructor names and argument
order (file:line cited per class); all arithmetic lives in libpbrtb200.so — the host mirror
(`pbh_*`, C++) builds matrices / BVH / film tables and the CUDA back end (`pbrtb200_*`) renders.
Tests and bench.py drive the product through these classes; the CPU oracle consumes the very same
descriptions (oracle/orc.py), so both sides see identical inputs.
"""
import ctypes as C
import os

import numpy as np

from . import _ffi
from ._ffi import lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class PbrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


# ---------------------------------------------------------------------------------------------
class Transform:
    """src/transform/transform.rs:14-35 — (m, m_inv) pair of row-major 4x4 f32 matrices."""

    def __init__(self, m=None, m_inv=None):
        self.m = _f(np.eye(4) if m is None else m).reshape(4, 4)
        self.m_inv = _f(np.eye(4) if m_inv is None else m_inv).reshape(4, 4)

    @staticmethod
    def new():
        return Transform()

    @staticmethod
    def _mk(fn, *args):
        m, mi = np.zeros(16, np.float32), np.zeros(16, np.float32)
        rc = fn(*args, _fp(m), _fp(mi))
        if rc not in (None, 0):
            raise PbrtError(rc, "Singular matrix!")
        return Transform(m, mi)

    @staticmethod
    def translate(v):
        return Transform._mk(lib().pbh_translate, _fp(_f(v)))

    @staticmethod
    def scale(x, y, z):
        return Transform._mk(lib().pbh_scale, x, y, z)

    @staticmethod
    def rotate_x(deg):
        return Transform._mk(lib().pbh_rotate_x, deg)

    @staticmethod
    def rotate_y(deg):
        return Transform._mk(lib().pbh_rotate_y, deg)

    @staticmethod
    def rotate_z(deg):
        return Transform._mk(lib().pbh_rotate_z, deg)

    @staticmethod
    def look_at(pos, look, up):
        return Transform._mk(lib().pbh_look_at, _fp(_f(pos)), _fp(_f(look)), _fp(_f(up)))

    def inverse(self):
        return Transform(self.m_inv.copy(), self.m.copy())

    def __mul__(self, o):
        m, mi = np.zeros(16, np.float32), np.zeros(16, np.float32)
        lib().pbh_mul(_fp(_f(self.m)), _fp(_f(self.m_inv)), _fp(_f(o.m)), _fp(_f(o.m_inv)), _fp(m), _fp(mi))
        return Transform(m, mi)


class UVMapping2D:
    """src/texture/mapping2d.rs:49-77"""
    kind = 0

    def __init__(self, su=1.0, sv=1.0, du=0.0, dv=0.0):
        self.params = [su, sv, du, dv] + [0] * 8


class PlanarMapping2D:
    """src/texture/mapping2d.rs:175-210"""
    kind = 1

    def __init__(self, vs=(1, 0, 0), vt=(0, 1, 0), ds=0.0, dt=0.0):
        self.params = [*vs, *vt, ds, dt, 0, 0, 0, 0]


def _w2t_params(xf):
    """rows 0..2 of world_to_texture.m; only affine transforms are supported on the device."""
    xf = Transform.new() if xf is None else xf
    out = np.zeros(12, np.float32)
    rc = lib().pbh_mapping_from_transform(_fp(_f(xf.m)), _fp(out))
    if rc:
        raise PbrtError(rc, "world_to_texture must be an affine transform (w row 0 0 0 1)")
    return out.tolist()


class SphericalMapping2D:
    """src/texture/mapping2d.rs:106-143: SphericalMapping2D::new() / new_with(world_to_texture)"""
    kind = 2

    def __init__(self, world_to_texture=None):
        self.params = _w2t_params(world_to_texture)


class CylindricalMapping2D:
    """src/texture/mapping2d.rs:145-173: CylindricalMapping2D::new() / new_with(world_to_texture)"""
    kind = 3

    def __init__(self, world_to_texture=None):
        self.params = _w2t_params(world_to_texture)


class IdentityMapping3D:
    """src/texture/mapping3d.rs:43-66: IdentityMapping3D::new() / new_with(world_to_texture)"""
    kind = 4

    def __init__(self, world_to_texture=None):
        self.params = _w2t_params(world_to_texture)


class Texture:
    """src/texture/mod.rs:52-86 (Constant, Scale), checkerboard.rs:24-95, uv.rs:20-26, mix.rs, bilerp.rs,
    dots.rs, fbm.rs.  Float textures are RGB textures with three equal channels."""

    def __init__(self, kind, value=(0, 0, 0), mapping=None, tex1=None, tex2=None, aa=0, tex3=None):
        self.kind, self.value, self.mapping, self.tex1, self.tex2, self.aa = kind, value, mapping, tex1, tex2, aa
        self.tex3 = tex3

    def children(self):
        return {1: (self.tex1, self.tex2), 4: (self.tex1, self.tex2), 5: (self.tex1, self.tex2, self.tex3),
                7: (self.tex1, self.tex2)}.get(self.kind, ())

    def value12(self):
        v = [float(x) for x in np.ravel(self.value)]
        return v + [0.0] * (12 - len(v))

    @staticmethod
    def constant(v):
        v = (v, v, v) if np.isscalar(v) else tuple(v)
        return Texture(0, v)

    @staticmethod
    def checkerboard(mapping, t1, t2, antialiased=False):
        return Texture(1, mapping=mapping, tex1=t1, tex2=t2, aa=1 if antialiased else 0)

    @staticmethod
    def uv(mapping):
        return Texture(2, mapping=mapping)

    @staticmethod
    def scale(t1, t2):
        """ScaleTexture::new (texture/mod.rs:74-78): t1 * t2"""
        return Texture(4, tex1=t1, tex2=t2)

    @staticmethod
    def mix(t1, t2, amount):
        """MixTexture::new (texture/mix.rs:16-19): t1.lerp(t2, amount)"""
        return Texture(5, tex1=t1, tex2=t2, tex3=amount)

    @staticmethod
    def bilerp(mapping, v00, v01, v10, v11):
        """BilerpTexture::new (texture/bilerp.rs:19-26)"""
        c = [(v, v, v) if np.isscalar(v) else tuple(v) for v in (v00, v01, v10, v11)]
        return Texture(6, value=[x for q in c for x in q], mapping=mapping)

    @staticmethod
    def dots(mapping, inside, outside):
        """DotsTexture::new (texture/dots.rs:17-20)"""
        return Texture(7, mapping=mapping, tex1=inside, tex2=outside)

    @staticmethod
    def fbm(octaves, roughness, mapping=None):
        """FBmTexture::new(oct, roughness, map) (texture/fbm.rs:15-19)"""
        return Texture(8, value=(roughness, 0, 0), mapping=mapping or IdentityMapping3D(), aa=int(octaves))

    @staticmethod
    def wrinkled(octaves, roughness, mapping=None):
        """WrinkledTexture::new(oct, roughness, map) (texture/fbm.rs:36-40)"""
        return Texture(9, value=(roughness, 0, 0), mapping=mapping or IdentityMapping3D(), aa=int(octaves))

    WRAP = {"repeat": 0, "black": 1, "clamp": 2}  # ImageWrap (texture/imagewrap.rs)

    @staticmethod
    def image(mapping, image, spectrum=True, do_trilinear=False, max_aniso=8.0, wrap="repeat", scale=1.0, gamma=1.0):
        """TextureCache::<Spectrum | f32>::new_texture (texture/imagemap.rs:128-138, 183-193).
        `image`: a PNG file name, or an (h, w, 3) array of read_image texels (byte / 255), or None
        (unreadable file -> 1x1 map of scale^gamma, imagemap.rs:116-120)."""
        if isinstance(image, str):
            from .imageio import read_image
            try:
                image = read_image(image)
            except (OSError, ValueError):
                image = None
        t = Texture(3, mapping=mapping)
        t.image = dict(texels=None if image is None else _f(image).reshape(image.shape[0], image.shape[1], 3),
                       spectrum=bool(spectrum), do_trilinear=bool(do_trilinear), max_aniso=float(max_aniso),
                       wrap=Texture.WRAP[wrap] if isinstance(wrap, str) else int(wrap), scale=float(scale),
                       gamma=float(gamma))
        return t


class Material:
    """src/material/mod.rs:79-99; bump_map: Option<ScalarTextureReference> (material::bump, mod.rs:23-77)"""

    def __init__(self, kind, kd, sigma=None, ks=None, roughness=None, bump_map=None):
        self.kind, self.kd, self.sigma, self.ks, self.roughness = kind, kd, sigma, ks, roughness
        self.bump_map = bump_map

    @staticmethod
    def matte(kd, sigma, bump_map=None):
        return Material(0, kd, sigma=sigma, bump_map=bump_map)

    @staticmethod
    def plastic(kd, ks, roughness, bump_map=None):
        return Material(1, kd, ks=ks, roughness=roughness, bump_map=bump_map)


class Shape:
    """src/shape/mod.rs:158-189"""

    def __init__(self, kind, o2w, w2o, ro, **kw):
        self.kind, self.o2w, self.w2o, self.ro = kind, o2w, w2o, bool(ro)
        self.__dict__.update(kw)

    @staticmethod
    def sphere(o2w, w2o, ro, rad, z0, z1, pm):
        return Shape("sphere", o2w, w2o, ro, rad=rad, z0=z0, z1=z1, pm=pm)

    @staticmethod
    def cylinder(o2w, w2o, ro, rad, z0, z1, pm):
        """Shape::cylinder (src/shape/cylinder.rs:27-38)"""
        return Shape("cylinder", o2w, w2o, ro, rad=rad, z0=z0, z1=z1, pm=pm)

    @staticmethod
    def disk(o2w, w2o, ro, height, radius, inner_radius, pm):
        """Shape::disk (src/shape/disk.rs:24-35)"""
        return Shape("disk", o2w, w2o, ro, height=height, rad=radius, ri=inner_radius, pm=pm)

    @staticmethod
    def triangle_mesh(o2w, w2o, ro, vi, P, N=None, S=None, uv=None):
        vi = np.ascontiguousarray(vi, dtype=np.uint32).reshape(-1)
        P = _f(P).reshape(-1, 3)
        N = None if N is None else _f(N).reshape(-1, 3)
        S = None if S is None else _f(S).reshape(-1, 3)
        uv = None if uv is None else _f(uv).reshape(-1, 2)
        return Shape("mesh", o2w, w2o, ro, vi=vi, P=P, N=N, S=S, uv=uv)


class AreaLight:
    """Extension: diffuse area light (the reference's src/area_light.rs is a stub)."""

    def __init__(self, L, num_samples=1):
        self.L = (L, L, L) if np.isscalar(L) else tuple(L)
        self.num_samples = int(num_samples)
        self.kind = "area"


class Light:
    """src/light/point.rs:21-25, src/light/spot.rs:24-35"""

    def __init__(self, kind, l2w, I, width=0.0, fall=0.0):
        self.kind, self.l2w, self.width, self.fall = kind, l2w, width, fall
        self.I = (I, I, I) if np.isscalar(I) else tuple(I)

    @staticmethod
    def point(l2w, I):
        return Light("point", l2w, I)

    @staticmethod
    def spot(l2w, I, width, fall):
        return Light("spot", l2w, I, width, fall)


class Primitive:
    """src/primitive/mod.rs:78-122"""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    @staticmethod
    def geometric(shape, material):
        return Primitive("geometric", shape=shape, material=material, area_light=None)

    @staticmethod
    def geometric_area_light(shape, material, area_light):
        return Primitive("geometric", shape=shape, material=material, area_light=area_light)

    @staticmethod
    def bvh(prims, max_prims, sm):
        return Primitive("bvh", prims=list(prims), max_prims=int(max_prims), sm=sm)


class Scene:
    """src/scene.rs:15-44"""

    def __init__(self, aggregate, lights):
        if aggregate.kind != "bvh":
            raise ValueError("the GPU back end flattens BVH aggregates only (Grid/KdTree: out of scope)")
        self.aggregate, self.lights = aggregate, list(lights)

    @staticmethod
    def new_with(aggregate, lights, volume_region=None):
        if volume_region is not None:
            raise ValueError("volumes are out of scope (VolumeIntegrator is a stub in the reference)")
        return Scene(aggregate, lights)

    def all_lights(self):
        """Scene lights followed by the area lights attached to primitives (first use order)."""
        out = list(self.lights)
        for p in self.aggregate.prims:
            if p.area_light is not None and p.area_light not in out:
                out.append(p.area_light)
        return out


class Filter:
    """src/filter.rs:44-85"""

    def __init__(self, ty, xw, yw, p0=0.0, p1=0.0):
        self.ty, self.xw, self.yw, self.p0, self.p1 = ty, float(xw), float(yw), float(p0), float(p1)

    mean = staticmethod(lambda xw, yw: Filter(0, xw, yw))
    triangle = staticmethod(lambda xw, yw: Filter(1, xw, yw))
    gaussian = staticmethod(lambda xw, yw, a: Filter(2, xw, yw, a))
    mitchell = staticmethod(lambda xw, yw, b, c: Filter(3, xw, yw, b, c))
    lanczos = staticmethod(lambda xw, yw, tau: Filter(4, xw, yw, tau))


class Film:
    """src/camera/film.rs:69-122"""

    def __init__(self, xres, yres, filt, crop):
        self.x_res, self.y_res, self.filter, self.crop = int(xres), int(yres), filt, tuple(float(c) for c in crop)
        self.desc = _ffi.Film()
        rc = lib().pbh_film_image(self.x_res, self.y_res, filt.ty, filt.xw, filt.yw, filt.p0, filt.p1,
                                  _fp(_f(self.crop)), C.byref(self.desc))
        if rc:
            raise PbrtError(rc, "Film::image failed")

    @staticmethod
    def image(xres, yres, filt, crop=(0.0, 1.0, 0.0, 1.0), filename="", open_window=False):
        return Film(xres, yres, filt, crop)

    def get_sample_extent(self):
        e = (C.c_int32 * 4)()
        lib().pbh_film_sample_extent(C.byref(self.desc), e)
        return tuple(e)

    def get_pixel_extent(self):
        d = self.desc
        return (d.x_pixel_start, d.x_pixel_start + d.x_pixel_count, d.y_pixel_start,
                d.y_pixel_start + d.y_pixel_count)

    @property
    def shape(self):
        return (self.desc.y_pixel_count, self.desc.x_pixel_count)


class Camera:
    """src/camera/mod.rs:105-135 (Perspective only; Orthographic/Environment out of scope)"""

    def __init__(self, cam2world, screen_window, sopen, sclose, lensr, focald, fov, film):
        self.cam2world, self.screen_window = cam2world, tuple(float(x) for x in screen_window)
        self.sopen, self.sclose, self.lensr, self.focald, self.fov, self.film = sopen, sclose, lensr, focald, fov, film
        self.desc = _ffi.Camera()
        rc = lib().pbh_camera_perspective(_fp(_f(cam2world.m)), _fp(_f(self.screen_window)), sopen, sclose,
                                          lensr, focald, fov, film.x_res, film.y_res, C.byref(self.desc))
        if rc:
            raise PbrtError(rc, "Singular matrix!")

    @staticmethod
    def perspective(cam2world, screen_window, sopen, sclose, lensr, focald, fov, film):
        return Camera(cam2world, screen_window, sopen, sclose, lensr, focald, fov, film)


class Sampler:
    """src/sampler/mod.rs:30-50"""

    def __init__(self, kind, ext, xs, ys, jitter, sopen, sclose):
        self.kind, self.ext, self.xs, self.ys = kind, tuple(int(e) for e in ext), int(xs), int(ys)
        self.jitter, self.sopen, self.sclose = bool(jitter), float(sopen), float(sclose)

    @staticmethod
    def stratified(x_start, x_end, y_start, y_end, xs, ys, jitter, sopen, sclose):
        return Sampler(0, (x_start, x_end, y_start, y_end), xs, ys, jitter, sopen, sclose)

    @staticmethod
    def low_discrepancy(x_start, x_end, y_start, y_end, spp, sopen, sclose):
        return Sampler(1, (x_start, x_end, y_start, y_end), spp, 1, True, sopen, sclose)

    @staticmethod
    def halton(x_start, x_end, y_start, y_end, spp, sopen, sclose):
        """Sampler::halton (src/sampler/mod.rs:36-40): a variable number of samples per pixel."""
        return Sampler(2, (x_start, x_end, y_start, y_end), spp, 1, True, sopen, sclose)

    def samples_per_pixel(self):
        if self.kind == 0:
            return self.xs * self.ys
        if self.kind == 2:
            return self.xs
        p = 1
        while p < self.xs:
            p <<= 1
        return p


class SurfaceIntegrator:
    """src/integrator/mod.rs:148-162"""

    def __init__(self, max_depth, strict_flags=False):
        self.max_depth, self.strict_flags = int(max_depth), bool(strict_flags)

    @staticmethod
    def whitted(max_depth):
        return SurfaceIntegrator(max_depth)


# ---------------------------------------------------------------------------------------------
class HostScene:
    """Runs the host mirror: Primitive::bvh(...) + the flatten shim -> pbrtb200_scene."""

    def __init__(self, scene):
        L = lib()
        self.h = L.pbh_scene_new()
        self.scene = scene
        tex_ids, mat_ids, self.light_ids = {}, {}, {}
        self.tex_ids = tex_ids  # id(api.Texture) -> index into the flat texture table

        def tex(t):
            if t is None:
                return 0
            if id(t) in tex_ids:
                return tex_ids[id(t)]
            if t.kind == 0:
                i = L.pbh_texture_constant(self.h, _fp(_f(t.value)))
            elif t.kind == 1:
                a, b = tex(t.tex1), tex(t.tex2)
                i = L.pbh_texture_checkerboard(self.h, t.mapping.kind, _fp(_f(t.mapping.params)), a, b, t.aa)
            elif t.kind == 3:
                im = t.image
                tx = im["texels"]
                i = L.pbh_texture_image(self.h, t.mapping.kind, _fp(_f(t.mapping.params)),
                                        None if tx is None else _fp(tx), 0 if tx is None else tx.shape[1],
                                        0 if tx is None else tx.shape[0], int(im["spectrum"]),
                                        int(im["do_trilinear"]), im["max_aniso"], im["wrap"], im["scale"], im["gamma"])
                if i < 0:
                    raise PbrtError(L.pbh_last_error(self.h).decode())
            elif t.kind == 2:
                i = L.pbh_texture_uv(self.h, t.mapping.kind, _fp(_f(t.mapping.params)))
            elif t.kind == 4:
                a, b = tex(t.tex1), tex(t.tex2)
                i = L.pbh_texture_scale(self.h, a, b)
            elif t.kind == 5:
                a, b, c = tex(t.tex1), tex(t.tex2), tex(t.tex3)
                i = L.pbh_texture_mix(self.h, a, b, c)
            elif t.kind == 6:
                v = _f(t.value12())
                i = L.pbh_texture_bilerp(self.h, t.mapping.kind, _fp(_f(t.mapping.params)), _fp(v[0:3].copy()),
                                         _fp(v[3:6].copy()), _fp(v[6:9].copy()), _fp(v[9:12].copy()))
            elif t.kind == 7:
                a, b = tex(t.tex1), tex(t.tex2)
                i = L.pbh_texture_dots(self.h, t.mapping.kind, _fp(_f(t.mapping.params)), a, b)
            elif t.kind in (8, 9):
                fn = L.pbh_texture_fbm if t.kind == 8 else L.pbh_texture_wrinkled
                i = fn(self.h, t.aa, float(t.value[0]), _fp(_f(t.mapping.params)))
                if i < 0:
                    raise PbrtError(i, L.pbh_last_error(self.h).decode())
            else:
                raise PbrtError(_ffi.EINVAL, f"unknown texture kind {t.kind}")
            tex_ids[id(t)] = i
            return i

        def mat(m):
            if id(m) in mat_ids:
                return mat_ids[id(m)]
            bm = -1 if m.bump_map is None else tex(m.bump_map)
            if m.kind == 0:
                i = L.pbh_material_matte(self.h, tex(m.kd), tex(m.sigma), bm)
            else:
                i = L.pbh_material_plastic(self.h, tex(m.kd), tex(m.ks), tex(m.roughness), bm)
            mat_ids[id(m)] = i
            return i

        for lt in scene.all_lights():
            if lt.kind == "area":
                i = L.pbh_light_area(self.h, _fp(_f(lt.L)), lt.num_samples)
            elif lt.kind == "point":
                i = L.pbh_light_point(self.h, _fp(_f(lt.l2w.m)), _fp(_f(lt.l2w.m_inv)), _fp(_f(lt.I)))
            else:
                i = L.pbh_light_spot(self.h, _fp(_f(lt.l2w.m)), _fp(_f(lt.l2w.m_inv)), _fp(_f(lt.I)), lt.width, lt.fall)
            self.light_ids[id(lt)] = i
        agg = scene.aggregate
        for p in agg.prims:
            s = p.shape
            m = mat(p.material) if p.material is not None else 0
            if s.kind == "sphere":
                rc = L.pbh_add_sphere(self.h, _fp(_f(s.o2w.m)), _fp(_f(s.o2w.m_inv)), int(s.ro), s.rad, s.z0, s.z1, s.pm, m)
            elif s.kind == "cylinder":
                rc = L.pbh_add_cylinder(self.h, _fp(_f(s.o2w.m)), _fp(_f(s.o2w.m_inv)), int(s.ro), s.rad, s.z0, s.z1, s.pm, m)
            elif s.kind == "disk":
                rc = L.pbh_add_disk(self.h, _fp(_f(s.o2w.m)), _fp(_f(s.o2w.m_inv)), int(s.ro), s.height, s.rad, s.ri, s.pm, m)
            else:
                al = -1 if p.area_light is None else self.light_ids[id(p.area_light)]
                npn = lambda a: None if a is None else _fp(a)
                rc = L.pbh_add_triangle_mesh(
                    self.h, _fp(_f(s.o2w.m)), _fp(_f(s.o2w.m_inv)), int(s.ro),
                    s.vi.ctypes.data_as(C.POINTER(C.c_uint32)), s.vi.size, _fp(s.P), s.P.shape[0],
                    npn(s.N), npn(s.S), npn(s.uv), m, al)
            if rc < 0:
                self._fail(rc)
        rc = L.pbh_build_bvh(self.h, agg.max_prims, agg.sm.encode())
        if rc:
            self._fail(rc)
        self.flat = L.pbh_flat_scene(self.h)

    def _fail(self, rc):
        raise PbrtError(rc, lib().pbh_last_error(self.h).decode())

    def nodes(self):
        f = self.flat.contents
        return np.ctypeslib.as_array(C.cast(f.nodes, C.POINTER(C.c_uint8)), shape=(f.n_nodes * 32,)).copy().view(
            np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("offset", "<u4"), ("count", "<u2"),
                      ("axis", "u1"), ("is_leaf", "u1")]))

    def prim_order(self):
        n = self.flat.contents.n_prims
        out = np.zeros((n, 3), np.uint32)
        lib().pbh_prim_order(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def close(self):
        if self.h:
            lib().pbh_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One pbrtb200_ctx (one CUDA device)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().pbrtb200_create(device, C.byref(self.h))
        if rc:
            raise PbrtError(rc, lib().pbrtb200_last_error(None).decode())
        self.device = device
        self.scene_key, self.host_scene = None, None

    def check(self, rc):
        if rc:
            raise PbrtError(rc, lib().pbrtb200_last_error(self.h).decode())

    def set_stream(self, cuda_stream):
        """Issue this ctx's work on the given cudaStream_t (int handle), e.g. torch's current stream."""
        self.check(lib().pbrtb200_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def upload(self, host_scene, scene_key=None):
        """Uploads a flattened scene.  The ctx remembers WHICH scene it holds (`scene_key`,
        `host_scene`): renderers that share a ctx compare against that, not against their own
        last upload, so one renderer can never render another renderer's scene by accident."""
        self.scene_key, self.host_scene = None, None
        self.check(lib().pbrtb200_upload_scene(self.h, host_scene.flat))
        self.scene_key, self.host_scene = scene_key, host_scene

    def close(self):
        if self.h:
            lib().pbrtb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """pbrtb200_group: every GPU in `devices` behind one render call (include/pbrtb200.h), the
    counterpart of the reference's thread pool behind SamplerRenderer::render
    (src/sampler_renderer.rs:168-173).  Duck-types Context for GpuRenderer."""

    def __init__(self, devices):
        self.devices = list(devices)
        self.h = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        rc = lib().pbrtb200_group_create(arr, len(self.devices), C.byref(self.h))
        if rc:
            raise PbrtError(rc, lib().pbrtb200_group_last_error(None).decode())
        self.scene_key, self.host_scene = None, None
        self._pinned = None

    def check(self, rc):
        if rc:
            raise PbrtError(rc, lib().pbrtb200_group_last_error(self.h).decode())

    def upload(self, host_scene, scene_key=None):
        self.scene_key, self.host_scene = None, None
        self.check(lib().pbrtb200_group_upload_scene(self.h, host_scene.flat))
        self.scene_key, self.host_scene = scene_key, host_scene

    def bands(self):
        """(row bounds of the last frame, per-device device ms)"""
        n = len(self.devices)
        b, t = (C.c_int32 * (n + 1))(), np.zeros(n, np.float32)
        self.check(lib().pbrtb200_group_bands(self.h, b, _fp(t)))
        return list(b), t.tolist()

    def device_stats(self):
        """Per-device stats dicts of the last frame."""
        out = []
        for i in range(len(self.devices)):
            st = _ffi.Stats()
            self.check(lib().pbrtb200_group_device_stats(self.h, i, C.byref(st)))
            out.append(st.as_dict())
        return out

    def pin_host_film(self, film):
        """Page-locks and maps `film` (a numpy array) for every device, so that the film kernels store
        into it directly.  The group holds a reference until unpin_host_film() / close(): the buffer
        cannot be freed while it is registered."""
        self.check(lib().pbrtb200_group_pin_host_film(self.h, C.c_void_p(film.ctypes.data), film.nbytes))
        self._pinned = film

    def unpin_host_film(self):
        self.check(lib().pbrtb200_group_unpin_host_film(self.h))
        self._pinned = None

    def close(self):
        if self.h:
            lib().pbrtb200_group_destroy(self.h)   # (unpins)
            self.h = None
        self._pinned = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DevicePtr:
    """A raw device address (e.g. a peer GPU's film mapped with pbrtb200_peer_film_open)."""

    def __init__(self, addr):
        self.addr = int(addr)


def _ptr(a):
    """numpy array -> host pointer ; torch CUDA tensor / DevicePtr -> device pointer"""
    if a is None:
        return None, 0
    if isinstance(a, DevicePtr):
        return C.c_void_p(a.addr), 1
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data), 0
    return C.c_void_p(a.data_ptr()), (1 if a.is_cuda else 0)


def _like(dev_tensor, n, dtype):
    """An uninitialised output buffer of n elements on the device `dev_tensor` lives on (torch)."""
    import torch
    return torch.empty(n, dtype={np.uint8: torch.uint8}[dtype], device=dev_tensor.device)


class GpuRenderer:
    """Drop-in for SamplerRenderer (src/sampler_renderer.rs:26-54): same constructor arguments
    (sampler, camera, surface integrator; the volume integrator is a stub in the reference), and
    `render(scene)` replaces `Renderer::render` (src/renderer.rs:9)."""

    def __init__(self, sampler, camera, surf, vol=None, num_cpus=8, device=0, ctx=None, devices=None):
        """devices: a list of CUDA device indices renders every frame on all of them (Group)."""
        self.sampler, self.camera, self.surf = sampler, camera, surf
        film = camera.film
        # sampler_renderer.rs:39-44 (num_cpus::get() is a property of the host running the crate)
        self.num_tasks = int(lib().pbh_num_tasks(num_cpus, film.x_res * film.y_res))
        self.ctx = ctx or (Group(devices) if devices is not None and len(devices) > 1 else
                           Context(devices[0] if devices else device))
        self.last_stats = None

    @property
    def host_scene(self):
        return self.ctx.host_scene

    def sampler_desc(self):
        s = self.sampler
        return _ffi.Sampler(s.kind, s.ext[0], s.ext[1], s.ext[2], s.ext[3], s.xs, s.ys, int(s.jitter),
                            s.sopen, s.sclose, self.num_tasks)

    def halton_layout(self):
        """(slots per pixel, real samples) of a HaltonSampler frame over the whole sampler extent."""
        cap, n = C.c_uint32(0), C.c_uint64(0)
        smp = self.sampler_desc()
        self.ctx.check(lib().pbrtb200_halton_layout(self.ctx.h, C.byref(smp), C.byref(cap), C.byref(n)))
        return int(cap.value), int(n.value)

    def preprocess(self, scene):
        """Builds/flattens/uploads the scene once (the reference builds its BVH at scene creation)."""
        if self.ctx.scene_key is not scene:
            self.ctx.upload(HostScene(scene), scene_key=scene)

    def render(self, scene, tiles=None, out=None, keep_others=False):
        """Returns the film as an (H, W, 4) array: sum(w*XYZ), sum(w).  `out` may be a CUDA tensor
        (float32, H*W*4) or a DevicePtr to keep the film in HBM.  keep_others: with `tiles`, leave
        the pixels outside the tiles untouched (another GPU owns them) instead of zeroing them."""
        self.preprocess(scene)
        film = self.camera.film
        h, w = film.shape
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        ts = None
        if tiles is not None:  # an EMPTY tile list is a tile set without pixels, not "the whole film"
            rects = np.ascontiguousarray(tiles, dtype=np.int32).reshape(-1, 4)
            ts = _ffi.TileSet(rects.ctypes.data_as(C.POINTER(C.c_int32)), rects.shape[0], 1 if keep_others else 0)
        integ = _ffi.Integrator(0, self.surf.max_depth, int(self.surf.strict_flags))
        st = _ffi.Stats()
        smp = self.sampler_desc()
        p, is_dev = _ptr(out)
        if isinstance(self.ctx, Group):
            if ts is not None:
                raise PbrtError(_ffi.EINVAL, "a Group partitions the film itself: tiles are not accepted")
            rc = lib().pbrtb200_group_render(self.ctx.h, C.byref(self.camera.desc), C.byref(smp), C.byref(film.desc),
                                             C.byref(integ), p, is_dev, C.byref(st))
        else:
            rc = lib().pbrtb200_render(self.ctx.h, C.byref(self.camera.desc), C.byref(smp), C.byref(film.desc),
                                       C.byref(integ), C.byref(ts) if ts is not None else None, p, is_dev,
                                       C.byref(st))
        self.last_stats = st.as_dict()
        self.ctx.check(rc)
        return out

    def primary_hits(self, scene, want_samples=False, want_rays=False):
        """Config-2 hook: per camera sample (raster order over the sampler extent) the closest hit."""
        self.preprocess(scene)
        s = self.sampler
        per_pixel = s.samples_per_pixel()
        if s.kind == 2:  # HaltonSampler: padded layout, `cap` slots per pixel (pbrtb200_halton_layout)
            per_pixel = self.halton_layout()[0]
        n = (s.ext[1] - s.ext[0]) * (s.ext[3] - s.ext[2]) * per_pixel
        hits = np.zeros(n, dtype=HIT_DTYPE)
        smp_out = np.zeros((n, 5), np.float32) if want_samples else None
        rays = np.zeros((n, 8), np.float32) if want_rays else None
        st = _ffi.Stats()
        smp = self.sampler_desc()
        rc = lib().pbrtb200_primary_hits(self.ctx.h, C.byref(self.camera.desc), C.byref(smp),
                                         C.c_void_p(hits.ctypes.data),
                                         None if smp_out is None else C.c_void_p(smp_out.ctypes.data),
                                         None if rays is None else C.c_void_p(rays.ctypes.data), 0, C.byref(st))
        self.last_stats = st.as_dict()
        self.ctx.check(rc)
        return hits, smp_out, rays

    # Scene::intersect / intersect_p over a batch (src/scene.rs:60-67)
    def intersect(self, scene, rays, hits=None):
        self.preprocess(scene)
        n = rays.shape[0]
        pr, dev = _ptr(rays)
        if hits is None:
            hits = _like(rays, n * 16, np.uint8) if dev else np.zeros(n, dtype=HIT_DTYPE)
        st = _ffi.Stats()
        ph, hdev = _ptr(hits)
        if hdev != dev:  # the C ABI takes ONE is_device flag for both buffers
            raise PbrtError(_ffi.EINVAL, "intersect: rays and hits must both be host arrays or both be device tensors")
        rc = lib().pbrtb200_trace_closest(self.ctx.h, pr, n, ph, dev, C.byref(st))
        self.last_stats = st.as_dict()
        self.ctx.check(rc)
        return hits

    def intersect_p(self, scene, rays, occluded=None):
        self.preprocess(scene)
        n = rays.shape[0]
        pr, dev = _ptr(rays)
        if occluded is None:
            occluded = _like(rays, n, np.uint8) if dev else np.zeros(n, np.uint8)
        st = _ffi.Stats()
        po, odev = _ptr(occluded)
        if odev != dev:
            raise PbrtError(_ffi.EINVAL, "intersect_p: rays and occluded must both be host arrays or both be device tensors")
        rc = lib().pbrtb200_trace_any(self.ctx.h, pr, n, po, dev, C.byref(st))
        self.last_stats = st.as_dict()
        self.ctx.check(rc)
        return occluded


    def cost_profile(self, scene, stride=4):
        """pbrtb200_cost_profile: relative cost of each film row (what Group balances its bands with)."""
        self.preprocess(scene)
        film = self.camera.film
        out = np.zeros(film.shape[0], np.float32)
        h = lib().pbrtb200_group_ctx(self.ctx.h, 0) if isinstance(self.ctx, Group) else self.ctx.h
        rc = lib().pbrtb200_cost_profile(h, C.byref(self.camera.desc), C.byref(film.desc), stride, _fp(out))
        self.ctx.check(rc)
        return out

    def develop(self, film, want_rgb=False):
        """Film::write_image's pixel pipeline on the device (film.rs:316-354 as intended, D6):
        film (H, W, 4) host array or CUDA tensor -> (H, W, 3) uint8 [, (H, W, 3) float32 RGB]."""
        pf, dev = _ptr(film)
        shape = tuple(film.shape[:-1]) if len(film.shape) > 1 else (film.shape[0] // 4,)
        n = int(np.prod(shape))
        rgb8 = np.zeros(shape + (3,), np.uint8)
        rgb = np.zeros(shape + (3,), np.float32) if want_rgb else None
        rc = lib().pbrtb200_film_develop(self.ctx.h, pf, dev, n, None if rgb is None else C.c_void_p(rgb.ctypes.data),
                                         C.c_void_p(rgb8.ctypes.data), 0)
        self.ctx.check(rc)
        return (rgb8, rgb) if want_rgb else rgb8


HIT_DTYPE = np.dtype([("prim", "<u4"), ("t", "<f4"), ("b1", "<f4"), ("b2", "<f4")])


def rgb_to_bytes(rgb):
    """write_img's quantisation (film.rs:21-23): (255 * p^(1/2.2) + 0.5).clamp(0, 255) as u8."""
    a = _f(rgb)
    out = np.zeros(a.shape, np.uint8)
    lib().pbh_rgb_to_bytes(_fp(a), a.size, C.c_void_p(out.ctypes.data))
    return out


def write_image(filename, film=None, rgb8=None):
    """Film::write_image (film.rs:316-354 as intended): develop `film` (H, W, 4) on the host
    mirror — or take an already developed (H, W, 3) uint8 image — and write an 8-bit PNG; a name
    ending in .pfm stores the linear float RGB instead."""
    name = os.fsencode(filename)
    if filename.lower().endswith(".pfm"):
        rgb = np.ascontiguousarray(film_to_rgb(film), np.float32)
        h, w = rgb.shape[:2]
        rc = lib().pbh_write_pfm(name, _fp(rgb), w, h)
    else:
        if rgb8 is None:
            rgb8 = rgb_to_bytes(film_to_rgb(film))
        rgb8 = np.ascontiguousarray(rgb8, np.uint8)
        h, w = rgb8.shape[:2]
        rc = lib().pbh_write_png(name, C.c_void_p(rgb8.ctypes.data), w, h)
    if rc != 0:
        raise OSError(f"cannot write {filename} (rc={rc})")


def film_to_rgb(xyzw):
    """Film::write_image's pixel conversion as intended (film.rs:331-346)."""
    a = _f(xyzw).reshape(-1, 4)
    out = np.zeros((a.shape[0], 3), np.float32)
    lib().pbh_film_to_rgb(_fp(a), a.shape[0], _fp(out))
    return out.reshape(*np.shape(xyzw)[:-1], 3)
