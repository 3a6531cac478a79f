"""ctypes bindings for libpbrtb200.so (include/pbrtb200.h + include/pbrtb200_host.h).

The library is the product: there is no Python or CPU fallback.  Importing this module fails loudly
when the shared object has not been built (`python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBRTB200_LIB names an alternative build of the same library (A/B measurements of kernel variants
# built by __graft_entry__.build_variant); the default is the product.
LIB_PATH = os.environ.get("PBRTB200_LIB") or os.path.join(_HERE, "libpbrtb200.so")

OK, EINVAL, ENODEV, ENAN, ESTACK, ESINGULAR, ENOMEM = 0, -1, -2, -3, -4, -5, -6
MISS = 0xFFFFFFFF

f32, u32, i32, u64, u16, u8 = C.c_float, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint16, C.c_uint8
P = C.POINTER


class Node32(C.Structure):
    _fields_ = [("bmin", f32 * 3), ("bmax", f32 * 3), ("offset", u32), ("count", u16),
                ("axis", u8), ("is_leaf", u8)]


class Tri48(C.Structure):
    _fields_ = [("p1", f32 * 3), ("mesh", u32), ("p2", f32 * 3), ("attr", u32), ("p3", f32 * 3),
                ("user", u32)]


class Sphere80(C.Structure):
    _fields_ = [("w2o", f32 * 12), ("radius", f32), ("z_min", f32), ("z_max", f32),
                ("phi_max", f32), ("theta_min", f32), ("theta_max", f32), ("material", u32),
                ("flip", u32)]


class Mesh(C.Structure):
    _fields_ = [("o2w", f32 * 12), ("o2w_inv", f32 * 12), ("material", u32), ("area_light", i32),
                ("flip", u32), ("has_uv", u32), ("has_n", u32), ("has_s", u32)]


class Texture(C.Structure):
    _fields_ = [("kind", i32), ("value", f32 * 12), ("map_kind", i32), ("map", f32 * 12),
                ("tex1", i32), ("tex2", i32), ("tex3", i32), ("aa", i32)]


class Material(C.Structure):
    _fields_ = [("kind", i32), ("kd", i32), ("sigma", i32), ("ks", i32), ("roughness", i32),
                ("bump", i32)]


class Light(C.Structure):
    _fields_ = [("kind", i32), ("pos", f32 * 3), ("intensity", f32 * 3), ("w2l", f32 * 12),
                ("cos_total_width", f32), ("cos_falloff_start", f32), ("num_samples", i32),
                ("first_tri", u32), ("n_tris", u32), ("total_area", f32)]


class MipMap(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("n_levels", u32), ("do_trilinear", u32),
                ("max_anisotropy", f32), ("wrap", u32), ("texel_offset", u64)]


class Scene(C.Structure):
    _fields_ = [("nodes", P(Node32)), ("n_nodes", u32), ("leaf_prim", P(u32)), ("n_prims", u32),
                ("tris", P(Tri48)), ("n_tris", u32), ("spheres", P(Sphere80)),
                ("sphere_o2w", P(f32)), ("n_spheres", u32), ("meshes", P(Mesh)), ("n_meshes", u32),
                ("tri_uv", P(f32)), ("tri_n", P(f32)), ("tri_s", P(f32)), ("n_attr", u32),
                ("materials", P(Material)), ("n_materials", u32), ("textures", P(Texture)),
                ("n_textures", u32), ("lights", P(Light)), ("n_lights", u32),
                ("area_prims", P(u32)), ("n_area_prims", u32),
                ("mipmaps", P(MipMap)), ("n_mipmaps", u32), ("texels", P(f32)), ("n_texels", u64)]


class Camera(C.Structure):
    _fields_ = [("raster_to_camera", f32 * 16), ("camera_to_world", f32 * 16),
                ("dx_camera", f32 * 3), ("dy_camera", f32 * 3), ("shutter_open", f32),
                ("shutter_close", f32), ("lens_radius", f32), ("focal_distance", f32)]


class Sampler(C.Structure):
    _fields_ = [("kind", i32), ("x_start", i32), ("x_end", i32), ("y_start", i32), ("y_end", i32),
                ("xs", i32), ("ys", i32), ("jitter", i32), ("shutter_open", f32),
                ("shutter_close", f32), ("num_tasks", i32)]


class Film(C.Structure):
    _fields_ = [("x_res", i32), ("y_res", i32), ("x_pixel_start", i32), ("y_pixel_start", i32),
                ("x_pixel_count", i32), ("y_pixel_count", i32), ("filter_xw", f32),
                ("filter_yw", f32), ("filter_table", f32 * 256)]


class Integrator(C.Structure):
    _fields_ = [("kind", i32), ("max_depth", i32), ("strict_flags", i32)]


class TileSet(C.Structure):
    _fields_ = [("rects", P(i32)), ("n_rects", u32), ("flags", u32)]


class Stats(C.Structure):
    _fields_ = [("camera_rays", u64), ("camera_hits", u64), ("shadow_rays", u64),
                ("ms_total", f32), ("ms_raygen", f32), ("ms_trace", f32), ("ms_shade", f32),
                ("ms_shadow", f32), ("ms_film", f32), ("kernel_launches", u32),
                ("nan_samples", u32), ("stack_overflows", u32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol the two headers declare: (name, restype, argtypes)
_vp = C.c_void_p
_fp = P(f32)
SYMBOLS = [
    ("pbrtb200_create", i32, [C.c_int, P(_vp)]),
    ("pbrtb200_destroy", None, [_vp]),
    ("pbrtb200_last_error", C.c_char_p, [_vp]),
    ("pbrtb200_set_stream", i32, [_vp, _vp]),
    ("pbrtb200_upload_scene", i32, [_vp, P(Scene)]),
    ("pbrtb200_render", i32, [_vp, P(Camera), P(Sampler), P(Film), P(Integrator), P(TileSet), _vp,
                              C.c_int, P(Stats)]),
    ("pbrtb200_trace_closest", i32, [_vp, _vp, u64, _vp, C.c_int, P(Stats)]),
    ("pbrtb200_trace_any", i32, [_vp, _vp, u64, _vp, C.c_int, P(Stats)]),
    ("pbrtb200_halton_layout", i32, [_vp, P(Sampler), P(u32), P(u64)]),
    ("pbrtb200_primary_hits", i32, [_vp, P(Camera), P(Sampler), _vp, _vp, _vp, C.c_int, P(Stats)]),
    ("pbh_translate", None, [_fp, _fp, _fp]),
    ("pbh_scale", None, [f32, f32, f32, _fp, _fp]),
    ("pbh_rotate_x", None, [f32, _fp, _fp]),
    ("pbh_rotate_y", None, [f32, _fp, _fp]),
    ("pbh_rotate_z", None, [f32, _fp, _fp]),
    ("pbh_mul", None, [_fp, _fp, _fp, _fp, _fp, _fp]),
    ("pbh_invert", i32, [_fp, _fp]),
    ("pbh_look_at", i32, [_fp, _fp, _fp, _fp, _fp]),
    ("pbh_scene_new", _vp, []),
    ("pbh_scene_free", None, [_vp]),
    ("pbh_last_error", C.c_char_p, [_vp]),
    ("pbh_texture_constant", i32, [_vp, _fp]),
    ("pbh_texture_checkerboard", i32, [_vp, C.c_int, _fp, C.c_int, C.c_int, C.c_int]),
    ("pbh_texture_uv", i32, [_vp, C.c_int, _fp]),
    ("pbh_texture_image", i32, [_vp, C.c_int, _fp, _fp, u32, u32, C.c_int, C.c_int, f32, C.c_int, f32, f32]),
    ("pbh_mapping_from_transform", i32, [_fp, _fp]),
    ("pbh_texture_scale", i32, [_vp, C.c_int, C.c_int]),
    ("pbh_texture_mix", i32, [_vp, C.c_int, C.c_int, C.c_int]),
    ("pbh_texture_bilerp", i32, [_vp, C.c_int, _fp, _fp, _fp, _fp, _fp]),
    ("pbh_texture_dots", i32, [_vp, C.c_int, _fp, C.c_int, C.c_int]),
    ("pbh_texture_fbm", i32, [_vp, C.c_int, f32, _fp]),
    ("pbh_texture_wrinkled", i32, [_vp, C.c_int, f32, _fp]),
    ("pbh_material_matte", i32, [_vp, C.c_int, C.c_int, C.c_int]),
    ("pbh_material_plastic", i32, [_vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("pbh_light_point", i32, [_vp, _fp, _fp, _fp]),
    ("pbh_light_spot", i32, [_vp, _fp, _fp, _fp, f32, f32]),
    ("pbh_light_area", i32, [_vp, _fp, C.c_int]),
    ("pbh_add_triangle_mesh", i32, [_vp, _fp, _fp, C.c_int, P(u32), u64, _fp, u64, _fp, _fp, _fp,
                                    C.c_int, C.c_int]),
    ("pbh_add_sphere", i32, [_vp, _fp, _fp, C.c_int, f32, f32, f32, f32, C.c_int]),
    ("pbh_add_cylinder", i32, [_vp, _fp, _fp, C.c_int, f32, f32, f32, f32, C.c_int]),
    ("pbh_add_disk", i32, [_vp, _fp, _fp, C.c_int, f32, f32, f32, f32, C.c_int]),
    ("pbh_build_bvh", i32, [_vp, u32, C.c_char_p]),
    ("pbh_flat_scene", P(Scene), [_vp]),
    ("pbh_prim_order", None, [_vp, P(u32)]),
    ("pbh_camera_perspective", i32, [_fp, _fp, f32, f32, f32, f32, f32, C.c_int, C.c_int,
                                     P(Camera)]),
    ("pbh_film_image", i32, [C.c_int, C.c_int, C.c_int, f32, f32, f32, f32, _fp, P(Film)]),
    ("pbh_film_sample_extent", None, [P(Film), P(i32)]),
    ("pbh_num_tasks", u32, [u32, u32]),
    ("pbh_film_to_rgb", None, [_fp, u64, _fp]),
    ("pbh_rgb_to_bytes", None, [_fp, u64, _vp]),
    ("pbh_write_png", i32, [C.c_char_p, _vp, u32, u32]),
    ("pbh_write_pfm", i32, [C.c_char_p, _fp, u32, u32]),
    ("pbrtb200_peer_film_create", i32, [_vp, u64, P(_vp), C.c_char_p]),
    ("pbrtb200_peer_film_open", i32, [_vp, C.c_char_p, P(_vp)]),
    ("pbrtb200_peer_film_close", i32, [_vp, _vp]),
    ("pbrtb200_film_develop", i32, [_vp, _vp, C.c_int, u64, _vp, _vp, C.c_int]),
    ("pbrtb200_cost_profile", i32, [_vp, P(Camera), P(Film), C.c_int, _fp]),
    ("pbrtb200_group_create", i32, [P(C.c_int), C.c_int, P(_vp)]),
    ("pbrtb200_group_destroy", None, [_vp]),
    ("pbrtb200_group_last_error", C.c_char_p, [_vp]),
    ("pbrtb200_group_size", i32, [_vp]),
    ("pbrtb200_group_ctx", _vp, [_vp, C.c_int]),
    ("pbrtb200_group_upload_scene", i32, [_vp, P(Scene)]),
    ("pbrtb200_group_render", i32, [_vp, P(Camera), P(Sampler), P(Film), P(Integrator), _vp, C.c_int, P(Stats)]),
    ("pbrtb200_group_pin_host_film", i32, [_vp, _vp, u64]),
    ("pbrtb200_group_unpin_host_film", i32, [_vp]),
    ("pbrtb200_group_bands", i32, [_vp, P(i32), _fp]),
    ("pbrtb200_group_device_stats", i32, [_vp, C.c_int, P(Stats)]),
    ("pbrtb200_cut_bands", i32, [_fp, C.c_int, C.c_int, C.c_int, P(i32)]),
    ("pbrtb200_work_list", i32, [P(Sampler), P(Film), P(TileSet), P(u32), P(i32), P(u32), P(u32), P(i32)]),
    ("pbrtb200_bands_new", _vp, [_fp, C.c_int, C.c_int, C.c_int]),
    ("pbrtb200_bands_free", None, [_vp]),
    ("pbrtb200_bands_update", i32, [_vp, _fp]),
    ("pbrtb200_bands_get", i32, [_vp, P(i32)]),
]

_lib = None


def lib():
    """Load the shared library once; raise if it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing — build it with __graft_entry__.build(); "
                "pbrt_rust_b200 has no CPU or pure-Python fallback")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
