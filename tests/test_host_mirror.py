"""CPU tests of the product's host side (libpbrtb200.so `pbh_*` + the C-ABI surface) against the
oracle: same BVH node for node, same camera/film/sampler descriptors.  No compute call needs a GPU."""
import ctypes as C
import re
import os

import numpy as np
import pytest

import pbrt_rust_b200 as pb
from pbrt_rust_b200 import _ffi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_library_exports_every_declared_symbol():
    """Every function the two public headers declare is exported by the .so and bound in _ffi."""
    declared = set()
    for h in ("pbrtb200.h", "pbrtb200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        declared |= set(re.findall(r"\b(pbrtb200_[a-z_0-9]+|pbh_[a-z_0-9]+)\s*\(", src))
    declared -= {"pbrtb200_ctx", "pbh_scene"}
    L = pb.lib()
    bound = {n for n, _, _ in _ffi.SYMBOLS}
    for name in sorted(declared):
        assert hasattr(L, name), name
        assert name in bound, f"{name} declared in include/ but not bound in _ffi.SYMBOLS"


def test_struct_sizes_match_the_abi():
    assert C.sizeof(_ffi.Node32) == 32 and C.sizeof(_ffi.Tri48) == 48 and C.sizeof(_ffi.Sphere80) == 80
    assert C.sizeof(_ffi.Film) == 4 * 8 + 256 * 4


def test_ctypes_layouts_match_the_c_header(tmp_path):
    """Every struct of include/pbrtb200.h: size and the offset of every field as the C compiler lays
    them out == the ctypes mirror in _ffi.py (what a Rust #[repr(C)] binding must reproduce too)."""
    import subprocess
    pairs = {"pbrtb200_node32": _ffi.Node32, "pbrtb200_tri48": _ffi.Tri48, "pbrtb200_sphere80": _ffi.Sphere80,
             "pbrtb200_mesh": _ffi.Mesh, "pbrtb200_texture": _ffi.Texture, "pbrtb200_mipmap": _ffi.MipMap,
             "pbrtb200_material": _ffi.Material, "pbrtb200_light": _ffi.Light, "pbrtb200_scene": _ffi.Scene,
             "pbrtb200_camera": _ffi.Camera, "pbrtb200_sampler": _ffi.Sampler, "pbrtb200_film": _ffi.Film,
             "pbrtb200_integrator": _ffi.Integrator, "pbrtb200_tileset": _ffi.TileSet, "pbrtb200_stats": _ffi.Stats}
    header = open(os.path.join(ROOT, "include", "pbrtb200.h")).read()
    declared = set(re.findall(r"}\s*(pbrtb200_[a-z0-9_]+);", header)) - {"pbrtb200_ray32", "pbrtb200_hit16"}
    assert declared == set(pairs), declared ^ set(pairs)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pbrtb200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    assert C.sizeof(_ffi.Texture) == 120 and C.sizeof(_ffi.Material) == 24


def test_rust_ffi_is_generated_from_the_header():
    """integration/rust/src/gpu/ffi.rs (the binding a maintainer adds to the crate) is exactly what
    scripts/gen_rust_ffi.py produces from the current header, declares every pbrtb200_* entry point
    and mirrors every public struct with its size."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "scripts", "gen_rust_ffi.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    text, fns = g.generate()
    assert open(g.OUT).read() == text, "run python scripts/gen_rust_ffi.py"
    header = open(os.path.join(ROOT, "include", "pbrtb200.h")).read()
    declared = set(re.findall(r"\b(pbrtb200_[a-z_0-9]+)\s*\(", header)) - {"pbrtb200_ctx"}
    assert declared == set(fns), declared ^ set(fns)
    for cname in re.findall(r"}\s*(pbrtb200_[a-z0-9_]+);", header):
        assert f"pub struct {cname} " in text, cname
    assert "size_of::<pbrtb200_texture>() == 120" in text and "size_of::<pbrtb200_scene>()" in text


def test_host_check_mode_is_not_part_of_the_product():
    """PB_HOST_CHECK (the host compilation of the device source used by tests/devsrc/) is never
    defined by the product build, and libpbrtb200.so exports none of the harness's entry points."""
    import subprocess
    import __graft_entry__ as g
    assert not any("PB_HOST_CHECK" in f for f in g.NVCC_FLAGS)
    for root_, _, files in os.walk(os.path.join(ROOT, "pbrt_rust_b200")):
        for f in files:
            if f.endswith((".cu", ".cpp", ".py")):
                assert "define PB_HOST_CHECK" not in open(os.path.join(root_, f), errors="ignore").read(), f
    syms = subprocess.check_output(["nm", "-D", "--defined-only", pb.LIB_PATH], text=True)
    assert "devsrc_" not in syms and "orc_" not in syms


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(pb.PbrtError) as e:
        pb.Context(0)
    assert e.value.code == _ffi.ENODEV


@pytest.mark.parametrize("sm", ["sah", "middle", "equal"])
@pytest.mark.parametrize("which", ["c1", "c2", "c3", "c4", "quadrics"])
def test_host_bvh_equals_reference_tree(orc, which, sm):
    """The in-place host builder must emit BVHAccelerator::new's tree (bvh.rs:189-362) node for node
    and the same ordered primitive list; checked bit-exactly against the line-faithful oracle."""
    cfg = {"c1": lambda: scenes.config1(), "c2": lambda: scenes.config2(n=20000),
           "c3": lambda: scenes.config3(nx=80, nz=40),
           "c4": lambda: scenes.config4(n_ground=(40, 20), n_spheres=200),
           "quadrics": lambda: scenes.quadrics()}[which]()
    cfg["scene"].aggregate.sm = sm
    hs = pb.HostScene(cfg["scene"])
    osc = orc.OracleScene(cfg["scene"])
    hn = hs.nodes()
    ob, om = osc.nodes()
    assert len(hn) == len(ob)
    assert np.array_equal(np.concatenate([hn["bmin"], hn["bmax"]], 1).view(np.uint32), ob.view(np.uint32))
    assert np.array_equal(hn["offset"], om[:, 0])
    assert np.array_equal(hn["is_leaf"], om[:, 2])
    assert np.array_equal(np.where(hn["is_leaf"] == 1, hn["count"], hn["axis"]), om[:, 1])
    assert np.array_equal(hs.prim_order(), osc.prim_order())


def test_flatten_vertex_order_is_refine_reversed():
    """shape/mesh.rs:329-331: triangle j has (p1,p2,p3) = (P[vi[3j+2]], P[vi[3j+1]], P[vi[3j]])."""
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    vi = np.array([0, 3, 2, 0, 1, 2, 0, 3, 1, 1, 2, 3], np.uint32)
    mesh = pb.Shape.triangle_mesh(pb.Transform.new(), pb.Transform.new(), False, vi, P)
    hs = pb.HostScene(pb.Scene.new_with(pb.Primitive.bvh([pb.Primitive.geometric(mesh, None)], 1, "sah"), []))
    f = hs.flat.contents
    order = hs.prim_order()
    for i in range(f.n_prims):
        j = int(order[i, 2])
        t = f.tris[i]
        assert list(t.p1) == P[vi[3 * j + 2]].tolist()
        assert list(t.p2) == P[vi[3 * j + 1]].tolist()
        assert list(t.p3) == P[vi[3 * j]].tolist()
        assert t.user == j


def test_bad_mesh_is_rejected():
    mesh = pb.Shape.triangle_mesh(pb.Transform.new(), pb.Transform.new(), False, [0, 1], np.zeros((3, 3)))
    with pytest.raises(pb.PbrtError):
        pb.HostScene(pb.Scene.new_with(pb.Primitive.bvh([pb.Primitive.geometric(mesh, None)], 1, "sah"), []))


def test_singular_matrix_is_an_error_code():
    out = np.zeros(16, np.float32)
    assert pb.lib().pbh_invert(np.zeros(16, np.float32).ctypes.data_as(C.POINTER(C.c_float)),
                               out.ctypes.data_as(C.POINTER(C.c_float))) == _ffi.ESINGULAR


def test_transforms_match_oracle(orc):
    L = orc.lib()
    for kind, args, mk in [(0, (1.5, -2.0, 3.25), lambda a: pb.Transform.translate(a)),
                           (1, (2.0, 0.5, 4.0), lambda a: pb.Transform.scale(*a)),
                           (2, (33.0, 0, 0), lambda a: pb.Transform.rotate_x(a[0])),
                           (3, (90.0, 0, 0), lambda a: pb.Transform.rotate_y(a[0])),
                           (4, (-17.5, 0, 0), lambda a: pb.Transform.rotate_z(a[0]))]:
        m, mi = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
        L.orc_transform(kind, _p(np.array(args, np.float32)), _p(m), _p(mi))
        t = mk(args)
        assert np.array_equal(t.m.view(np.uint32), m.view(np.uint32))
        assert np.array_equal(t.m_inv.view(np.uint32), mi.view(np.uint32))
    m, mi = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
    a = [np.array(v, np.float32) for v in ((1, 1, -6), (1, 1, 1), (0, 1, 0))]
    L.orc_look_at(_p(a[0]), _p(a[1]), _p(a[2]), _p(m), _p(mi))
    t = pb.Transform.look_at(*a)
    assert np.array_equal(t.m.view(np.uint32), m.view(np.uint32))
    assert np.array_equal(t.m_inv.view(np.uint32), mi.view(np.uint32))


def test_camera_film_sampler_descriptors_match_oracle(orc):
    for cfg in (scenes.config1(), scenes.config2(n=10), scenes.config3(nx=4, nz=4),
                scenes.config1(xres=142, yres=12, crop=(1 / 3, 2 / 3, 1 / 3, 2 / 3), filt=pb.Filter.mean(3.0, 3.0)),
                scenes.config1(filt=pb.Filter.gaussian(2.0, 2.0, 1.5)),
                scenes.config1(filt=pb.Filter.mitchell(2.0, 2.5, 1 / 3, 1 / 3)),
                scenes.config1(filt=pb.Filter.lanczos(3.0, 3.0, 3.0)),
                scenes.config1(filt=pb.Filter.triangle(1.5, 2.0))):
        cam, film = cfg["camera"], cfg["film"]
        oc = orc.render_config(cam, cfg["sampler"], num_cpus=8)
        r2c, r2ci, dxdy = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32), np.zeros(6, np.float32)
        assert orc.lib().orc_perspective(C.byref(oc), _p(r2c), _p(r2ci), _p(dxdy)) == 0
        d = cam.desc
        assert np.array_equal(np.array(d.raster_to_camera, np.float32).view(np.uint32), r2c.reshape(-1).view(np.uint32))
        assert np.array_equal(np.array(list(d.dx_camera) + list(d.dy_camera), np.float32).view(np.uint32), dxdy.view(np.uint32))
        lay = orc.layout(oc)
        assert tuple(lay["sample_ext"]) == film.get_sample_extent()
        assert tuple(lay["pixel_ext"]) == film.get_pixel_extent()
        tab = np.zeros(256, np.float32)
        f = film.filter
        orc.lib().orc_filter_table(f.ty, f.xw, f.yw, f.p0, f.p1, _p(tab))
        assert np.array_equal(np.array(film.desc.filter_table, np.float32).view(np.uint32), tab.view(np.uint32))
        r = pb.lib().pbh_num_tasks(8, film.x_res * film.y_res)
        assert r == lay["num_tasks"]


def test_film_extents_kat():
    """camera/film.rs:374-408 through the product's host mirror."""
    f = pb.Film.image(142, 12, pb.Filter.mean(3.0, 3.0), (np.float32(1) / np.float32(3), np.float32(2) / np.float32(3),
                                                           np.float32(1) / np.float32(3), np.float32(2) / np.float32(3)))
    assert f.get_sample_extent() == (45, 98, 1, 11) and f.get_pixel_extent() == (48, 95, 4, 8)
    f = pb.Film.image(142, 12, pb.Filter.mean(3.0, 3.0), (0.0, np.float32(1) / np.float32(3),
                                                           np.float32(1) / np.float32(3), np.float32(2) / np.float32(3)))
    assert f.get_sample_extent() == (-3, 51, 1, 11) and f.get_pixel_extent() == (0, 48, 4, 8)


def test_film_to_rgb_matches_oracle(orc):
    cfg = scenes.config1(xres=32, yres=24)
    osc = orc.OracleScene(cfg["scene"])
    ref = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8))
    assert np.array_equal(pb.film_to_rgb(ref["film"]).view(np.uint32), ref["rgb"].view(np.uint32))


# ---- film output ("next" row 3: Film::write_image, camera/film.rs:15-33, 316-354) ------------

def _oracle_bytes(orc, rgb):
    a = np.ascontiguousarray(rgb, np.float32)
    out = np.zeros(a.shape, np.uint8)
    L = orc.lib()
    L.orc_rgb_to_bytes(C.c_void_p(a.ctypes.data), C.c_uint64(a.size), C.c_void_p(out.ctypes.data))
    return out


def test_rgb_to_bytes_matches_oracle_and_hand_values(orc):
    """write_img's to_byte (film.rs:21-23): known answers + the whole [0, 1.2] range against the
    oracle restatement (same glibc powf on both sides: bit-exact)."""
    hand = np.array([0.0, 1.0, 2.0, -1.0, np.nan, 0.5, 0.2140411], np.float32)
    # 0 -> 0.5 -> 0; 1 -> 255.5 -> 255; 2 -> clamp 255; negative base -> NaN -> 0; NaN -> 0;
    # 0.5^(1/2.2) = 0.72974 -> 186.58 -> 186; 0.2140411^(1/2.2) ~ 0.49626 -> 127.05 -> 127
    assert pb.rgb_to_bytes(hand).tolist() == [0, 255, 255, 0, 0, 186, 127]
    sweep = np.linspace(0.0, 1.2, 200001, dtype=np.float32)
    assert np.array_equal(pb.rgb_to_bytes(sweep), _oracle_bytes(orc, sweep))


def _decode_png(path):
    """Minimal PNG reader (8-bit RGB, filter 0..4) — test-side check of the writer."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        ln, ty = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + ln]
        crc, = struct.unpack(">I", raw[pos + 8 + ln:pos + 12 + ln])
        assert crc == (zlib.crc32(ty + body) & 0xFFFFFFFF), ty
        if ty == b"IHDR":
            w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, comp, filt, inter) == (8, 2, 0, 0, 0)
        elif ty == b"IDAT":
            idat += body
        pos += 12 + ln
    data = zlib.decompress(idat)
    stride = 3 * w
    img = np.zeros((h, stride), np.uint8)
    for y in range(h):
        row = data[y * (stride + 1):(y + 1) * (stride + 1)]
        assert row[0] == 0
        img[y] = np.frombuffer(row[1:], np.uint8)
    return img.reshape(h, w, 3)


def test_write_image_png_round_trip(orc, tmp_path):
    cfg = scenes.config1(xres=48, yres=20)
    osc = orc.OracleScene(cfg["scene"])
    ref = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8))
    path = str(tmp_path / "frame.png")
    pb.write_image(path, ref["film"])
    img = _decode_png(path)
    assert img.shape == (20, 48, 3)
    assert np.array_equal(img, _oracle_bytes(orc, ref["rgb"]).reshape(20, 48, 3))
    assert img.max() > 0  # the spheres are lit
    # float companion
    pfm = str(tmp_path / "frame.pfm")
    pb.write_image(pfm, ref["film"])
    raw = open(pfm, "rb").read()
    head = b"PF\n48 20\n-1.0\n"
    assert raw.startswith(head)
    px = np.frombuffer(raw[len(head):], "<f4").reshape(20, 48, 3)[::-1]
    assert np.array_equal(px.view(np.uint32), ref["rgb"].reshape(20, 48, 3).view(np.uint32))


def test_write_image_rejects_bad_path():
    with pytest.raises(OSError):
        pb.write_image("/nonexistent-dir/x.png", np.ones((2, 2, 4), np.float32))


# ---- ImageTexture / MIPMap ("next" row 2) -----------------------------------------------------

def _host_pyramid(host, mip_index=0):
    """Levels of MIPMap `mip_index` out of the flattened scene's texel pool."""
    f = host.flat.contents
    mm = f.mipmaps[mip_index]
    tex = np.ctypeslib.as_array(f.texels, shape=(f.n_texels * 4,)).reshape(-1, 4)
    off, w, h, out = mm.texel_offset, mm.width, mm.height, []
    for _ in range(mm.n_levels):
        out.append(tex[off:off + w * h, :3].reshape(h, w, 3).copy())
        off += w * h
        w, h = max(w >> 1, 1), max(h >> 1, 1)
    return mm, out


@pytest.mark.parametrize("wrap", ["repeat", "black", "clamp"])
@pytest.mark.parametrize("spectrum", [True, False])
def test_host_mipmap_pyramid_matches_oracle(orc, wrap, spectrum):
    """The product's host-side MIPMap::new (pbh_texture_image) against the oracle's line-faithful
    one: same level count, same sizes, every texel bit-identical — non-power-of-two input, so the
    sinc resize (mipmap.rs:45-106, incl. its in-place t pass) is exercised for all wrap modes."""
    img = scenes.procedural_image(200, 120)
    t = pb.Texture.image(pb.UVMapping2D(1, 1, 0, 0), img, spectrum=spectrum, wrap=wrap, scale=0.9, gamma=2.2)
    mat = pb.Material.matte(t, pb.Texture.constant(0.0))
    tri = pb.Shape.triangle_mesh(pb.Transform.new(), pb.Transform.new(), False, [0, 1, 2],
                                 [[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    scene = pb.Scene.new_with(pb.Primitive.bvh([pb.Primitive.geometric(tri, mat)], 1, "sah"), [])
    host = pb.HostScene(scene)
    mm, levels = _host_pyramid(host)
    ref = orc.OracleMIPMap(img, spectrum=spectrum, wrap=pb.Texture.WRAP[wrap], scale=0.9, gamma=2.2)
    assert (mm.width, mm.height, mm.n_levels) == (256, 128, 9) and ref.levels() == 9
    for i, lv in enumerate(levels):
        r = ref.level(i)
        assert lv.shape == r.shape
        assert np.array_equal(lv.view(np.uint32), r.view(np.uint32)), f"level {i}"
    if not spectrum:
        assert np.array_equal(levels[0][..., 0], levels[0][..., 1])


def test_host_mipmap_unreadable_file_and_png_reader(orc, tmp_path):
    """imagemap.rs:116-120: a file that cannot be read gives the 1x1 scale^gamma map; and the
    package's PNG reader round-trips the product's own PNG writer."""
    t = pb.Texture.image(pb.UVMapping2D(1, 1, 0, 0), str(tmp_path / "missing.png"), scale=0.5, gamma=2.0)
    assert t.image["texels"] is None
    mat = pb.Material.matte(t, pb.Texture.constant(0.0))
    tri = pb.Shape.triangle_mesh(pb.Transform.new(), pb.Transform.new(), False, [0, 1, 2],
                                 [[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    host = pb.HostScene(pb.Scene.new_with(pb.Primitive.bvh([pb.Primitive.geometric(tri, mat)], 1, "sah"), []))
    mm, levels = _host_pyramid(host)
    assert (mm.width, mm.height, mm.n_levels) == (1, 1, 1) and levels[0].ravel().tolist() == [0.25, 0.25, 0.25]
    from pbrt_rust_b200.imageio import read_png_rgb8
    rgb8 = (scenes.procedural_image(37, 21) * 255).round().astype(np.uint8)
    pb.write_image(str(tmp_path / "a.png"), rgb8=rgb8)
    assert np.array_equal(read_png_rgb8(str(tmp_path / "a.png")), rgb8)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "checkerboard_stretched.npz"))["rgb8"]
    pb.write_image(str(tmp_path / "b.png"), rgb8=golden)
    assert np.array_equal(read_png_rgb8(str(tmp_path / "b.png")), golden)


def test_upload_rejects_bad_mipmap_tables():
    """Validation happens before any device call is needed for the table itself."""
    assert C.sizeof(_ffi.MipMap) == 32


def test_pixel_work_list_builder_equals_the_per_pixel_construction(tmp_path):
    """pbh::build_pixel_list (per-interval task tables, separable need masks, threaded tile-major fill)
    against the straightforward per-pixel construction, on 800 random (film, filter, task count, tile
    set) cases including threaded 1080p-sized ones: tests/devsrc/pixel_list_check.cpp."""
    import subprocess
    exe = str(tmp_path / "pixel_list_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "devsrc", "pixel_list_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout[-2000:]


def test_work_list_hook_lists_every_needed_pixel_once_in_tile_order():
    """pbrtb200_work_list (the list pbrtb200_render uploads, as host arithmetic): the whole film visits
    every sampler pixel once, 8 x 4 tiles row-major; a band lists its rows plus the filter halo, halo
    pixels flagged; k is the raster index inside the owning task's window and the index inverts the list."""
    L = _ffi.lib()
    cfg = scenes.config1(xres=50, yres=30, filt=pb.Filter.gaussian(2.0, 2.0, 2.0))
    film, s = cfg["camera"].film, cfg["sampler"]
    nt = int(L.pbh_num_tasks(8, film.x_res * film.y_res))
    smp = _ffi.Sampler(s.kind, s.ext[0], s.ext[1], s.ext[2], s.ext[3], s.xs, s.ys, int(s.jitter), s.sopen, s.sclose, nt)
    sw, sh = s.ext[1] - s.ext[0], s.ext[3] - s.ext[2]

    def work_list(tiles):
        ts = None
        if tiles is not None:
            rects = np.ascontiguousarray(tiles, np.int32)
            ts = _ffi.TileSet(rects.ctypes.data_as(C.POINTER(C.c_int32)), rects.shape[0], 1)
        n = C.c_uint32(0)
        assert L.pbrtb200_work_list(C.byref(smp), C.byref(film.desc), C.byref(ts) if ts else None, C.byref(n), None, None, None, None) == 0
        xy, k, task = np.zeros(n.value, np.int32), np.zeros(n.value, np.uint32), np.zeros(n.value, np.uint32)
        index = np.zeros(sw * sh, np.int32)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        assert L.pbrtb200_work_list(C.byref(smp), C.byref(film.desc), C.byref(ts) if ts else None, C.byref(n), p(xy, C.c_int32),
                                    p(k, C.c_uint32), p(task, C.c_uint32), p(index, C.c_int32)) == 0
        x = (xy & 0xFFFF).astype(np.int16).astype(int)
        y = ((xy >> 16) & 0xFFFF).astype(np.int16).astype(int)
        return x, y, k, task, index.reshape(sh, sw)

    x, y, k, task, index = work_list(None)
    assert x.size == sw * sh and (task >> 31).max() == 0
    assert np.array_equal(index[y - s.ext[2], x - s.ext[0]], np.arange(x.size))          # the index inverts the list
    tile = ((y - s.ext[2]) // 4) * ((sw + 7) // 8) + (x - s.ext[0]) // 8
    assert (np.diff(tile) >= 0).all()                                                      # tiles in row-major order
    for t in np.unique(task):                                                              # k: raster index in the task window
        w = (C.c_int32 * 4)()
        m = task == t
        x0, y0, tw = x[m].min(), y[m].min(), x[m].max() - x[m].min() + 1
        assert np.array_equal(k[m], (y[m] - y0) * tw + (x[m] - x0))
    xb, yb, kb, taskb, indexb = work_list([(0, 8, 50, 16)])                                # a band of film rows 8..15
    halo = (taskb >> 31) == 1
    assert ((yb[~halo] >= 8) & (yb[~halo] < 16) & (xb[~halo] >= 0) & (xb[~halo] < 50)).all()
    assert halo.any() and yb.min() < 8 and yb.max() >= 16 and yb.min() >= 8 - 4 and yb.max() <= 16 + 3
    whole_k = {(a, b): c for a, b, c in zip(x, y, k)}
    assert all(whole_k[(a, b)] == c for a, b, c in zip(xb, yb, kb))                       # same RNG offsets as the whole film
    assert (indexb >= 0).sum() == xb.size
