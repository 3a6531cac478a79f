"""The device source of the texture stage, compiled for the host, against the oracle (CPU only).

pbrt_rust_b200/csrc/shade_tex.cuh (mappings, noise, every texture kind except image lookups, bump
mapping) is plain arithmetic; tests/devsrc/device_source_host.cpp compiles that very source with g++
(PB_HOST_CHECK) so its logic is checked here, where no GPU exists.  Texture tables come from the
product's host mirror (the flatten shim), descriptions from scenes.TexGen; the oracle gets the same
descriptions.  The `-m gpu` suite repeats the comparison on the device through whole renders.
This is a checker only: nothing in the product loads libdevsrc.so.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pbrt_rust_b200 as pb
from pbrt_rust_b200 import _ffi, scenes
from pbrt_rust_b200.api import HostScene, Material, Primitive, Scene, Shape, Texture, Transform

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "devsrc", "device_source_host.cpp")
LIB = os.path.join(HERE, "devsrc", "libdevsrc.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "pbrt_rust_b200", "csrc", f) for f in ("shade_tex.cuh", "shade_mip.cuh", "dmath.cuh", "halton_math.cuh", "trace_math.cuh", "trace_core.cuh", "host_logic.hpp", "leaf_ref.h", "scene.cuh", "shade_math.cuh", "film_math.cuh", "film.cuh", "shade.cuh", "trace.cuh", "raygen.cuh", "halton.cuh")] + \
    [os.path.join(HERE, "..", "include", "pbrtb200.h")]


@pytest.fixture(scope="module")
def dev():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-Wno-unknown-pragmas", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.devsrc_noise.restype = C.c_float
    L.devsrc_noise.argtypes = [C.c_float] * 3
    L.devsrc_fbm.restype = C.c_float
    L.devsrc_fbm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
    L.devsrc_tex_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.devsrc_tex_eval_img.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.devsrc_bump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.devsrc_radical_inverse.restype = C.c_double
    L.devsrc_radical_inverse.argtypes = [C.c_uint64, C.c_uint32]
    L.devsrc_halton_image.argtypes = [C.c_void_p, C.c_float, C.c_uint64, C.c_void_p]
    L.devsrc_halton_prime.restype = C.c_uint32
    L.devsrc_slab.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.devsrc_tri_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.devsrc_quadratic.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.devsrc_solve2x2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.devsrc_coordinate_system.argtypes = [C.c_void_p, C.c_void_p]
    L.devsrc_chacha12_block.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.devsrc_bsdf_f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.devsrc_fresnel_dielectric.restype = C.c_float
    L.devsrc_fresnel_dielectric.argtypes = [C.c_float] * 3
    L.devsrc_compute_differentials.argtypes = [C.c_void_p] * 3
    L.devsrc_vis_segment.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
    L.devsrc_quadric_dg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    L.devsrc_van_der_corput.restype = C.c_float
    L.devsrc_van_der_corput.argtypes = [C.c_uint32, C.c_uint32]
    L.devsrc_sobol2.restype = C.c_float
    L.devsrc_sobol2.argtypes = [C.c_uint32, C.c_uint32]
    L.devsrc_stream_floats.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    L.devsrc_shuffle.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int]
    L.devsrc_film_weights.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    L.devsrc_camera_ray.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.devsrc_tri_surface.argtypes = [C.c_void_p] * 7 + [C.c_float] * 3 + [C.c_void_p] * 2
    L.devsrc_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    L.devsrc_film_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.devsrc_raygen.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.devsrc_halton.restype = C.c_uint64
    L.devsrc_halton.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.devsrc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _scene_of(materials):
    """one tiny triangle per material: enough for the host mirror / oracle to register every texture"""
    prims = []
    for i, m in enumerate(materials):
        P = np.array([[i, 0, 0], [i + 0.5, 0, 0], [i, 0.5, 0]], np.float32)
        prims.append(Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), False,
                                                             np.arange(3, dtype=np.uint32), P), m))
    return Scene.new_with(Primitive.bvh(prims, 1, "sah"), [])


def _random_dg(rng):
    """struct DG as 30 floats: p, nn, u, v, dpdu, dpdv, dndu, dndv, dpdx, dpdy, dudx, dudy, dvdx, dvdy"""
    g = np.zeros(30, np.float32)
    g[0:3] = rng.uniform(-4, 4, 3)
    dpdu, dpdv = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    nn = np.cross(dpdu, dpdv)
    g[3:6] = nn / np.linalg.norm(nn)
    g[6:8] = rng.uniform(0, 1, 2)
    g[8:11], g[11:14] = dpdu, dpdv
    g[14:17], g[17:20] = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3)
    scale = 10.0 ** rng.uniform(-3, 0)
    g[20:23], g[23:26] = rng.uniform(-1, 1, 3) * scale, rng.uniform(-1, 1, 3) * scale
    g[26:30] = rng.uniform(-1, 1, 4) * scale
    if rng.integers(6) == 0:
        g[20:30] = 0.0  # no differentials
    return g


def _dg15(g):
    return np.concatenate([g[0:3], g[20:23], g[23:26], g[6:8], g[26:30]]).astype(np.float32)


def _dgs33(g, flip):
    return np.concatenate([g[0:3], g[8:11], g[11:14], g[14:17], g[17:20], g[3:6], g[6:8], g[26:30],
                           [1.0 if flip else 0.0, 0, 0], g[20:23], g[23:26]]).astype(np.float32)


def test_noise_and_fbm_match_the_oracle_bit_for_bit(dev, orc):
    """noise is +,-,* only: the device source must agree with the oracle exactly; fbm / turbulence
    add one log2f, the same glibc call on both sides here."""
    L = orc.lib()
    rng = np.random.default_rng(7)
    for _ in range(4000):
        x, y, z = (np.float32(v) for v in rng.uniform(-300, 300, 3))
        assert dev.devsrc_noise(x, y, z) == L.orc_noise(x, y, z)
    for _ in range(1500):
        p = rng.uniform(-20, 20, 3).astype(np.float32)
        dx = (rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(-4, 0)).astype(np.float32)
        dy = (rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(-4, 0)).astype(np.float32)
        om, octs, turb = np.float32(rng.uniform(0.2, 0.9)), int(rng.integers(0, 9)), int(rng.integers(2))
        assert dev.devsrc_fbm(turb, _p(p), _p(dx), _p(dy), om, octs) == L.orc_fbm(turb, _p(p), _p(dx), _p(dy), om, octs)
    zero = np.zeros(3, np.float32)  # s2 = 0: log2(0) = -inf -> all octaves
    assert dev.devsrc_fbm(0, _p(p), _p(zero), _p(zero), 0.5, 5) == L.orc_fbm(0, _p(p), _p(zero), _p(zero), 0.5, 5)


def test_device_texture_evaluator_matches_the_oracle(dev, orc):
    """Random texture trees over every mapping and kind (no images), random shading geometry:
    tex_eval_ext of the device source == the oracle's Texture::evaluate restatement, bit for bit
    (both sides call the same libm here; the GPU run allows the CUDA-libm tolerance)."""
    from oracle import orc as O
    rng = np.random.default_rng(2024)
    gen = scenes.TexGen(rng, None, ext=True, images=False)
    texs = [gen.spectrum_tex() for _ in range(150)] + [gen.unit_tex() for _ in range(40)] + \
        [gen.float_tex(0.02, 0.4) for _ in range(30)] + [gen.noise_tex() for _ in range(30)]
    kinds = set()

    def walk(t):
        kinds.add((t.kind, t.mapping.kind if t.mapping is not None else -1))
        for c in t.children():
            walk(c)

    for t in texs:
        walk(t)
    assert {k for k, _ in kinds} == {0, 1, 2, 4, 5, 6, 7, 8, 9}, kinds
    assert {m for _, m in kinds} >= {0, 1, 2, 3, 4}
    scene = _scene_of([Material.matte(t, Texture.constant(0.0)) for t in texs])
    hs, osc = HostScene(scene), O.OracleScene(scene)
    flat = hs.flat.contents
    table = C.cast(flat.textures, C.c_void_p)
    n = 0
    for t in texs:
        hid, oid = hs.tex_ids[id(t)], osc.tex_ids[id(t)]
        for _ in range(40):
            g = _random_dg(rng)
            got, want, q = np.zeros(3, np.float32), np.zeros(3, np.float32), _dg15(g)
            dev.devsrc_tex_eval(table, hid, _p(g), _p(got))
            O.lib().orc_texture_eval(osc.h, oid, _p(q), _p(want))
            assert got.tolist() == want.tolist(), (t.kind, got, want)
            n += 1
    assert n == 40 * len(texs)


def test_device_bump_matches_the_oracle(dev, orc):
    """material::bump of the device source == the oracle's, over random displacement maps"""
    from oracle import orc as O
    rng = np.random.default_rng(99)
    gen = scenes.TexGen(rng, None, ext=True, images=False)
    bumps = []
    while len(bumps) < 60:
        b = gen.bump_tex()
        if b is not None:
            bumps.append(b)
    scene = _scene_of([Material.matte(Texture.constant(0.5), Texture.constant(0.0), bump_map=b) for b in bumps])
    hs, osc = HostScene(scene), O.OracleScene(scene)
    flat = hs.flat.contents
    assert all(flat.materials[i].bump > 0 for i in range(flat.n_materials))
    table = C.cast(flat.textures, C.c_void_p)
    for b in bumps:
        hid, oid = hs.tex_ids[id(b)], osc.tex_ids[id(b)]
        for _ in range(25):
            g = _random_dg(rng)
            flip = bool(rng.integers(2))
            ng = g[3:6] * np.float32(-1.0 if rng.integers(2) else 1.0)
            got, want, q = np.zeros(9, np.float32), np.zeros(9, np.float32), _dgs33(g, flip)
            dev.devsrc_bump(table, hid, _p(g), _p(ng), int(flip), _p(got))
            O.lib().orc_bump(osc.h, oid, _p(q), _p(ng), _p(want))
            assert got.tolist() == want.tolist()


def test_flat_texture_table_layout():
    """the host mirror fills the fields the device evaluator reads (include/pbrtb200.h)"""
    a, b = Texture.constant((1, 2, 3)), Texture.constant(0.25)
    mix = Texture.mix(a, Texture.scale(a, b), b)
    bil = Texture.bilerp(pb.api.UVMapping2D(2, 3, 0.5, 0.25), 1.0, 2.0, 3.0, (4, 5, 6))
    fbm = Texture.fbm(5, 0.5, pb.api.IdentityMapping3D(Transform.translate((1, 2, 3))))
    dots = Texture.dots(pb.api.SphericalMapping2D(), a, b)
    scene = _scene_of([Material.matte(mix, b, bump_map=fbm), Material.plastic(bil, a, b), Material.matte(dots, b)])
    hs = HostScene(scene)
    f = hs.flat.contents
    t = lambda x: f.textures[hs.tex_ids[id(x)]]
    assert t(mix).kind == 5 and (t(mix).tex1, t(mix).tex3) == (hs.tex_ids[id(a)], hs.tex_ids[id(b)])
    assert f.textures[t(mix).tex2].kind == 4
    assert t(bil).kind == 6 and list(t(bil).value) == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 5, 6]
    assert list(t(bil).map)[:4] == [2, 3, 0.5, 0.25] and t(bil).map_kind == 0
    assert t(fbm).kind == 8 and t(fbm).aa == 5 and t(fbm).value[0] == 0.5 and t(fbm).map_kind == 4
    assert list(t(fbm).map) == [1, 0, 0, 1, 0, 1, 0, 2, 0, 0, 1, 3]
    assert t(dots).kind == 7 and t(dots).map_kind == 2 and list(t(dots).map) == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    assert [f.materials[i].bump for i in range(3)] == [hs.tex_ids[id(fbm)] + 1, 0, 0]
    with pytest.raises(pb.api.PbrtError):  # projective world_to_texture: not representable in map[12]
        m = np.eye(4, dtype=np.float32)
        m[3, 0] = 0.5
        pb.api.SphericalMapping2D(Transform(m, m))
    with pytest.raises(pb.api.PbrtError):  # f32::clamp(0.0, max) panics for max < 0 (noise.rs:116)
        HostScene(_scene_of([Material.matte(Texture.fbm(-1, 0.5), b)]))


def test_device_halton_arithmetic_matches_the_oracle(dev, orc):
    """radical_inverse_ and halton_image of csrc/halton_math.cuh against the oracle's HaltonSampler
    (sampler/halton.rs:57-76, montecarlo.rs:7-20): f64 radical inverses bit for bit, and for whole
    task windows the same accepted candidates with the same image positions."""
    L = orc.lib()
    rng = np.random.default_rng(5)
    primes = [dev.devsrc_halton_prime(k) for k in range(40)]
    assert primes[:8] == [2, 3, 5, 7, 11, 13, 17, 19] and all(all(p % q for q in range(2, int(p ** 0.5) + 1)) for p in primes)
    for _ in range(20000):
        n, b = int(rng.integers(0, 2 ** 40)), primes[int(rng.integers(0, 40))]
        assert dev.devsrc_radical_inverse(n, b) == L.orc_radical_inverse(n, b)
    from pbrt_rust_b200 import scenes
    cfg = scenes.config1(xres=37, yres=23, sampler="halton")
    c = orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0)
    cs, _, counts = orc.halton_samples(c)
    lay = orc.layout(c)
    se, nt = lay["sample_ext"], lay["num_tasks"]
    ext = np.array(se, np.int32)
    seen = np.zeros_like(counts)
    out = np.zeros(2, np.float32)
    for t in range(nt):
        win = np.zeros(4, np.int32)
        L.orc_compute_sub_window(_p(ext), t, nt, _p(win))
        x0, x1, y0, y1 = (int(v) for v in win)
        if x0 == x1 or y0 == y1:
            continue
        dx, dy = x1 - x0, y1 - y0
        delta = np.float32(max(float(dy), float(dx)))
        for i in range(max(dx, dy) ** 2 * 4):
            if not dev.devsrc_halton_image(_p(win), delta, i, _p(out)):
                continue
            px = min(max(int(np.floor(out[0])), x0), x1 - 1) - se[0]
            py = min(max(int(np.floor(out[1])), y0), y1 - 1) - se[2]
            k = seen[py, px]
            assert k < counts[py, px] and cs[py, px, k, 0] == out[0] and cs[py, px, k, 1] == out[1]
            seen[py, px] += 1
    assert np.array_equal(seen, counts)


def _edge_rays(rng, n):
    """rays with zero / tiny / huge / NaN direction components, origins on box planes, empty ranges"""
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    k = rng.integers(0, 8, n)
    for a in range(3):
        d[(k == 1 + a) | (k == 6), a] = 0.0
    d[k == 4] *= np.float32(1e-30)
    d[k == 5] *= np.float32(1e30)
    d[k == 7, rng.integers(0, 3)] = np.nan
    m = rng.integers(0, 4, n) == 0
    o[m] = np.round(o[m])  # on integer planes
    mint = np.where(rng.integers(0, 4, n) == 0, rng.uniform(0, 4, n), 0.0).astype(np.float32)
    maxt = np.where(rng.integers(0, 3, n) == 0, rng.uniform(0, 8, n), 3.4028235e38).astype(np.float32)
    return np.concatenate([o, mint[:, None], d, maxt[:, None]], axis=1).astype(np.float32)


def test_device_slab_and_triangle_tests_match_the_oracle(dev, orc):
    """slab_test / slab_test_finite (bbox.rs:185-209) and tri_hit (mesh.rs:41-72) of
    csrc/trace_math.cuh against the oracle on random and degenerate inputs: decisions equal, entry
    distance equal, (t, b1, b2) bit for bit.  The min/max variant is only used when 1/d is finite on every
    axis (trace_ray), which is where it must equal the reference's swap form."""
    L = orc.lib()
    rng = np.random.default_rng(11)
    rays = np.concatenate([_edge_rays(rng, 6000), _edge_rays(np.random.default_rng(12), 6000)])
    n_fin = n_hit = n_tri = 0
    for r in rays:
        lo = np.round(rng.uniform(-3, 2, 3)) if rng.integers(3) == 0 else rng.uniform(-3, 2, 3)
        box = np.concatenate([lo, lo + rng.uniform(0, 3, 3) * (rng.integers(0, 5, 3) > 0)]).astype(np.float32)
        r = r.copy()
        if rng.integers(2) and np.isfinite(r[4:7]).all() and (r[4:7] != 0).all():  # aim half of the plain rays at the box
            r[4:7] = ((box[:3] + box[3:]) * 0.5 + rng.uniform(-0.6, 0.6, 3) - r[0:3]).astype(np.float32)
        t01, t0 = np.zeros(2, np.float32), np.zeros(1, np.float32)
        want = L.orc_bbox_intersect(_p(box), _p(r), _p(t01))
        got = dev.devsrc_slab(_p(box), _p(r), 0, _p(t0))
        assert got == want
        if want:
            # (the entry distance only feeds comparisons; max(-0.0, +0.0) may come out with either
            # sign depending on how fmax is lowered, so compare the value, not the bits)
            assert t0[0] == t01[0]
            n_hit += 1
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.float32(1.0) / r[4:7]
        if np.isfinite(inv).all():
            n_fin += 1
            got = dev.devsrc_slab(_p(box), _p(r), 1, _p(t0))
            assert got == want
            if want:
                assert t0[0] == t01[0]
        centre = (r[0:3] + r[4:7] * np.float32(rng.uniform(0.2, 2.0))) if rng.integers(2) and np.isfinite(r[4:7]).all() else rng.uniform(-3, 3, 3)
        p9 = (np.asarray(centre, np.float64)[None, :] + rng.uniform(-1.5, 1.5, (3, 3))).astype(np.float32).reshape(-1)
        if rng.integers(8) == 0:
            p9[6:9] = p9[0:3]  # degenerate triangle
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        want = L.orc_tri_intersect(_p(p9), _p(r), _p(a))
        got = dev.devsrc_tri_hit(_p(p9), _p(r), _p(b))
        assert got == want
        if want:
            n_tri += 1
            nan = np.isnan(a)
            assert np.array_equal(np.isnan(b), nan) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))
    assert n_fin > 3000 and n_hit > 1000 and n_tri > 1000, (n_fin, n_hit, n_tri)


def test_device_math_helpers_match_the_oracle(dev, orc):
    """quadratic_ (utils/mod.rs:56-92), solve2x2_ (:94-110), coordinate_system_ (vector.rs:184-195 as
    written) and chacha12_block (rand_chacha 0.3 ChaCha12) of csrc/dmath.cuh against the oracle."""
    L = orc.lib()
    rng = np.random.default_rng(21)
    for _ in range(5000):
        a, b, c = (np.float32(v) for v in rng.uniform(-4, 4, 3))
        if rng.integers(6) == 0:
            a = np.float32(0.0)
        if rng.integers(6) == 0:
            c = np.float32(b * b / (4 * a)) if a != 0 else c  # discriminant near zero
        w0, w1, g = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(2, np.float32)
        want = L.orc_quadratic(a, b, c, _p(w0), _p(w1))
        got = dev.devsrc_quadratic(a, b, c, _p(g))
        assert got == want
        if want:
            assert g.view(np.uint32).tolist() == [w0.view(np.uint32)[0], w1.view(np.uint32)[0]]
        a4, b2 = rng.uniform(-2, 2, 4).astype(np.float32), rng.uniform(-2, 2, 2).astype(np.float32)
        if rng.integers(5) == 0:
            a4[2:] = a4[:2] * np.float32(2.0)  # singular
        xw, xg = np.zeros(2, np.float32), np.zeros(2, np.float32)
        want = L.orc_solve_2x2(_p(a4), _p(b2), _p(xw))
        got = dev.devsrc_solve2x2(_p(a4), _p(b2), _p(xg))
        assert got == want and (not want or np.array_equal(xw.view(np.uint32), xg.view(np.uint32)))
        v = rng.uniform(-2, 2, 3).astype(np.float32)
        cw, cg, zero = np.zeros(6, np.float32), np.zeros(6, np.float32), np.zeros(3, np.float32)
        L.orc_vec_op(4, _p(v), _p(zero), _p(cw))
        dev.devsrc_coordinate_system(_p(v), _p(cg))
        assert np.array_equal(cw.view(np.uint32), cg.view(np.uint32))
    for seed in (0, 1, 12, 255):
        key = np.zeros(8, np.uint32)
        L.orc_seed_from_u64(C.c_uint64(seed), _p(key))
        for blk in (0, 1, 7, 2 ** 32 + 5):
            want, got = np.zeros(16, np.uint32), np.zeros(16, np.uint32)
            L.orc_stream_words(_p(key), C.c_uint64(16 * blk), C.c_uint64(16), _p(want))
            dev.devsrc_chacha12_block(_p(key), blk, _p(got))
            assert np.array_equal(want, got)


def _unit(rng):
    v = rng.normal(size=3)
    return (v / np.linalg.norm(v)).astype(np.float32)


def test_device_bsdf_differentials_and_segments_match_the_oracle(dev, orc):
    """BSDF::f over Lambertian / Oren-Nayar / plastic (bsdf/*.rs), Fresnel::Dielectric, DifferentialGeometry::
    compute_differentials (diff_geom.rs:81-152) and VisibilityTester::segment of csrc/shade_math.cuh
    against the oracle, bit for bit (both sides call the same libm here)."""
    L = orc.lib()
    L.orc_fresnel_dielectric.restype = C.c_float
    L.orc_fresnel_dielectric.argtypes = [C.c_float] * 3
    L.orc_bsdf_f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_vis_segment.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
    rng = np.random.default_rng(31)
    nonblack = 0
    for _ in range(6000):
        kind = int(rng.integers(0, 3))
        kd, ks = rng.uniform(0, 1, 3).astype(np.float32), rng.uniform(0, 1, 3).astype(np.float32)
        par = np.float32(rng.uniform(0.5, 60.0) if kind == 1 else rng.uniform(0.0005, 0.6))
        nn = _unit(rng)
        ng = nn if rng.integers(3) else _unit(rng)
        dpdu = (np.cross(nn, _unit(rng)) * rng.uniform(0.2, 3.0)).astype(np.float32)
        frame = np.concatenate([nn, ng, dpdu]).astype(np.float32)
        wo = _unit(rng)
        wi = (_unit(rng) * np.float32(rng.uniform(0.5, 8.0) if rng.integers(2) else 1.0)).astype(np.float32)  # point.rs: wi un-normalised
        if rng.integers(10) == 0:
            wi = (nn * np.float32(2.0)).astype(np.float32)   # along the normal: sin_theta = 0 paths
        want, got = np.zeros(3, np.float32), np.zeros(3, np.float32)
        strict = int(rng.integers(8) == 0)
        L.orc_bsdf_f(kind, _p(kd), _p(ks), par, _p(frame), _p(wo), _p(wi), strict, _p(want))
        dev.devsrc_bsdf_f(kind, _p(kd), _p(ks), par, _p(frame), _p(wo), _p(wi), strict, _p(got))
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan) and np.array_equal(want[~nan].view(np.uint32), got[~nan].view(np.uint32)), (kind, want, got)
        nonblack += bool((want != 0).any())
        c, ei, et = np.float32(rng.uniform(-1.2, 1.2)), np.float32(rng.uniform(1.0, 2.0)), np.float32(rng.uniform(1.0, 2.0))
        assert dev.devsrc_fresnel_dielectric(c, ei, et) == L.orc_fresnel_dielectric(c, ei, et)
        g12 = np.concatenate([rng.uniform(-3, 3, 3), nn, dpdu, np.cross(nn, dpdu) * rng.uniform(0.3, 2.0)]).astype(np.float32)
        rd12 = np.concatenate([rng.uniform(-5, 5, 6), _unit(rng), _unit(rng)]).astype(np.float32)
        if rng.integers(10) == 0:
            rd12[6:9] = np.cross(nn, _unit(rng))   # differential ray parallel to the tangent plane
        dw, dg_ = np.zeros(10, np.float32), np.zeros(10, np.float32)
        L.orc_compute_differentials(_p(g12), _p(rd12), _p(dw))
        dev.devsrc_compute_differentials(_p(g12), _p(rd12), _p(dg_))
        nan = np.isnan(dw)
        assert np.array_equal(np.isnan(dg_), nan) and np.array_equal(dw[~nan].view(np.uint32), dg_[~nan].view(np.uint32))
        p1, p2 = rng.uniform(-5, 5, 3).astype(np.float32), rng.uniform(-5, 5, 3).astype(np.float32)
        e1, e2 = np.float32(rng.uniform(0, 1e-2)), np.float32(rng.choice([0.0, 1e-3]))
        rw, rg = np.zeros(8, np.float32), np.zeros(8, np.float32)
        L.orc_vis_segment(_p(p1), e1, _p(p2), e2, _p(rw))
        dev.devsrc_vis_segment(_p(p1), e1, _p(p2), e2, _p(rg))
        assert np.array_equal(rw.view(np.uint32), rg.view(np.uint32))
    assert nonblack > 2000


def test_device_quadric_dg_matches_the_oracle(dev, orc):
    """The dg of Sphere / Cylinder / Disk hits (sphere.rs:143-180, cylinder.rs:127-154, disk.rs:107-133,
    helpers.rs:17-43) of csrc/shade_math.cuh, fed with the host mirror's flattened records, against
    the oracle's Shape::intersect: p, nn, (u, v), dpdu, dpdv bit for bit."""
    from pbrt_rust_b200.api import Light
    L = orc.lib()
    L.orc_quadric_intersect.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(41)
    U = lambda a, b: float(rng.uniform(a, b))
    hits = {0: 0, 1: 0, 2: 0}
    for _ in range(300):
        t = Transform.translate((U(-2, 2), U(-2, 2), U(-2, 2))) * Transform.rotate_x(U(0, 360)) * Transform.rotate_y(U(0, 360)) \
            * Transform.scale(U(0.5, 1.5), U(0.5, 1.5), U(0.5, 1.5) * (-1.0 if rng.integers(4) == 0 else 1.0))
        ro, pm = bool(rng.integers(2)), 360.0 if rng.integers(2) else U(90, 350)
        shape = int(rng.integers(3))
        if shape == 0:
            prm = (U(0.4, 1.2), U(-0.9, -0.2), U(0.2, 0.9))
            prm = (prm[0], prm[1] * prm[0], prm[2] * prm[0])
            sh = Shape.sphere(t, t.inverse(), ro, prm[0], prm[1], prm[2], pm)
        elif shape == 1:
            prm = (U(0.3, 1.0), U(-1.0, 0.0), U(0.1, 1.0))
            sh = Shape.cylinder(t, t.inverse(), ro, prm[0], prm[1], prm[2], pm)
        else:
            prm = (U(-0.3, 0.3), U(0.6, 1.2), U(0.0, 0.3))
            sh = Shape.disk(t, t.inverse(), ro, prm[0], prm[1], prm[2], pm)
        m = Material.matte(Texture.constant(0.5), Texture.constant(0.0))
        hs = HostScene(Scene.new_with(Primitive.bvh([Primitive.geometric(sh, m)], 1, "sah"), []))
        f = hs.flat.contents
        rec = C.cast(f.spheres, C.c_void_p)
        o2w = C.cast(f.sphere_o2w, C.c_void_p)
        centre = np.asarray(t.m, np.float32).reshape(4, 4)[:3, 3]
        for _ in range(12):
            o = (centre + _unit(rng) * np.float32(U(2.5, 5.0))).astype(np.float32)
            d = (centre + rng.uniform(-0.5, 0.5, 3) - o).astype(np.float32)
            ray = np.concatenate([o, [0.0], d, [3.4028235e38]]).astype(np.float32)
            o3, want, props = np.zeros(3, np.float32), np.zeros(14, np.float32), np.zeros(13, np.float32)
            m_, mi_ = np.asarray(t.m, np.float32), np.asarray(t.m_inv, np.float32)
            if shape == 0:
                ok = L.orc_sphere_intersect(_p(m_), _p(mi_), int(ro), C.c_float(prm[0]), C.c_float(prm[1]), C.c_float(prm[2]),
                                            C.c_float(pm), _p(ray), _p(o3), _p(want))
            else:
                ok = L.orc_quadric_intersect(shape, _p(m_), _p(mi_), int(ro), prm[0], prm[1], prm[2], pm, _p(ray), _p(o3),
                                             _p(want), _p(props))
            if not ok:
                continue
            got = np.zeros(14, np.float32)
            dev.devsrc_quadric_dg(rec, o2w, _p(o), _p(d), o3[0], o3[2], _p(got))
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (shape, want, got)
            hits[shape] += 1
    assert min(hits.values()) > 100, hits


def test_device_sampler_helpers_match_the_oracle(dev, orc):
    """van_der_corput / sobol2 (sampler/utils.rs:6-35, as written), RNG::random_float and RNG::shuffle
    (rng.rs:15-33) over the ChaCha12 word stream of csrc/dmath.cuh against the oracle."""
    L = orc.lib()
    L.orc_van_der_corput.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_sobol2.argtypes = [C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(51)
    for _ in range(20000):
        n, sc = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32))
        assert dev.devsrc_van_der_corput(n, sc) == L.orc_van_der_corput(n, sc)
        assert dev.devsrc_sobol2(n, sc) == L.orc_sobol2(n, sc)
    for seed in (0, 3, 12, 4095):
        key = np.zeros(8, np.uint32)
        L.orc_seed_from_u64(C.c_uint64(seed), _p(key))
        want, got = np.zeros(1000, np.float32), np.zeros(1000, np.float32)
        L.orc_rng_floats(C.c_uint64(seed), C.c_uint64(1000), _p(want))
        dev.devsrc_stream_floats(_p(key), 0, 1000, _p(got))
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
        for dims, count in ((1, 1), (1, 17), (2, 16), (2, 63)):
            v = rng.uniform(0, 1, count * dims).astype(np.float32)
            a, b = v.copy(), v.copy()
            L.orc_rng_shuffle(C.c_uint64(seed), _p(a), C.c_uint64(count * dims), C.c_uint64(dims))
            dev.devsrc_shuffle(_p(key), 0, _p(b), count, dims)
            assert np.array_equal(a, b) and sorted(a.tolist()) == sorted(v.tolist())


def test_device_film_arithmetic_matches_the_oracle(dev, orc):
    """Film::add_sample's extent and filter-table arithmetic (film.rs:192-249) of csrc/film_math.cuh,
    with the host mirror's film record and table, against the oracle's Film for every filter, crop
    windows and samples on / outside the film border: the weight every pixel receives, bit for bit."""
    from pbrt_rust_b200.api import Camera, Film, Filter, Sampler
    L = orc.lib()
    L.orc_film_add_sample.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(61)
    U = lambda a, b: float(rng.uniform(a, b))
    touched = 0
    for trial in range(60):
        filt = [Filter.mean(0.5, 0.5), Filter.mean(U(0.5, 2.5), U(0.5, 2.5)), Filter.triangle(U(1, 2.5), U(1, 2.5)),
                Filter.gaussian(U(1, 2.5), U(1, 2.5), U(0.5, 2)), Filter.mitchell(2.0, 2.0, 1 / 3, 1 / 3),
                Filter.lanczos(U(1.5, 3), U(1.5, 3), 3.0)][trial % 6]
        crop = (0, 1, 0, 1) if trial % 3 else tuple(sorted([U(0, 0.5), U(0.5, 1)]) + sorted([U(0, 0.5), U(0.5, 1)]))
        xres, yres = int(rng.integers(6, 30)), int(rng.integers(5, 24))
        film = Film.image(xres, yres, filt, crop)
        cam = Camera.perspective(Transform.new(), (-1.0, 1.0, -1.0, 1.0), 0.0, 0.0, 0.0, 1e6, 50.0, film)
        e = film.get_sample_extent()
        cfg = orc.render_config(cam, Sampler.stratified(e[0], e[1], e[2], e[3], 1, 1, False, 0.0, 0.0), num_cpus=8, mode=0)
        h, w = film.shape
        px = film.get_pixel_extent()
        for _ in range(40):
            k = int(rng.integers(4))
            sx = U(e[0] - 1.0, e[1] + 1.0) if k else float(rng.integers(px[0], px[1] + 1)) + (0.5 if rng.integers(2) else 0.0)
            sy = U(e[2] - 1.0, e[3] + 1.0) if k else float(rng.integers(px[2], px[3] + 1)) + (0.5 if rng.integers(2) else 0.0)
            want, got = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
            rc = L.orc_film_add_sample(C.byref(cfg), sx, sy, _p(want))
            assert rc == 0
            dev.devsrc_film_weights(C.byref(film.desc), sx, sy, _p(got))
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (trial, sx, sy)
            touched += int((want != 0).sum())
    assert touched > 5000


@pytest.mark.parametrize("sampler,lensr", [("stratified", 0.0), ("stratified", 0.15), ("ld", 0.05)])
def test_device_camera_ray_matches_the_oracle(dev, orc, sampler, lensr):
    """camera_ray of csrc/trace_math.cuh (camera/mod.rs:168-271 incl. depth of field, projective.rs:79-97)
    fed with the host mirror's camera record, against Camera::generate_ray of the oracle for every
    camera sample of a small frame: origin and direction bit for bit."""
    cfg = scenes.config1(xres=24, yres=18, sampler=sampler)
    cam = cfg["camera"]
    if lensr:
        from pbrt_rust_b200.api import Camera
        cam = Camera.perspective(cam.cam2world, cam.screen_window, 0.0, 0.0, lensr, 7.5, cam.fov, cam.film)
    ocfg = orc.render_config(cam, cfg["sampler"], num_cpus=8, mode=0)
    se = orc.layout(ocfg)["sample_ext"]
    spp = cfg["sampler"].samples_per_pixel()
    cs, rays, _, _ = orc.camera_samples(ocfg, 0, se[0], se[1], se[2], se[3], spp)
    assert cs.shape[0] == (se[1] - se[0]) * (se[3] - se[2]) * spp
    if lensr:
        assert np.abs(rays[:, 0:3] - rays[0, 0:3]).max() > 1e-3   # origins move over the lens
    out = np.zeros(6, np.float32)
    for k in range(cs.shape[0]):
        dev.devsrc_camera_ray(C.byref(cam.desc), spp, _p(cs[k]), _p(out))
        assert np.array_equal(out[0:3].view(np.uint32), rays[k, 0:3].view(np.uint32))
        assert np.array_equal(out[3:6].view(np.uint32), rays[k, 4:7].view(np.uint32))


def test_device_triangle_surface_matches_the_oracle(dev, orc):
    """tri_dg and tri_shading_geometry of csrc/shade_math.cuh (mesh.rs:105-193 as written — the (ss, ts)
    tuple binding, zeroed differentials —, :220-262), fed with the host mirror's flattened triangle,
    mesh and attribute records, against the oracle's Triangle for meshes with / without normals,
    tangents and uvs under random (also mirroring) transforms: geometric and shading dg bit for bit."""
    L = orc.lib()
    L.orc_tri_surface.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_float] * 3 + [C.c_void_p] * 2
    rng = np.random.default_rng(71)
    U = lambda a, b: float(rng.uniform(a, b))
    done = set()
    for trial in range(400):
        t = Transform.translate((U(-2, 2), U(-2, 2), U(-2, 2))) * Transform.rotate_x(U(0, 360)) * Transform.rotate_z(U(0, 360)) \
            * Transform.scale(U(0.5, 1.5), U(0.5, 1.5), U(0.5, 1.5) * (-1.0 if rng.integers(3) == 0 else 1.0))
        if rng.integers(4) == 0:
            t = Transform.new()
        ro = bool(rng.integers(2))
        P = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        has_n, has_s, has_uv = bool(rng.integers(2)), bool(rng.integers(2)), bool(rng.integers(2))
        N = (np.cross(P[1] - P[0], P[2] - P[0])[None, :] + rng.uniform(-0.3, 0.3, (3, 3))).astype(np.float32) if has_n else None
        S = ((P[1] - P[0])[None, :] + rng.uniform(-0.2, 0.2, (3, 3))).astype(np.float32) if has_s else None
        UV = rng.uniform(0, 1, (3, 2)).astype(np.float32) if has_uv else None
        if has_uv and rng.integers(6) == 0:
            UV[2] = UV[0]  # degenerate uv parameterisation -> coordinate_system branch
        mat = Material.matte(Texture.constant(0.5), Texture.constant(0.0))
        sh = Shape.triangle_mesh(t, t.inverse(), ro, np.arange(3, dtype=np.uint32), P, N, S, UV)
        hs = HostScene(Scene.new_with(Primitive.bvh([Primitive.geometric(sh, mat)], 1, "sah"), []))
        f = hs.flat.contents
        tri = f.tris[0]
        pw = np.array([*tri.p1, *tri.p2, *tri.p3], np.float32)
        attr = int(tri.attr)
        n9 = np.ctypeslib.as_array(f.tri_n, shape=(f.n_attr * 9,))[9 * attr:9 * attr + 9].copy() if has_n else None
        s9 = np.ctypeslib.as_array(f.tri_s, shape=(f.n_attr * 9,))[9 * attr:9 * attr + 9].copy() if has_s else None
        uv6 = np.ctypeslib.as_array(f.tri_uv, shape=(f.n_attr * 6,))[6 * attr:6 * attr + 6].copy() if has_uv else None
        rev = [2, 1, 0]  # the Triangle holds (P[vi[2]], P[vi[1]], P[vi[0]]) (mesh.rs:329-331)
        oP = np.ascontiguousarray(P[rev]).reshape(-1)
        oN = np.ascontiguousarray(N[rev]).reshape(-1) if has_n else None
        oS = np.ascontiguousarray(S[rev]).reshape(-1) if has_s else None
        oUV = np.ascontiguousarray(UV[rev]).reshape(-1) if has_uv else None
        m_, mi_ = np.asarray(t.m, np.float32), np.asarray(t.m_inv, np.float32)
        mesh_rec = C.byref(f.meshes[0])
        for _ in range(4):
            bb = rng.dirichlet([1, 1, 1])
            hit = bb[0] * pw[0:3] + bb[1] * pw[3:6] + bb[2] * pw[6:9]
            o = (hit + _unit(rng) * U(1.0, 4.0)).astype(np.float32)
            d = (hit - o).astype(np.float32) * np.float32(U(0.5, 2.0))
            ray = np.concatenate([o, [0.0], d, [3.4028235e38]]).astype(np.float32)
            tbb = np.zeros(3, np.float32)
            if not L.orc_tri_intersect(_p(pw), _p(ray), _p(tbb)):
                continue
            wa, wb, ga, gb = np.zeros(14, np.float32), np.zeros(17, np.float32), np.zeros(14, np.float32), np.zeros(17, np.float32)
            pn = lambda a: None if a is None else _p(a)
            rc = L.orc_tri_surface(_p(m_), _p(mi_), int(ro), _p(oP), pn(oN), pn(oS), pn(oUV), _p(ray), tbb[0], tbb[1], tbb[2],
                                   _p(wa), _p(wb))
            assert rc == 0
            dev.devsrc_tri_surface(mesh_rec, _p(pw), pn(n9), pn(s9), pn(uv6), _p(o), _p(d), tbb[0], tbb[1], tbb[2], _p(ga), _p(gb))
            for want, got in ((wa, ga), (wb, gb)):
                nan = np.isnan(want)
                assert np.array_equal(np.isnan(got), nan) and np.array_equal(want[~nan].view(np.uint32), got[~nan].view(np.uint32)), \
                    (has_n, has_s, has_uv, want, got)
            done.add((has_n, has_s, has_uv))
    assert len(done) == 8


def test_device_image_textures_match_the_oracle(dev, orc):
    """MIPMap::lookup (trilinear and EWA, all three wrap modes, spectrum and float maps; mipmap.rs:
    112-341) of csrc/shade_mip.cuh over the host mirror's pyramids (MIPMap::new, texel pool layout),
    reached through the general evaluator, against the oracle's ImageTexture — bit for bit — alone and
    nested inside checkerboard / mix / scale trees (scenes.TexGen with images)."""
    from oracle import orc as O
    from pbrt_rust_b200.api import PlanarMapping2D, UVMapping2D
    rng = np.random.default_rng(81)
    img = scenes.procedural_image(37, 23, seed=5)   # not a power of two: resized by MIPMap::new
    texs = []
    for tri in (True, False):
        for wrap in ("repeat", "black", "clamp"):
            for spectrum in (True, False):
                mp = UVMapping2D(float(rng.uniform(0.5, 3)), float(rng.uniform(0.5, 3)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1))) \
                    if rng.integers(2) else PlanarMapping2D(tuple(rng.uniform(-0.5, 0.5, 3)), tuple(rng.uniform(-0.5, 0.5, 3)), 0.1, 0.2)
                texs.append(Texture.image(mp, img, spectrum=spectrum, do_trilinear=tri, max_aniso=float(rng.choice([2.0, 8.0])),
                                          wrap=wrap, scale=float(rng.uniform(0.6, 1.0)), gamma=float(rng.uniform(1.0, 2.2))))
    gen = scenes.TexGen(np.random.default_rng(82), img, ext=True, images=True)
    nested = [gen.spectrum_tex() for _ in range(80)]
    texs += nested

    def has_image(t):
        return t.kind == 3 or any(has_image(c) for c in t.children())

    assert sum(has_image(t) for t in nested) >= 15
    scene = _scene_of([Material.matte(t, Texture.constant(0.0)) for t in texs])
    hs, osc = HostScene(scene), O.OracleScene(scene)
    f = hs.flat.contents
    assert f.n_mipmaps >= 12
    table, mips, texels = C.cast(f.textures, C.c_void_p), C.cast(f.mipmaps, C.c_void_p), C.cast(f.texels, C.c_void_p)
    for t in texs:
        hid, oid = hs.tex_ids[id(t)], osc.tex_ids[id(t)]
        for _ in range(30):
            g = _random_dg(rng)
            g[20:30] *= np.float32(0.05)   # modest footprints: the EWA ellipse stays a few texels wide
            got, want, q = np.zeros(3, np.float32), np.zeros(3, np.float32), _dg15(g)
            dev.devsrc_tex_eval_img(table, mips, texels, hid, _p(g), _p(got))
            O.lib().orc_texture_eval(osc.h, oid, _p(q), _p(want))
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(got), nan) and np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32)), (t.kind, got, want)


def _scene_rays(rng, n, lo, hi, spread):
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    tgt = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 4:7], rays[:, 7] = o, tgt - o, 3.4028235e38
    k = rng.integers(0, 6, n)
    rays[k == 0, 7] = rng.uniform(0.2, 1.5, int((k == 0).sum()))      # short segments (shadow-ray like)
    rays[k == 1, 3] = rng.uniform(0.0, 0.8, int((k == 1).sum()))      # mint > 0
    return rays


@pytest.mark.parametrize("box,mixed", [(1, 0), (2, 0), (3, 0), (2, 5), (3, 7), (3, 2)])
@pytest.mark.parametrize("kind", ["triangles_single", "triangles_multi", "mixed_quadrics", "random_scenes"])
def test_device_traversal_matches_the_oracle(dev, orc, kind, box, mixed):
    """The kernels' traversal source (csrc/trace_core.cuh: pair-node steps, shared-memory-style stack,
    leaf decoding, triangle / sphere / cylinder / disk tests, the finite and the NaN-faithful slab
    paths) run on the CPU over the host mirror's flattened scene with the product's own pair-node
    packing, against BVHAccelerator::intersect / intersect_p of the oracle: hit ids, t and the
    barycentrics bit for bit (phi of quadrics within the libm-free tolerance 0 here: same libm), the
    occlusion bit for all five any-hit loop shapes (ordered, unordered, postponed leaves).  `box` = the box-test family
    (trace_core.cuh child_box): exact min / max, exact per octant, and conservative FFMA inner tests with
    the reference's exact test applied in the leaf phase — all three must give the same bits.  `mixed`
    forces the per-axis "the warp mixes signs" form of the specialised tests (a single emulated lane
    never does by itself)."""
    from oracle import orc as O
    dev.devsrc_set_box(box)
    dev.devsrc_force_mixed_axes(mixed)
    rng = np.random.default_rng({"triangles_single": 1, "triangles_multi": 2, "mixed_quadrics": 3, "random_scenes": 4}[kind])
    if kind == "triangles_single":
        cfgs = [scenes.config3(nx=40, nz=20, xres=16, yres=16, xs=1, ys=1)]
    elif kind == "triangles_multi":
        cfgs = [scenes.config2(n=4000, xres=16, yres=16)]
    elif kind == "mixed_quadrics":
        cfgs = [scenes.config4(n_ground=(30, 15), n_spheres=200, xres=16, yres=16, xs=1, ys=1)]
    else:
        cfgs = [scenes.random_scene(seed) for seed in (3, 11, 29, 42, 57, 64, 71, 88, 93, 105, 117, 123)]
    for cfg in cfgs:
        hs, osc = HostScene(cfg["scene"]), O.OracleScene(cfg["scene"])
        f = hs.flat.contents
        lo = np.array([f.nodes[0].bmin[i] for i in range(3)]) - 3.0
        hi = np.array([f.nodes[0].bmax[i] for i in range(3)]) + 3.0
        rays = np.concatenate([_scene_rays(rng, 3000, lo, hi, float(max(abs(lo).max(), abs(hi).max())) * 0.7),
                               _edge_rays(rng, 600) * np.float32([4, 4, 4, 1, 1, 1, 1, 1])]).astype(np.float32)
        prim, tbb, _ = osc.trace_closest(rays)
        got = np.zeros((rays.shape[0], 4), np.float32)
        assert dev.devsrc_trace(C.byref(f), _p(rays), rays.shape[0], -1, _p(got)) == 0
        gprim = got[:, 0].copy().view(np.uint32)
        hit = prim != 0xFFFFFFFF
        assert hit.mean() > 0.1
        # A ray with a NaN direction component "hits" every triangle it tests with t = NaN (all of the
        # test's comparisons are false), and a NaN maxt no longer bounds later box tests, so the
        # reference may go on to subtrees the device pruned while maxt was still finite: both sides
        # report a NaN hit, possibly at different primitives (float-edge cases).
        nan_dir = np.isnan(rays[:, 4:7]).any(axis=1)
        assert np.array_equal(gprim[~nan_dir], prim[~nan_dir])
        assert np.array_equal(gprim[nan_dir] != 0xFFFFFFFF, hit[nan_dir])
        nan = np.isnan(tbb[:, 0])
        assert np.array_equal(np.isnan(got[hit, 1]), nan[hit])
        ok = hit & ~nan
        assert np.array_equal(got[ok, 1].view(np.uint32), tbb[ok, 0].view(np.uint32))
        assert np.array_equal(got[ok, 2:4].view(np.uint32), tbb[ok, 1:3].view(np.uint32))
        occ, _ = osc.trace_any(rays)
        for mode in range(5):
            g = np.zeros((rays.shape[0], 4), np.float32)
            assert dev.devsrc_trace(C.byref(f), _p(rays), rays.shape[0], mode, _p(g)) == 0
            assert np.array_equal((g[:, 0].copy().view(np.uint32) != 0xFFFFFFFF).astype(np.uint8), occ), mode
    dev.devsrc_set_box(3)
    dev.devsrc_force_mixed_axes(0)


@pytest.mark.parametrize("sampler", ["stratified", "ld"])
def test_device_film_gather_matches_the_oracle(dev, orc, sampler):
    """film_pixel (csrc/film.cuh: the per-pixel gather over neighbouring sampler pixels in raster
    order, add_sample's arithmetic, the radiance fold and to_xyz) run for every film pixel of a frame
    on the oracle's own camera samples and random radiance, against a sequential Film::add_sample
    sweep of the oracle in raster order: xyz sums and weight sums bit for bit, for every filter and
    for crop windows (the deterministic-accumulation contract of DESIGN.md §4)."""
    from pbrt_rust_b200.api import Camera, Film, Filter, Sampler
    L = orc.lib()
    L.orc_film_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    rng = np.random.default_rng(91)
    U = lambda a, b: float(rng.uniform(a, b))
    for trial in range(12):
        filt = [Filter.mean(0.5, 0.5), Filter.mean(U(0.5, 2.0), U(0.5, 2.0)), Filter.triangle(U(1, 2), U(1, 2)),
                Filter.gaussian(U(1, 2.5), U(1, 2.5), U(0.5, 2)), Filter.mitchell(2.0, 2.0, 1 / 3, 1 / 3),
                Filter.lanczos(U(1.5, 3), U(1.5, 3), 3.0)][trial % 6]
        crop = (0, 1, 0, 1) if trial < 6 else tuple(sorted([U(0, 0.4), U(0.6, 1)]) + sorted([U(0, 0.4), U(0.6, 1)]))
        film = Film.image(int(rng.integers(12, 28)), int(rng.integers(10, 22)), filt, crop)
        cam = Camera.perspective(Transform.new(), (-1.0, 1.0, -1.0, 1.0), 0.0, 0.0, 0.0, 1e6, 50.0, film)
        e = film.get_sample_extent()
        smp = Sampler.stratified(e[0], e[1], e[2], e[3], 2, 2, True, 0.0, 0.0) if sampler == "stratified" \
            else Sampler.low_discrepancy(e[0], e[1], e[2], e[3], 4, 0.0, 0.0)
        cfg = orc.render_config(cam, smp, num_cpus=8, mode=0)
        cs, _, _, _ = orc.camera_samples(cfg, 0, e[0], e[1], e[2], e[3], 4)   # raster order, 4 per pixel
        img = np.ascontiguousarray(cs[:, 0:2])
        rgb = rng.uniform(0, 3, (cs.shape[0], 3)).astype(np.float32)
        h, w = film.shape
        want, got = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
        assert L.orc_film_accumulate(C.byref(cfg), _p(img), _p(rgb), cs.shape[0], _p(want)) == 0
        ext = np.array(e, np.int32)
        dev.devsrc_film_gather(C.byref(film.desc), _p(ext), 4, _p(img), _p(rgb), _p(got))
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (trial, float(np.abs(want - got).max()))
        assert (want[..., 3] != 0).mean() > 0.9  # (negative lobes of mitchell / lanczos are fine)


@pytest.mark.parametrize("which", ["config1", "config2", "config4"])
def test_device_primary_hits_of_a_frame_match_the_oracle(dev, orc, which):
    """The config-2 check on the CPU: every camera sample of a small frame through the device source's
    camera_ray and traversal (host mirror's flattened scene, product pair-node packing) against the
    oracle's render — primary-hit ids 100 % equal, t bit for bit."""
    from oracle import orc as O
    cfg = {"config1": lambda: scenes.config1(xres=40, yres=30),
           "config2": lambda: scenes.config2(n=5000, xres=48, yres=32),
           "config4": lambda: scenes.config4(n_ground=(30, 15), n_spheres=150, xres=40, yres=24, xs=2, ys=1)}[which]()
    hs, osc = HostScene(cfg["scene"]), O.OracleScene(cfg["scene"])
    ocfg = O.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0, primary_only=True)
    ref = O.render(osc, ocfg, want_hits=True)
    se = O.layout(ocfg)["sample_ext"]
    spp = cfg["sampler"].samples_per_pixel()
    # (the light-sample floats of area lights sit in the same per-pixel stream: same pair count as the render)
    pairs = sum(l.num_samples for l in cfg["scene"].all_lights() if l.kind == "area")
    cs, _, _, _ = O.camera_samples(ocfg, pairs, se[0], se[1], se[2], se[3], spp)
    rays = np.zeros((cs.shape[0], 8), np.float32)
    out = np.zeros(6, np.float32)
    for k in range(cs.shape[0]):
        dev.devsrc_camera_ray(C.byref(cfg["camera"].desc), spp, _p(cs[k]), _p(out))
        rays[k, 0:3], rays[k, 4:7] = out[0:3], out[3:6]
    rays[:, 7] = 3.4028235e38
    got = np.zeros((rays.shape[0], 4), np.float32)
    assert dev.devsrc_trace(C.byref(hs.flat.contents), _p(rays), rays.shape[0], -1, _p(got)) == 0
    prim = got[:, 0].copy().view(np.uint32)
    assert np.array_equal(prim, ref["hit_ids"])
    hit = prim != 0xFFFFFFFF
    assert 0.1 < hit.mean()
    assert np.array_equal(got[hit, 1].view(np.uint32), ref["hit_ts"][hit].view(np.uint32))


def _device_source_frame(dev, orc, cfg, strict_flags=False):
    """Renders cfg through the device source on the CPU (oracle camera samples in, film out) and
    with the oracle; returns (film_dev, oracle result, stats3)."""
    from oracle import orc as O
    hs, osc = HostScene(cfg["scene"]), O.OracleScene(cfg["scene"])
    ocfg = O.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0, count_traversal=True)
    ref = O.render(osc, ocfg, want_hits=True, strict_flags=strict_flags)
    se = O.layout(ocfg)["sample_ext"]
    spp = cfg["sampler"].samples_per_pixel()
    pairs = sum(l.num_samples for l in cfg["scene"].all_lights() if l.kind == "area")
    cs, _, _, lu = O.camera_samples(ocfg, pairs, se[0], se[1], se[2], se[3], spp)
    img, lens = np.ascontiguousarray(cs[:, 0:2]), np.ascontiguousarray(cs[:, 2:4])
    h, w = cfg["film"].shape
    film = np.zeros((h, w, 4), np.float32)
    stats = np.zeros(3, np.uint64)
    ext = np.array(se, np.int32)
    rc = dev.devsrc_render(C.byref(hs.flat.contents), C.byref(cfg["camera"].desc), C.byref(cfg["film"].desc), _p(ext), spp,
                           _p(img), _p(lens), _p(lu) if pairs else None, int(strict_flags), _p(film), _p(stats))
    assert rc == 0, rc
    return film, ref, stats


@pytest.mark.parametrize("which", ["config1", "config1_ld_dof", "config3", "config4", "quadrics_wide_filter"])
def test_device_source_renders_whole_frames_like_the_oracle(dev, orc, which):
    """A complete frame through the device source on the CPU — camera_ray, the closest-hit traversal,
    the k_shade kernel itself (run one emulated thread per block: dg, shading geometry, textures,
    BSDFs, light sampling, the shadow-ray queue), the any-hit pass, film_pixel — against the oracle's
    render of the same scene: films bit for bit (xyz and weight sums), hit and shadow-ray counts equal.
    With both sides on the same libm this is exact; the GPU suite allows the CUDA-libm tolerance."""
    from pbrt_rust_b200.api import Camera, Filter, Film
    if which == "config1":
        cfg = scenes.config1(xres=48, yres=36)
    elif which == "config1_ld_dof":
        cfg = scenes.config1(xres=40, yres=30, sampler="ld")
        cam = cfg["camera"]
        cfg["camera"] = Camera.perspective(cam.cam2world, cam.screen_window, 0.0, 0.0, 0.1, 7.0, cam.fov, cam.film)
    elif which == "config3":
        cfg = scenes.config3(nx=40, nz=20, xres=48, yres=32, xs=2, ys=2)
    elif which == "config4":
        cfg = scenes.config4(n_ground=(30, 15), n_spheres=120, xres=48, yres=28, xs=2, ys=1)
    else:
        cfg = scenes.config4(n_ground=(20, 10), n_spheres=60, xres=36, yres=24, xs=1, ys=2, filt=Filter.gaussian(2.0, 1.5, 1.0)) \
            if "filt" in scenes.config4.__code__.co_varnames else scenes.config1(xres=36, yres=24, filt=Filter.gaussian(2.0, 1.5, 1.0))
    film, ref, stats = _device_source_frame(dev, orc, cfg)
    assert np.array_equal(film[..., 3].view(np.uint32), ref["film"][..., 3].view(np.uint32))
    assert np.array_equal(film.view(np.uint32), ref["film"].view(np.uint32)), float(np.abs(film - ref["film"]).max())
    assert int(stats[0]) == ref["stats"]["camera_hits"] and int(stats[1]) == ref["stats"]["shadow_rays"]
    assert film[..., :3].max() > 0


def test_device_source_renders_random_scenes_like_the_oracle(dev, orc):
    """The same whole-frame comparison over the differential fuzzer's scenes (base and extended
    generator: all shapes, materials, every texture kind incl. image maps, bump maps, all light kinds,
    filters, crops, depth of field; stratified and LD samplers)."""
    n = 0
    for ext in (False, True):
        for seed in range(14):
            cfg = scenes.random_scene(seed, ext=ext)
            if cfg["sampler"].kind == 2:
                continue  # Halton frames take the kernels' own sample generation (GPU suite)
            film, ref, stats = _device_source_frame(dev, orc, cfg)
            assert np.array_equal(film.view(np.uint32), ref["film"].view(np.uint32)), (ext, seed, float(np.abs(film - ref["film"]).max()))
            assert int(stats[0]) == ref["stats"]["camera_hits"] and int(stats[1]) == ref["stats"]["shadow_rays"]
            n += 1
    assert n >= 20


def _sampler_desc(smp, num_tasks):
    return _ffi.Sampler(smp.kind, smp.ext[0], smp.ext[1], smp.ext[2], smp.ext[3], smp.xs, smp.ys, int(smp.jitter),
                        smp.sopen, smp.sclose, num_tasks)


@pytest.mark.parametrize("kind,xs,ys,jitter,pairs", [("stratified", 2, 2, True, 0), ("stratified", 3, 2, True, 2),
                                                     ("stratified", 4, 4, True, 1), ("stratified", 2, 1, False, 0),
                                                     ("ld", 4, 1, True, 0), ("ld", 6, 1, True, 3)])
def test_device_raygen_kernels_match_the_oracle(dev, orc, kind, xs, ys, jitter, pairs):
    """k_raygen_groups and k_raygen_full, emulated thread by thread over a whole sampler extent with
    the reference's task windows and per-task StdRng keys, against the oracle's get_more_samples:
    image / lens / time samples and the light-sample floats bit for bit (stratified incl. the
    lens / time shuffles, unjittered, LD with its scrambles and shuffles)."""
    from pbrt_rust_b200.api import Sampler
    cfg = scenes.config1(xres=37, yres=22)
    e = cfg["sampler"].ext
    smp = Sampler.stratified(e[0], e[1], e[2], e[3], xs, ys, jitter, 0.25, 0.75) if kind == "stratified" \
        else Sampler.low_discrepancy(e[0], e[1], e[2], e[3], xs, 0.25, 0.75)
    ocfg = orc.render_config(cfg["camera"], smp, num_cpus=8, mode=0)
    ocfg.sopen, ocfg.sclose = 0.25, 0.75
    lay = orc.layout(ocfg)
    spp = smp.samples_per_pixel()
    cs, _, _, lu = orc.camera_samples(ocfg, pairs, e[0], e[1], e[2], e[3], spp)
    desc = _sampler_desc(smp, lay["num_tasks"])
    npx = (e[1] - e[0]) * (e[3] - e[2])
    for full in ([1] if kind == "ld" else [0, 1]):
        got = np.zeros((npx * spp, 5), np.float32)
        glu = np.zeros((npx * spp, max(1, 2 * pairs)), np.float32)
        edge = np.zeros(npx, np.uint32)
        dev.devsrc_raygen(C.byref(desc), pairs, C.byref(cfg["film"].desc), full, _p(got), _p(glu), _p(edge))
        cols = slice(0, 5) if full else slice(0, 2)
        assert np.array_equal(got[:, cols].view(np.uint32), cs[:, cols].view(np.uint32)), (full,)
        if pairs:
            assert np.array_equal(glu.view(np.uint32), lu.view(np.uint32))
        # box filter of half a pixel: a sample leaves its own pixel only on exact pixel borders
        assert edge.mean() < 0.2


@pytest.mark.parametrize("spp,pairs,res", [(4, 0, (33, 21)), (7, 2, (28, 19))])
def test_device_halton_kernels_match_the_oracle(dev, orc, spp, pairs, res):
    """k_halton_bin<0>, the host scan, k_halton_bin<1> (run in REVERSE candidate order here, so that
    only the per-pixel sort can put the samples back into generation order) and k_halton_samples,
    emulated thread by thread, against the oracle's HaltonSampler: per-pixel counts, every camera
    sample and the light-sample floats bit for bit."""
    from pbrt_rust_b200.api import Sampler
    cfg = scenes.config1(xres=res[0], yres=res[1])
    e = cfg["sampler"].ext
    smp = Sampler.halton(e[0], e[1], e[2], e[3], spp, 0.1, 0.9)
    ocfg = orc.render_config(cfg["camera"], smp, num_cpus=8, mode=0)
    ocfg.sopen, ocfg.sclose = 0.1, 0.9
    cs, lu, counts = orc.halton_samples(ocfg, pairs)
    desc = _sampler_desc(smp, orc.layout(ocfg)["num_tasks"])
    total = int(counts.sum())
    gc = np.zeros(counts.size, np.uint32)
    got = np.zeros((total, 5), np.float32)
    glu = np.zeros((total, max(1, 2 * pairs)), np.float32)
    assert dev.devsrc_halton(C.byref(desc), pairs, _p(gc), _p(got), _p(glu), total) == total
    assert np.array_equal(gc.reshape(counts.shape), counts)
    valid = ~np.isnan(cs[..., 0])
    assert np.array_equal(got.view(np.uint32), cs[valid].view(np.uint32))     # compact, pixel raster then generation order
    if pairs:
        assert np.array_equal(glu.view(np.uint32), lu[valid].view(np.uint32))
