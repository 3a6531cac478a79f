"""Pins the CPU oracle against every known-answer test the reference's own unit tests hold for the
hot path.  Each test names the reference test it transcribes (file:line)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

F32_MAX = float(np.finfo(np.float32).max)
f32 = np.float32


def _p(a):
    return a.ctypes.data_as(C.c_void_p)  # keeps the array alive for the duration of the call


def _ray(o, d, mint=0.0, maxt=F32_MAX):
    return np.array([*o, mint, *d, maxt], np.float32)


def _ident():
    return np.eye(4, dtype=np.float32)


# ---- utils/mod.rs ---------------------------------------------------------------------------
def test_quadratic(orc):
    """utils/mod.rs:366-388 it_can_solve_quadratic_equations"""
    L = orc.lib()

    def q(a, b, c):
        t0, t1 = C.c_float(), C.c_float()
        ok = L.orc_quadratic(a, b, c, C.byref(t0), C.byref(t1))
        return (t0.value, t1.value) if ok else None

    assert q(1.0, 0.0, 1.0) is None
    assert q(-1.0, 4.0, -4.0) == (2.0, 2.0)
    assert q(1.0, 0.0, 0.0) == (0.0, 0.0)
    assert q(1.0, -2.0, 0.0) == (0.0, 2.0)
    for i in range(2, 200):
        assert q(1.0, -float(i), 0.0) == (0.0, float(i))
    t0 = float(f32(np.sqrt(f32(3.0))) + f32(4.0))
    t1 = float(-f32(np.sqrt(f32(3.0))) + f32(4.0))
    assert q(-1.0, 8.0, -13.0) == (t1, t0)
    assert q(0.0, 1.0, 1.0) is None
    assert q(0.0, 0.0, 1.75) is None


def test_solve_linear_system_2x2(orc):
    """utils/mod.rs:392-437"""
    L = orc.lib()

    def s(a, b):
        x = np.zeros(2, np.float32)
        ok = L.orc_solve_2x2(_p(np.array(a, np.float32).reshape(-1)), _p(np.array(b, np.float32)), _p(x))
        return tuple(x) if ok else None

    nan = float("nan")
    assert s([[1, 0], [0, 1]], [3.5, -4.2]) == (f32(3.5), f32(-4.2))
    for a in ([[nan, 0], [0, 1]], [[1, nan], [0, 1]], [[1, 0], [nan, 1]], [[1, 0], [0, nan]]):
        assert s(a, [3.5, -4.2]) is None
    assert s([[1, 0], [0, 1]], [nan, -4.2]) is None
    assert s([[1, 0], [0, 1]], [3.5, nan]) is None
    for a in ([[0, 0], [0, 0]], [[1, 2], [1, 2]], [[3, 2], [6, 4]]):
        assert s(a, [3.5, -4.2]) is None
    h = 0.5 * math.sqrt(2.0)
    ox, oy = s([[h, -h], [h, h]], [3.0, 1.0])
    assert abs(ox - 2.0 * math.sqrt(2.0)) < 1e-6 and abs(oy + math.sqrt(2.0)) < 1e-6


def test_partition_by(orc):
    """utils/mod.rs:439-517 it_can_partition_slices"""
    L = orc.lib()

    def part(xs):
        a = np.array(xs, np.int32)
        L.orc_partition_by_i32(_p(a), C.c_uint64(len(a)))
        return a.tolist()

    def halves_ok(xs):
        m = len(xs) // 2
        return max(xs[:m]) <= min(xs[m:])

    assert halves_ok(part([2, 5, 6, 1, 0, 4, 3, 2, 5, 1, 3, 3, 5]))
    assert part([1, 2]) == [1, 2] and part([2, 1]) == [1, 2] and part([1]) == [1]
    assert halves_ok(part([1, 0, 0, 0, -1, 0]))
    assert part([3, 1, 1]) == [1, 1, 3]
    assert part([3, 1, 2]) == [1, 2, 3]
    assert part([3, 3, 1]) == [1, 3, 3]
    assert part([2, 2, 2]) == [2, 2, 2]
    assert halves_ok(part([-2, -4, 2, 2, 2, 4, 4, 4, 4, 4, 4, 4]))
    assert halves_ok(part([-2, -4, 2, 2, 2, 4, 4, 4, 4, 4, 4, -5]))


def test_get_crop_window(orc):
    """utils/mod.rs:520-529"""
    L = orc.lib()

    def cw(num, count, aspect):
        o = np.zeros(4, np.float32)
        L.orc_get_crop_window(num, count, aspect, _p(o))
        return tuple(float(x) for x in o)

    assert cw(0, 2, 2.0) == (0.0, 0.5, 0.0, 1.0)
    assert cw(1, 2, 2.0) == (0.5, 1.0, 0.0, 1.0)
    assert cw(0, 2, 0.5) == (0.0, 1.0, 0.0, 0.5)
    assert cw(1, 2, 0.5) == (0.0, 1.0, 0.5, 1.0)
    assert cw(0, 4, 1.0) == (0.0, 0.5, 0.0, 0.5)
    assert cw(1, 4, 1.0) == (0.5, 1.0, 0.0, 0.5)
    assert cw(2, 4, 1.0) == (0.0, 0.5, 0.5, 1.0)
    assert cw(3, 4, 1.0) == (0.5, 1.0, 0.5, 1.0)


def test_sampler_sub_windows(orc):
    """sampler/base.rs:68-89 its_base_can_tile_windows"""
    L = orc.lib()

    def sw(ext, num, count):
        o = np.zeros(4, np.int32)
        L.orc_compute_sub_window(_p(np.array(ext, np.int32)), C.c_uint64(num), C.c_uint64(count), _p(o))
        return tuple(int(x) for x in o)

    e = (0, 10, 0, 2)
    assert sw(e, 0, 20) == (0, 1, 0, 1)
    assert sw(e, 9, 20) == (9, 10, 0, 1)
    assert sw(e, 10, 20) == (0, 1, 1, 2)
    assert sw(e, 19, 20) == (9, 10, 1, 2)
    assert sw(e, 4, 5) == (8, 10, 0, 2)
    assert sw(e, 0, 1) == (0, 10, 0, 2)
    e = (0, 2, 0, 10)
    assert sw(e, 0, 20) == (0, 1, 0, 1)
    assert sw(e, 9, 20) == (1, 2, 4, 5)
    assert sw(e, 10, 20) == (0, 1, 5, 6)
    assert sw(e, 19, 20) == (1, 2, 9, 10)
    assert sw(e, 4, 5) == (0, 2, 8, 10)
    assert sw(e, 0, 1) == (0, 2, 0, 10)


def test_num_tasks_as_written(orc):
    """sampler_renderer.rs:41-44: 1080p -> 13, 4K -> 15"""
    L = orc.lib()
    assert L.orc_num_tasks_for(8, 1920 * 1080) == 13
    assert L.orc_num_tasks_for(8, 3840 * 2160) == 15
    assert L.orc_num_tasks_for(8, 640 * 480) == 11   # max(256, 1200) = 1200 -> ceil(log2) = 11
    assert L.orc_num_tasks_for(8, 16) == 8           # 256 = 2^8 exactly


# ---- bbox.rs --------------------------------------------------------------------------------
def _bbox_isect(orc, box, ray):
    t = np.zeros(2, np.float32)
    ok = orc.lib().orc_bbox_intersect(_p(np.array(box, np.float32)), _p(ray), _p(t))
    return (float(t[0]), float(t[1])) if ok else None


def test_bbox_intersect(orc):
    """bbox.rs:553-615 it_can_be_intersected"""
    b = [-1, -1, -1, 1, 1, 1]
    for axis in range(3):
        o = [0.0, 0.0, 0.0]
        for sign in (1.0, -1.0):
            o[axis] = 2.0 * sign
            d = [0.0, 0.0, 0.0]
            d[axis] = -sign
            assert _bbox_isect(orc, b, _ray(o, d)) == (1.0, 3.0)
            d[axis] = sign
            assert _bbox_isect(orc, b, _ray(o, d)) is None
    n = float(f32(1.0) / np.sqrt(f32(3.0)))
    dn = (f32(1.0) * (f32(1.0) / np.sqrt(f32(3.0))))
    d1, d2 = _bbox_isect(orc, b, _ray([-1, -1, -1], [dn, dn, dn]))
    assert d1 == 0.0 and abs(d2 - math.sqrt(12.0)) < 1e-6
    assert _bbox_isect(orc, b, _ray([1.5, 0.5, 0.5], [-0.5, 0.0, 0.5])) == (1.0, 1.0)
    assert _bbox_isect(orc, b, _ray([1.5, 0.5, 0.5], [-0.5, -0.5, 0.5])) == (1.0, 1.0)
    assert _bbox_isect(orc, b, _ray([2, 0, 0], [-1, 0, 0], mint=2.0)) == (2.0, 3.0)
    del n


def test_bbox_algebra(orc):
    """bbox.rs:225-401: empty(), surface_area, volume, max_extent tie rules"""
    L = orc.lib()

    def props(box):
        a, v, m, e = C.c_float(), C.c_float(), C.c_int32(), C.c_int32()
        L.orc_bbox_props(_p(np.array(box, np.float32)), C.byref(a), C.byref(v), C.byref(m), C.byref(e))
        return a.value, v.value, m.value, bool(e.value)

    empty = [F32_MAX] * 3 + [-F32_MAX] * 3
    assert props(empty)[3] and props(empty)[0] == 0.0 and props(empty)[2] == 2
    assert props([0, 0, 0, 0, 0, 0])[3] and props([0, 0, 0, 1, 0, 4])[3]
    assert props([2, 0, 0, 1, 3, 4])[3] and props([-2, 0, 0, 1, -3, 4])[3]
    assert not props([0, 0, 0, 1, 3, 4])[3]
    assert props([0, 0, 0, 1, 1, 1])[0] == 6.0
    assert props([1, 0, 0, 1, 1, 1])[0] == 2.0
    assert props([0, 0, 0, 2, 3, 4])[0] == 52.0
    assert props([1, 0, 0, 1, 1, 1])[1] == 0.0 and props([0, 0, 0, 2, 3, 4])[1] == 24.0
    assert props([0, 0, 0, 1, 1, 1])[2] == 2
    assert props([0, 0, 0, 2, 3, 4])[2] == 2
    assert props([0, 0, 0, 4, 5, 4])[2] == 1
    assert props([0, 10, 0, 4, 5, 4])[2] == 2


# ---- shape/mesh.rs --------------------------------------------------------------------------
TET_PTS = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
TET_TRIS = np.array([0, 3, 2, 0, 1, 2, 0, 3, 1, 1, 2, 3], np.uint32)


def _tet_tris(orc):
    v = np.zeros(12, np.uint32)
    orc.lib().orc_mesh_refine(_p(TET_TRIS), C.c_uint64(12), _p(v))
    return v.reshape(4, 3)


def test_mesh_refine_order_and_winding(orc):
    """shape/mesh.rs:425-436 it_can_be_refined_to_triangles"""
    v = _tet_tris(orc)
    assert v[3].tolist() == [2, 3, 0]
    assert v[2].tolist() == [2, 1, 0]
    assert v[1].tolist() == [1, 3, 0]
    assert v[0].tolist() == [3, 2, 1]


def test_triangles_can_be_intersected(orc):
    """shape/mesh.rs:478-497"""
    v = _tet_tris(orc)
    p9 = TET_PTS[v[0]].reshape(-1).copy()

    def hit(o, d):
        tbb = np.zeros(3, np.float32)
        return bool(orc.lib().orc_tri_intersect(_p(p9), _p(_ray(o, d)), _p(tbb)))

    assert hit([0, 0, 0], [1, 1, 1])
    assert not hit([1.5, 1.5, 1.5], [1, 1, 1])
    assert hit([1.5, 1.5, 1.5], [-1, -1, -1])
    assert not hit([1, 1, -1], [-1, -1, 1])


def test_triangle_areas(orc):
    """shape/mesh.rs:507-524"""
    L = orc.lib()
    L.orc_tri_area.restype = C.c_float
    v = _tet_tris(orc)
    for k in (1, 2, 3):
        assert L.orc_tri_area(_p(TET_PTS[v[k]].reshape(-1).copy())) == 0.5
        assert L.orc_tri_area(_p((2.0 * TET_PTS[v[k]]).reshape(-1).copy())) == 2.0


# ---- shape/sphere.rs ------------------------------------------------------------------------
def _sphere_isect(orc, o2w, o2w_inv, rad, z0, z1, pm, ray, want_dg=False):
    out3 = np.zeros(3, np.float32)
    dg = np.zeros(14, np.float32)
    ok = orc.lib().orc_sphere_intersect(_p(o2w), _p(o2w_inv), 0, rad, z0, z1, pm, _p(ray), _p(out3), _p(dg))
    return (out3, dg) if ok else None


def _translate(v):
    m, mi = _ident(), _ident()
    m[:3, 3] = v
    mi[:3, 3] = -np.array(v, np.float32)
    return m, mi


def test_sphere_creation(orc):
    """shape/sphere.rs:191-204 it_can_be_created"""
    o = np.zeros(6, np.float32)
    orc.lib().orc_sphere_props(1.0, -1.0, 1.0, 360.0, _p(o))
    assert o[0] == -1.0 and o[1] == 1.0
    assert o[2] == f32(np.arccos(f32(-1.0))) and o[3] == 0.0
    assert o[4] == f32(f32(np.pi) * f32(2.0))


def test_sphere_can_be_intersected(orc):
    """shape/sphere.rs:206-263"""
    m, mi = _translate([1.0, 2.0, 1.0])
    hit = lambda d: _sphere_isect(orc, m, mi, 1.0, -1.0, 1.0, 360.0, _ray([0, 0, 0], d)) is not None
    assert hit([1.0, 1.5, 1.0]) and hit([1.0, 1.0, 1.0]) and not hit([1.0, 0.5, 1.0])
    assert hit([0.0, 2.0, 1.0]) and hit([1.0, 2.0, 0.0])
    # partial sphere: translate(0,-3,0) * scale(2,2,2)
    m2 = np.array([[2, 0, 0, 0], [0, 2, 0, -3], [0, 0, 2, 0], [0, 0, 0, 1]], np.float32)
    m2i = np.zeros((4, 4), np.float32)
    assert orc.lib().orc_invert(_p(m2), _p(m2i)) == 0
    down = _ray([0, 0, 0], [0, -1, 0])
    assert _sphere_isect(orc, m2, m2i, 0.75, -0.75, -0.5, 180.0, down) is None
    assert _sphere_isect(orc, m2, m2i, 0.75, 0.5, 0.75, 180.0, down) is None
    assert _sphere_isect(orc, m2, m2i, 0.75, -0.5, 0.75, 180.0, down) is not None
    assert _sphere_isect(orc, m2, m2i, 0.75, -0.5, 0.75, 180.0, _ray([0, -3, 0], [0, -1, 0])) is None
    assert _sphere_isect(orc, m2, m2i, 0.75, -0.5, 0.75, 180.0, _ray([0, -3, 0], [0, 1, 0])) is not None
    assert _sphere_isect(orc, m2, m2i, 0.75, -0.5, 0.75, 180.0, _ray([0, -4, 10], [0, 0, -1])) is None


def test_sphere_intersection_information(orc):
    """shape/sphere.rs:266-292 it_has_intersection_information"""
    m, mi = _translate([0.0, -1.0, 0.0])
    out3, dg = _sphere_isect(orc, m, mi, 0.5, -0.5, 0.5, 360.0, _ray([0, 0, 0], [0, -1, 0]))
    assert out3[0] == 0.5 and out3[1] == f32(0.5) * f32(5e-4)
    assert dg[0:3].tolist() == [0.0, -0.5, 0.0]
    assert dg[3] == 0.0 and abs(dg[4] - 1.0) < 1e-6 and dg[5] == 0.0
    assert dg[6] == 0.25 and abs(dg[7] - 0.5) < 1e-6
    assert abs(dg[8] + math.pi) < 1e-6 and dg[9] == 0.0 and dg[10] == 0.0
    assert dg[11] == 0.0 and dg[12] == 0.0 and abs(dg[13] - math.pi / 2) < 1e-6


def test_sphere_areas(orc):
    """shape/sphere.rs:295-333 (transform-independent area)"""
    def area(rad, z0, z1, pm):
        o = np.zeros(6, np.float32)
        orc.lib().orc_sphere_props(rad, z0, z1, pm, _p(o))
        return float(o[5])
    pi = float(f32(np.pi))
    assert area(1.0, -1.0, 1.0, 360.0) == float(f32(4.0) * f32(np.pi))
    assert area(0.5, -1.0, 1.0, 360.0) == pi
    assert area(1.0, -1.0, 1.0, 180.0) == float(f32(2.0) * f32(np.pi))
    assert area(0.5, 0.0, 1.0, 360.0) == float(f32(0.5) * f32(np.pi))
    assert area(0.5, -1.0, 0.0, 360.0) == float(f32(0.5) * f32(np.pi))


# ---- primitive/aggregates --------------------------------------------------------------------
def _sphere_prims(pb, centres):
    mat = pb.Material.matte(pb.Texture.constant(0.5), pb.Texture.constant(0.0))
    out = []
    for v in centres:
        t = pb.Transform.translate(v)
        out.append(pb.Primitive.geometric(pb.Shape.sphere(t, t.inverse(), False, 1.0, -1.0, 1.0, 360.0), mat))
    return out


GET_SPHERES = [(0, 0, 0), (2, 0, 0), (0, 2, 0), (2, 2, 0), (0, 0, 2), (2, 0, 2), (0, 2, 2), (2, 2, 2)]


@pytest.mark.parametrize("sm", ["sah", "middle", "equal"])
def test_bvh_leaf_coverage(orc, sm):
    """bvh.rs:439-455 it_can_be_created"""
    import pbrt_rust_b200 as pb
    sc = pb.Scene.new_with(pb.Primitive.bvh(_sphere_prims(pb, GET_SPHERES), 1, sm), [])
    b, m = orc.OracleScene(sc).nodes()
    leaves = m[m[:, 2] == 1]
    assert np.all(leaves[:, 1] == 1)
    assert sorted(leaves[:, 0].tolist()) == list(range(8))


def _node_box(b, i):
    return b[i].tolist()


def test_bvh_arrange_by_middle(orc):
    """bvh.rs:490-511"""
    import pbrt_rust_b200 as pb
    cs = [(-4, 0, 0), (-2, 0, 0), (2, 0, 0), (4, 0, 0), (6, 0, 0), (8, 0, 0)]
    b, _ = orc.OracleScene(pb.Scene.new_with(pb.Primitive.bvh(_sphere_prims(pb, cs), 1, "middle"), [])).nodes()
    assert _node_box(b, 0) == [-5, -1, -1, 9, 1, 1]
    assert _node_box(b, 1) == [-5, -1, -1, -1, 1, 1]
    assert _node_box(b, 4) == [1, -1, -1, 9, 1, 1]


def test_bvh_arrange_by_equal_counts(orc):
    """bvh.rs:513-534"""
    import pbrt_rust_b200 as pb
    cs = [(-4, 0, 0), (-2, 0, 0), (2, 0, 0), (4, 0, 0), (6, 0, 0), (8, 0, 0)]
    b, _ = orc.OracleScene(pb.Scene.new_with(pb.Primitive.bvh(_sphere_prims(pb, cs), 1, "equal"), [])).nodes()
    assert _node_box(b, 0) == [-5, -1, -1, 9, 1, 1]
    assert _node_box(b, 1) == [-5, -1, -1, 3, 1, 1]
    assert _node_box(b, 6) == [3, -1, -1, 9, 1, 1]


def test_bvh_arrange_by_sah(orc):
    """bvh.rs:536-559"""
    import pbrt_rust_b200 as pb
    cs = [(-2, 2, 0), (-4, 0, 0), (2, 0, 0), (4, 2, -3), (4, 1.5, -1.5), (4, 1.5, 1.5), (4, 2, 3), (4, 0, 0)]
    b, _ = orc.OracleScene(pb.Scene.new_with(pb.Primitive.bvh(_sphere_prims(pb, cs), 1, "sah"), [])).nodes()
    assert _node_box(b, 0) == [-5, -1, -4, 5, 3, 4]
    assert _node_box(b, 1) == [-5, -1, -1, 3, 3, 1]
    assert _node_box(b, 6) == [3, -1, -4, 5, 3, 4]


@pytest.mark.parametrize("sm", ["sah", "middle", "equal"])
def test_aggregate_closest_hit_ids(orc, sm):
    """primitive/aggregates/mod.rs:93-138 test_intersection (BVH x3): rays hit ids[0], ids[4], ids[6]"""
    import pbrt_rust_b200 as pb
    osc = orc.OracleScene(pb.Scene.new_with(pb.Primitive.bvh(_sphere_prims(pb, GET_SPHERES), 1, sm), []))
    rays = np.stack([_ray([0, 0, -1], [0, 0, 1]), _ray([-1, 0, 2], [1, 0, 0]), _ray([4, 0, 0], [-2, 1, 1])])
    prim, tbb, _ = osc.trace_closest(rays)
    order = osc.prim_order()
    assert order[prim, 1].tolist() == [0, 4, 6]


def test_bvh_tetra_mesh_maxt(orc):
    """bvh.rs:457-488 it_can_refine_primitives: ray (0.25,-1,0.25)+(0,1,0) hits at t = 1"""
    import pbrt_rust_b200 as pb
    for mp, sm in ((1, "sah"), (10, "middle")):
        mesh = pb.Shape.triangle_mesh(pb.Transform.new(), pb.Transform.new(), False, TET_TRIS, TET_PTS)
        osc = orc.OracleScene(pb.Scene.new_with(pb.Primitive.bvh([pb.Primitive.geometric(mesh, None)], mp, sm), []))
        prim, tbb, _ = osc.trace_closest(_ray([0.25, -1.0, 0.25], [0, 1, 0])[None])
        assert prim[0] != 0xFFFFFFFF and tbb[0, 0] == 1.0
        occ, _ = osc.trace_any(_ray([0.25, -1.0, 0.25], [0, 1, 0])[None])
        assert occ[0] == 1


# ---- camera ---------------------------------------------------------------------------------
def _projection(orc, proj, sw):
    pi = np.zeros((4, 4), np.float32)
    assert orc.lib().orc_invert(_p(proj), _p(pi)) == 0
    r2s, s2r, r2c = (np.zeros((4, 4), np.float32) for _ in range(3))
    assert orc.lib().orc_projection(640, 480, _p(proj), _p(pi), _p(np.array(sw, np.float32)), _p(r2s), _p(s2r), _p(r2c)) == 0
    return r2s, s2r, r2c


def test_projection_matrices(orc):
    """camera/projective.rs:143-250"""
    ortho = _ident()
    ortho[2, 3] = -1.0
    flip = np.array([[1, 0, 0, 0], [0, -1, 0, 480], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    for proj in (ortho, _ident()):
        r2s, s2r, _ = _projection(orc, proj, [0, 640, 0, 480])
        assert np.array_equal(r2s, flip)
        fi = np.zeros((4, 4), np.float32)
        orc.lib().orc_invert(_p(flip), _p(fi))
        assert np.abs(s2r - fi).max() == 0.0 or np.allclose(s2r, fi, atol=0)
    r2s, s2r, _ = _projection(orc, _ident(), [0, 1, 0, 1])
    want = np.array([[1 / 640, 0, 0, 0], [0, -1 / 480, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    assert np.abs(r2s - want).max() < 5e-5
    # raster_to_camera (projective.rs:226-244): Vector transforms
    _, _, r2c = _projection(orc, ortho, [0, 1, 0, 1])
    xfv = lambda m, v: (m[:3, :3] @ np.array(v, np.float32)).astype(np.float32)
    assert xfv(r2c, [160, 120, 1]).tolist() == [0.25, -0.25, 1.0]
    assert np.sum((xfv(r2c, [480, 360, 0]) - np.array([0.75, -0.75, 0.0])) ** 2) < 1e-5
    _, _, r2c = _projection(orc, ortho, [-1, 1, -1, 1])
    assert xfv(r2c, [160, 120, 1]).tolist() == [0.5, -0.5, 1.0]
    assert np.sum((xfv(r2c, [480, 360, 0]) - np.array([1.5, -1.5, 0.0])) ** 2) < 1e-5


def test_film_extents(orc):
    """camera/film.rs:374-408 (cropped 142x12 film, 3x3 box filter)"""
    L = orc.lib()

    def ext(crop):
        s, p = np.zeros(4, np.int32), np.zeros(4, np.int32)
        L.orc_film_extents(142, 12, 3.0, 3.0, _p(np.array(crop, np.float32)), _p(s), _p(p))
        return tuple(int(x) for x in s), tuple(int(x) for x in p)

    ot, tt = float(f32(1.0) / f32(3.0)), float(f32(2.0) / f32(3.0))
    s, p = ext([ot, tt, ot, tt])
    assert s == (45, 98, 1, 11) and p == (48, 95, 4, 8)
    s, p = ext([0.0, ot, ot, tt])
    assert s == (-3, 51, 1, 11) and p == (0, 48, 4, 8)


def test_filters(orc):
    """filter.rs:136-210"""
    ev = orc.lib().orc_filter_eval
    for x in (0.0, 1.0, -1.0, 16.0, 0.001):
        for y in (2.0, 0.0, -0.01, math.pi):
            assert ev(0, 1.0, 1.0, 0, 0, x, y) == 1.0
    tri = lambda x, y: ev(1, 2.0, 2.0, 0, 0, x, y)
    assert tri(1, 0) == 0.5 and tri(0, 0) == 1.0 and tri(20, 0) == 0.0 and tri(-20, 0) == 0.0
    assert tri(0, 20) == 0.0 and tri(0, -20) == 0.0 and tri(1, 1) == 0.25 and tri(0.5, 0.5) == 0.5625
    assert tri(2, 0) == 0.0 and tri(0, -2) == 0.0
    for ty, p0, p1, centre in ((2, 1.0, 0.0, 0.9), (4, 1.0, 0.0, 0.9), (3, 0.2, 0.4, 0.8)):
        f = lambda x, y: ev(ty, 2.0, 2.0, p0, p1, x, y)
        for (x, y) in ((20, 0), (-20, 0), (0, 20), (0, -20), (2, 0), (0, -2)):
            assert f(x, y) == 0.0
        assert f(0, 0) > centre
    g = lambda x, y: ev(2, 2.0, 2.0, 1.0, 0, x, y)
    assert g(1, 1) > 0 and g(0.5, 0) > 0 and g(1, -0.5) > 0 and g(-1, -1.5) > 0


# ---- montecarlo.rs / rng.rs -------------------------------------------------------------------
def test_stratified_non_jittered(orc):
    """montecarlo.rs:194-249 (exact strata centres) + jittered strata bounds"""
    L = orc.lib()
    a = np.zeros(4, np.float32)
    L.orc_stratified_1d(C.c_uint64(0), C.c_uint64(4), 0, _p(a))
    assert a.tolist() == [0.5 / 4, 1.5 / 4, 2.5 / 4, 3.5 / 4]
    L.orc_stratified_1d(C.c_uint64(0), C.c_uint64(4), 1, _p(a))
    assert all(i / 4 <= a[i] <= (i + 1) / 4 for i in range(4))
    b = np.zeros(8, np.float32)
    L.orc_stratified_2d(C.c_uint64(0), C.c_uint64(2), C.c_uint64(2), 0, _p(b))
    assert b.tolist() == [0.25, 0.25, 0.75, 0.25, 0.25, 0.75, 0.75, 0.75]
    c = np.zeros(12, np.float32)
    L.orc_stratified_2d(C.c_uint64(0), C.c_uint64(3), C.c_uint64(2), 1, _p(c))
    for i in range(6):
        x, y = i % 3, i // 3
        assert x / 3 <= c[2 * i] <= (x + 1) / 3 and y / 2 <= c[2 * i + 1] <= (y + 1) / 2


def test_rng_shuffle_properties(orc):
    """rng.rs:49-87: shuffles are permutations; the lone 0 moves (seed 12), perm[0] != 0 (seed 120)"""
    L = orc.lib()
    xs = np.array([0] + [1] * 10, np.float32)
    L.orc_rng_shuffle(C.c_uint64(12), _p(xs), C.c_uint64(11), C.c_uint64(1))
    assert xs[0] == 1 and sorted(xs.tolist()) == [0.0] + [1.0] * 10
    perm = np.arange(11, dtype=np.float32)
    L.orc_rng_shuffle(C.c_uint64(120), _p(perm), C.c_uint64(11), C.c_uint64(1))
    assert perm[0] != 0 and sorted(perm.tolist()) == list(range(11))


def test_chacha_and_stdrng_known_answers(orc):
    """Third-party arithmetic (rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4, not vendored):
    RFC 8439 §2.3.2 block, all-zero-key ChaCha20/ChaCha12 keystreams, rand 0.8's own
    test_stdrng_construction value for StdRng::from_seed."""
    L = orc.lib()

    def block(state, rounds):
        i, o = np.array(state, np.uint32), np.zeros(16, np.uint32)
        L.orc_chacha_block(_p(i), rounds, _p(o))
        return o

    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    st = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + key + [1, 0x09000000, 0x4a000000, 0]
    assert block(st, 20)[:4].tolist() == [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3]
    z = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + [0] * 12
    assert block(z, 20).tobytes().hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    assert block(z, 12).tobytes().hex().startswith("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f")
    seed = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)
    k = np.frombuffer(seed, dtype="<u4").copy()
    w = np.zeros(2, np.uint32)
    L.orc_stream_words(_p(k), C.c_uint64(0), C.c_uint64(2), _p(w))
    assert int(w[0]) | (int(w[1]) << 32) == 10719222850664546238


def test_rng_float_range_and_determinism(orc):
    L = orc.lib()
    a, b = np.zeros(4096, np.float32), np.zeros(4096, np.float32)
    L.orc_rng_floats(C.c_uint64(3), C.c_uint64(4096), _p(a))
    L.orc_rng_floats(C.c_uint64(3), C.c_uint64(4096), _p(b))
    assert np.array_equal(a, b) and a.min() >= 0.0 and a.max() < 1.0 and 0.45 < a.mean() < 0.55
    L.orc_rng_floats(C.c_uint64(4), C.c_uint64(4096), _p(b))
    assert not np.array_equal(a, b)


def test_van_der_corput_as_written(orc):
    """sampler/utils.rs:6-35: the last bit-reversal step shifts by 2, not 1.  Checked
    against an independent integer restatement of the Rust lines, plus a few hand values."""
    L = orc.lib()
    M = 0xFFFFFFFF

    def vdc(n, scramble):
        n = ((n << 16) | (n >> 16)) & M
        n = (((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8)) & M
        n = (((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4)) & M
        n = (((n & 0x33333333) << 2) | ((n & 0xCCCCCCCC) >> 2)) & M
        n = (((n & 0x55555555) << 2) | ((n & 0xAAAAAAAA) >> 2)) & M   # as written
        n ^= scramble
        return float(np.float32(((n >> 8) & 0xffffff) / float(1 << 24)))

    def sobol2(n, s):
        v = 1 << 31
        while n:
            if (n & 1) == 0:
                s ^= v
            v ^= v >> 1
            n >>= 1
        return float(np.float32(((s >> 8) & 0xffffff) / float(1 << 24)))

    assert L.orc_van_der_corput(0, 0) == 0.0
    assert L.orc_van_der_corput(1, 0) == 0.0          # a true radical inverse would give 0.5
    for n in list(range(64)) + [255, 1023, 65535, 123456789]:
        for sc in (0, 0x9E3779B9, 0xFFFFFFFF):
            assert L.orc_van_der_corput(n, sc) == vdc(n, sc)
            assert L.orc_sobol2(n, sc) == sobol2(n, sc)


def test_d7_strict_flags_black_and_fixed_nonblack(orc):
    """BSDF::f as written is always black; the pbrt-semantics fix is not."""
    from pbrt_rust_b200 import scenes
    cfg = scenes.config1(xres=48, yres=36)
    osc = orc.OracleScene(cfg["scene"])
    oc = orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0)
    black = orc.render(osc, oc, strict_flags=True)
    lit = orc.render(osc, oc, strict_flags=False)
    assert black["rgb"].max() == 0.0 and lit["rgb"].max() > 0.1


def test_strict_and_default_modes_agree_for_box_filter(orc):
    """with a 0.5 box filter the two film modes only differ at samples that
    land exactly on a pixel boundary; images must agree everywhere else."""
    from pbrt_rust_b200 import scenes
    cfg = scenes.config3(nx=40, nz=20, xres=96, yres=54, xs=2, ys=2)
    osc = orc.OracleScene(cfg["scene"])
    a = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0))
    b = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=1))
    assert (np.abs(a["rgb"] - b["rgb"]).max(axis=-1) > 0).mean() <= 1e-3
    assert a["stats"]["camera_rays"] == b["stats"]["camera_rays"]


# ---- ImageTexture + MIPMap ("next" row 2): texture/imagemap.rs:211-418 -------------------------
import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


def _texels(name):
    """read_image (imagemap.rs:75-89) of the reference's fixture; texels committed by
    scripts/make_golden_textures.py."""
    return np.load(_os.path.join(_GOLD, name + ".npz"))["rgb8"].astype(np.float32) / np.float32(255)


def test_imagemap_rgb_textures(orc):  # imagemap.rs:222-244
    tex = orc.OracleMIPMap(_texels("checkerboard_square"), spectrum=True, do_trilinear=False, max_aniso=1.0,
                           wrap=0, scale=1.0, gamma=2.2)
    assert tex.eval_planar((0.25, 0.25, 0)).tolist() == [0, 0, 0]
    assert tex.eval_planar((0.75, 0.25, 0)).tolist() == [1, 1, 1]
    assert tex.eval_planar((0.25, 0.75, 0)).tolist() == [1, 1, 1]
    assert tex.eval_planar((0.75, 0.75, 0)).tolist() == [0, 0, 0]


def test_imagemap_float_textures(orc):  # imagemap.rs:246-267
    tex = orc.OracleMIPMap(_texels("checkerboard_stretched"), spectrum=False, do_trilinear=False, max_aniso=1.0,
                           wrap=0, scale=1.0, gamma=2.2)
    assert tex.eval_planar((0.25, 0.25, 0))[0] == 0.0
    assert tex.eval_planar((0.75, 0.25, 0))[0] == 1.0
    assert tex.eval_planar((0.25, 0.75, 0))[0] == 1.0
    assert tex.eval_planar((0.75, 0.75, 0))[0] == 0.0


def test_imagemap_repeat_wrap(orc):  # imagemap.rs:269-304
    tex = orc.OracleMIPMap(_texels("checkerboard_square"), True, False, 1.0, 0, 1.0, 2.2)
    x0 = tex.eval_planar((0.25, 0.25, 0)).tolist()
    for i in range(3):
        for j in range(3):
            for sx, sy in ((1, 1), (-1, 1), (1, -1), (-1, -1)):
                p = (np.float32(0.25) + np.float32(sx * i), np.float32(0.25) + np.float32(sy * j), 0)
                assert tex.eval_planar(p).tolist() == x0


def test_imagemap_black_wrap(orc):  # imagemap.rs:306-342
    tex = orc.OracleMIPMap(_texels("checkerboard_square"), True, False, 1.0, 1, 1.0, 2.2)
    for i in range(10):
        for j in range(10):
            dx = np.float32(1.0) + np.float32(i) * np.float32(0.1)
            dy = np.float32(1.0) + np.float32(j) * np.float32(0.1)
            for sx, sy in ((1, 1), (-1, 1), (1, -1), (-1, -1)):
                p = (np.float32(0.25) + sx * dx, np.float32(0.25) + sy * dy, 0)
                assert tex.eval_planar(p).tolist() == [0, 0, 0]


def test_imagemap_clamp_wrap(orc):  # imagemap.rs:344-374
    tex = orc.OracleMIPMap(_texels("checkerboard_square"), False, False, 1.0, 2, 1.0, 2.2)
    for i in range(10):
        for j in range(10):
            dx = np.float32(1.0) + np.float32(i) * np.float32(0.1)
            dy = np.float32(1.0) + np.float32(j) * np.float32(0.1)
            assert abs(1.0 - tex.eval_planar((np.float32(0.25) + dx, 0.25, 0))[0]) < 1e-4
            assert tex.eval_planar((np.float32(0.25) - dx, 0.25, 0))[0] == 0.0
            assert abs(1.0 - tex.eval_planar((np.float32(0.25) + dx, np.float32(0.25) - dy, 0))[0]) < 1e-4
            assert tex.eval_planar((np.float32(0.25) - dx, np.float32(0.25) - dy, 0))[0] == 0.0


def test_imagemap_isotropic_sampling(orc):  # imagemap.rs:376-391
    tex = orc.OracleMIPMap(_texels("checkerboard_stretched"), False, True, 1.0, 2, 1.0, 2.2)
    v = tex.eval_planar((0.51, 0.25, 0), (0.02, 0, 0), (0, 0.02, 0))[0]
    assert abs(v - 0.7) < 0.01, v


def test_imagemap_anisotropic_sampling(orc):  # imagemap.rs:393-417
    tex = orc.OracleMIPMap(_texels("checkerboard_stretched"), False, False, 100.0, 2, 1.0, 2.2)
    v = tex.eval_planar((0.51, 0.48, 0), (0.02, 0, 0), (0, 0.02, 0))[0]
    assert abs(v - 0.76) < 0.01, v
    v = tex.eval_planar((0.51, 0.48, 0), (0.02, 0, 0), (0, 0.002, 0))[0]
    assert abs(v - 0.88) < 0.01, v


def test_mipmap_structure_and_unreadable_file(orc):
    """mipmap.rs:159-204: 500x256 -> 512x256, ulog2(512) = 10 levels down to 1x1; an unreadable
    file gives a 1x1 map of scale^gamma (imagemap.rs:116-120)."""
    tex = orc.OracleMIPMap(_texels("checkerboard_stretched"), False, True, 1.0, 0, 1.0, 2.2)
    assert tex.levels() == 10
    assert [tex.level(i).shape[:2] for i in range(10)] == [(256 >> i if 256 >> i else 1, 512 >> i) for i in range(10)]
    one = orc.OracleMIPMap(None, True, True, 1.0, 0, 0.5, 2.0)
    assert one.levels() == 1 and one.level(0).ravel().tolist() == [0.25, 0.25, 0.25]
    L = orc.lib()
    assert [L.orc_modulo(a, 4) for a in (-5, -4, -1, 0, 3, 4, 9)] == [3, 0, 3, 0, 3, 0, 1]  # utils/mod.rs:219-223
    assert L.orc_sinc_1d(0.0, 2.0) == 1.0 and L.orc_sinc_1d(1.0, 2.0) == 0.0


# ---- Disk and Cylinder ("next" row 4): shape/disk.rs:139-305, shape/cylinder.rs:156-352 ---------

def _quadric(orc, shape, o2w, o2w_inv, a, b, c, pm, ray=None):
    """shape 1 = Cylinder(rad, z0, z1), 2 = Disk(height, radius, inner_radius).  Returns
    (hit, out3, dg14, props13)."""
    L = orc.lib()
    L.orc_quadric_intersect.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    out3, dg, props = np.zeros(3, np.float32), np.zeros(14, np.float32), np.zeros(13, np.float32)
    ok = L.orc_quadric_intersect(shape, _p(o2w), _p(o2w_inv), 0, a, b, c, pm, None if ray is None else _p(ray),
                                 _p(out3), _p(dg), _p(props))
    return bool(ok), out3, dg, props


def _xf(orc, kind, a3):
    m, mi = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
    orc.lib().orc_transform(kind, _p(np.array(a3, np.float32)), _p(m), _p(mi))
    return m, mi


def test_disk_creation_bounds_area(orc):
    """disk.rs:152-188 (creation, object bounds), :285-304 (areas)"""
    I = _ident()
    _, _, _, p = _quadric(orc, 2, I, I, 0.0, 1.0, 0.5, 360.0)
    assert p[4] == 0.0 and p[0] == 1.0 and p[5] == 0.5 and p[3] == f32(2.0) * f32(math.pi)
    assert p[7:13].tolist() == [-1.0, -1.0, 0.0, 1.0, 1.0, 0.0]
    m, mi = _xf(orc, 1, [1.0, 2.0, 3.0])
    _, _, _, p = _quadric(orc, 2, m, mi, 2.0, 0.0, 1.0, 90.0)
    assert p[4] == 2.0 and p[0] == 0.0 and p[5] == 1.0 and p[3] == f32(0.5) * f32(math.pi)
    assert p[7:13].tolist() == [0.0, 0.0, 2.0, 0.0, 0.0, 2.0]
    area = lambda ht, r, ri, pm: _quadric(orc, 2, I, I, ht, r, ri, pm)[3][6]
    assert area(0.0, 1.0, 0.0, 360.0) == f32(math.pi) and area(2381.0, 1.0, 0.0, 360.0) == f32(math.pi)
    assert area(0.0, 1.0, 0.0, 180.0) == f32(0.5) * f32(math.pi)
    assert area(0.0, 1.0, float(np.sqrt(f32(0.5))), 360.0) == f32(0.5) * f32(math.pi)


def test_disk_can_be_intersected(orc):
    """disk.rs:190-246"""
    I = _ident()
    down = lambda o: _ray(o, [0.0, 0.0, -1.0])
    hit = lambda *a, **k: _quadric(orc, 2, *a, **k)[0]
    assert hit(I, I, 0.0, 1.0, 0.0, 360.0, ray=down([0, 0, 1]))
    assert not hit(I, I, 0.0, 1.0, 0.0, 360.0, ray=down([2, 2, 1]))
    assert not hit(I, I, 0.0, 1.0, 0.5, 360.0, ray=down([0, 0, 1]))     # through the hole
    assert not hit(I, I, 2.0, 1.0, 0.0, 360.0, ray=down([0, 0, 1]))     # starts behind the disk
    assert hit(I, I, 0.0, 0.75, 0.25, 180.0, ray=down([0, 0.5, 1]))     # half pipe: top half
    assert not hit(I, I, 0.0, 0.75, 0.25, 180.0, ray=down([0, -0.5, 1]))
    m, mi = _xf(orc, 4, [180.0, 0, 0])                                  # rotate_z(180)
    assert not hit(m, mi, 0.0, 0.75, 0.25, 180.0, ray=down([0, 0.5, 1]))
    assert hit(m, mi, 0.0, 0.75, 0.25, 180.0, ray=down([0, -0.5, 1]))
    assert not hit(m, mi, 0.0, 0.75, 0.25, 180.0, ray=_ray([1.0, 10.5, 140.0], [-1.0, -10.5, -140.0]))


def test_disk_intersection_info(orc):
    """disk.rs:248-283"""
    I = _ident()
    ok, o3, dg, _ = _quadric(orc, 2, I, I, 0.0, 1.0, 0.0, 360.0, ray=_ray([0, 0, 1], [0, 0, -1]))
    assert ok and o3[0] == 1.0 and o3[1] == f32(5e-4)
    assert dg[0:3].tolist() == [0.0, 0.0, 0.0] and dg[3:6].tolist() == [0.0, 0.0, 1.0]
    ok, o3, dg, _ = _quadric(orc, 2, I, I, 0.0, 0.75, 0.25, 180.0, ray=_ray([0, 0.5, 1], [0, 0, -1]))
    assert ok and o3[0] == 1.0 and o3[1] == f32(5e-4)
    assert dg[0:3].tolist() == [0.0, 0.5, 0.0] and dg[3:6].tolist() == [0.0, 0.0, 1.0]
    assert dg[6] == 0.5 and dg[7] == 0.5


def test_cylinder_creation_bounds_area(orc):
    """cylinder.rs:171-205 (creation, bounds ignore transform and phi_max), :335-351 (areas)"""
    m, mi = _translate([1.0, 2.0, 3.0])
    _, _, _, p = _quadric(orc, 1, m, mi, 3.2, 14.0, -3.0, 16.0)
    assert p[0] == f32(3.2) and p[1] == -3.0 and p[2] == 14.0 and p[3] == f32(np.deg2rad(np.float64(16.0)))
    I = _ident()
    for o2w, pm in ((_ident(), 360.0), (_xf(orc, 1, [2.0, 3.0, 0.2])[0], 360.0), (_ident(), 180.0)):
        oi = np.zeros((4, 4), np.float32)
        assert orc.lib().orc_invert(_p(o2w), _p(oi)) == 0
        _, _, _, p = _quadric(orc, 1, o2w, oi, 0.5, -2.0, -1.0, pm)
        assert p[7:13].tolist() == [-0.5, -0.5, -2.0, 0.5, 0.5, -1.0]
    inv_pi = float(f32(1.0) / f32(math.pi))
    area = lambda o2w, oi, r, z0, z1, pm: _quadric(orc, 1, o2w, oi, r, z0, z1, pm)[3][6]
    assert area(I, I, inv_pi, 0.0, 1.0, 360.0) == 2.0 and area(I, I, inv_pi, 0.0, 1.0, 180.0) == 1.0
    s, si = _xf(orc, 1, [1.0, 2.0, 0.5])
    assert area(s, si, inv_pi, 0.0, 1.0, 360.0) == 2.0
    assert area(s, si, 0.0, 0.0, 1.0, 360.0) == 0.0 and area(s, si, 1.0, 0.0, 0.0, 360.0) == 0.0


def test_cylinder_can_be_intersected(orc):
    """cylinder.rs:207-303"""
    I = _ident()
    hit = lambda *a, **k: _quadric(orc, 1, *a, **k)[0]
    simple = lambda o, d: hit(I, I, 0.5, -1.0, 1.0, 360.0, ray=_ray(o, d))
    assert simple([1, 1, 1], [-1, -1, -1])
    assert not simple([1, 1, 1], [0, 0, -1])
    assert not simple([1, 1, 2], [-1, -1, 0])
    assert not simple([-1, 1, -2], [1, -1, 0])
    assert not simple([0.1, 0.1, -2], [0, 0, 1])            # down the middle
    assert simple([0.1, 0.1, -2], [-0.4, -0.4, 1.5])        # from the inside
    m, mi = _translate([1.0, 2.0, 3.0])
    partial = lambda o, d: hit(m, mi, 1.0, -3.0, 0.0, 90.0, ray=_ray(o, d))
    assert partial([2.0, 4.0, 1.5], [-1, -1, 0])
    assert not partial([1.0, 2.0, 1.5], [-1, -1, 0])
    assert partial([1.0, 2.0, 1.5], [1, 1, 0])
    assert not partial([2.5, 2.5, 10.0], [-1, -1, -10])     # barely misses
    for pm in (360.0, 0.1, 0.0, -1.0):                      # zero radius, tiny / zero / negative phi_max
        assert hit(I, I, 0.0, -1.0, 1.0, pm, ray=_ray([1, 1, -0.5], [-1, -1, 0]))
    assert hit(I, I, 0.0, 0.0, 0.0, 360.0, ray=_ray([1, 1, 1], [-1, -1, -1]))
    assert not hit(I, I, 0.0, -1.0, -1.0, 360.0, ray=_ray([1, 1, 1], [-1, -1, -1]))


def test_cylinder_intersection_information(orc):
    """cylinder.rs:305-333"""
    m, mi = _xf(orc, 3, [90.0, 0, 0])  # rotate_y(90)
    d = np.array([0.0, -1.0, -1.0], np.float32)
    d = d * (f32(1.0) / np.sqrt(f32(2.0)))
    ok, o3, dg, _ = _quadric(orc, 1, m, mi, 1.0, -1.0, 1.0, 180.0, ray=_ray([0, 1, 1], d.tolist()))
    s2 = math.sqrt(2.0)
    assert ok and abs(o3[0] - (s2 - 1.0)) < 1e-6 and abs(o3[1] - (s2 - 1.0) * 5e-4) < 1e-6
    assert np.sum((dg[0:3] - np.array([0.0, s2 / 2, s2 / 2])) ** 2) < 1e-6
    assert np.sum((dg[3:6] - np.array([0.0, 1.0, 1.0]) / s2) ** 2) < 1e-6
    assert dg[6] == 0.75 and abs(dg[7] - 0.5) < 1e-6
    assert np.sum((dg[8:11] - np.array([0.0, -math.pi * s2 / 2, math.pi * s2 / 2])) ** 2) < 1e-6
    assert np.sum((dg[11:14] - np.array([2.0, 0.0, 0.0])) ** 2) < 1e-6


# ---- textures and mappings: texture/{mod,checkerboard,uv,mapping2d}.rs tests ---------------------

def _dg15(p=(0, 0, 0), dpdx=(0, 0, 0), dpdy=(0, 0, 0), u=0.0, v=0.0, dudx=0.0, dudy=0.0, dvdx=0.0, dvdy=0.0):
    return np.array([*p, *dpdx, *dpdy, u, v, dudx, dudy, dvdx, dvdy], np.float32)


def _pad12(v):
    v = [float(x) for x in np.ravel(v)]
    return np.array(v + [0.0] * (12 - len(v)), np.float32)


def _map(orc, kind, params, dg):
    out, m = np.zeros(6, np.float32), _pad12(params)
    orc.lib().orc_mapping_map(kind, _p(m), _p(dg), _p(out))
    return out


class _TexScene:
    """An oracle scene used only as a texture table (orc_add_texture / orc_texture_eval)."""

    def __init__(self, orc):
        self.L = orc.lib()
        self.L.orc_scene_new.restype = C.c_void_p
        self.h = C.c_void_p(self.L.orc_scene_new())

    def add(self, kind, value=(0, 0, 0), map_kind=1, params=(1, 0, 0, 0, 1, 0, 0, 0), t1=0, t2=0, aa=0, t3=0):
        v, m = _pad12(value), _pad12(params)  # keep both arrays alive across the call
        return self.L.orc_add_texture(self.h, kind, _p(v), map_kind, _p(m), t1, t2, t3, aa)

    def eval(self, tex, dg):
        out = np.zeros(3, np.float32)
        self.L.orc_texture_eval(self.h, tex, _p(dg), _p(out))
        return out


_PLANAR_NEW = (1, 0, 0, 0, 1, 0, 0, 0)  # PlanarMapping2D::new(): vs = x, vt = y, ds = dt = 0


def test_constant_and_uv_textures(orc):
    """texture/mod.rs:93-97 const_texture_works, texture/uv.rs:37-49 uv_texture_works"""
    ts = _TexScene(orc)
    c = ts.add(0, value=(2.5, 2.5, 2.5))
    assert ts.eval(c, _dg15(p=(3, -1, 2), u=0.3)).tolist() == [2.5, 2.5, 2.5]
    uv = ts.add(2, map_kind=1, params=_PLANAR_NEW)
    assert ts.eval(uv, _dg15(p=(0.25, 0.25, 0.0))).tolist() == [0.25, 0.25, 0.0]
    assert ts.eval(uv, _dg15(p=(1.25, -2.5, 0.0))).tolist() == [0.25, 0.5, 0.0]


def test_checkerboard_texture(orc):
    """texture/checkerboard.rs:131-147 (point sampled) and :170-189 (closed-form antialiasing)"""
    ts = _TexScene(orc)
    one, zero = ts.add(0, value=(1, 1, 1)), ts.add(0, value=(0, 0, 0))
    chk = ts.add(1, map_kind=1, params=_PLANAR_NEW, t1=one, t2=zero, aa=0)
    assert ts.eval(chk, _dg15(p=(0.5, 0.5, 0)))[0] == 1.0
    assert ts.eval(chk, _dg15(p=(1.5, 0.5, 0)))[0] == 0.0
    assert ts.eval(chk, _dg15(p=(1.5, 1.5, 0)))[0] == 1.0
    aa = ts.add(1, map_kind=1, params=_PLANAR_NEW, t1=one, t2=zero, aa=1)
    assert abs(ts.eval(aa, _dg15(p=(1.1, 0.5, 0), dpdx=(0.2, 0, 0)))[0] - 0.25) < 1e-3
    assert abs(ts.eval(aa, _dg15(p=(1.1, 0.5, 0), dpdx=(0.2, 0.2, 0)))[0] - 0.25) < 1e-3
    v = ts.eval(aa, _dg15(p=(1.1, 0.9, 0), dpdx=(0.2, 0, 0), dpdy=(0, 0.2, 0)))[0]
    assert abs(v - 2.0 * (0.25 * 0.75)) < 1e-3


def test_uv_mapping(orc):
    """texture/mapping2d.rs:293-333 uv_mapping_can_map_coords / can_scale_coords (+ :216-246)"""
    new = (1, 1, 0, 0)
    assert _map(orc, 0, new, _dg15()).tolist() == [0, 0, 0, 0, 0, 0]
    assert _map(orc, 0, new, _dg15(u=0.5, v=0.2)).tolist() == [0.5, f32(0.2), 0, 0, 0, 0]
    assert _map(orc, 0, new, _dg15(u=0.5, v=0.2, dudx=10, dudy=12, dvdx=-1, dvdy=0)).tolist() == [0.5, f32(0.2), 10, -1, 12, 0]
    sc = (2.0, 0.5, 1.0, -0.3)
    assert _map(orc, 0, sc, _dg15()).tolist() == [1.0, f32(-0.3), 0, 0, 0, 0]
    assert _map(orc, 0, sc, _dg15(p=(0.1, 10.0, -13.0))).tolist() == [1.0, f32(-0.3), 0, 0, 0, 0]
    et = f32(0.2) * f32(0.5) - f32(0.3)
    assert _map(orc, 0, sc, _dg15(u=0.5, v=0.2)).tolist() == [2.0, et, 0, 0, 0, 0]
    assert _map(orc, 0, sc, _dg15(u=0.5, v=0.2, dudx=10, dudy=12, dvdx=-1, dvdy=0)).tolist() == [2.0, et, 20.0, -0.5, 24.0, 0.0]
    for params in (new, sc):  # test_uv_mapping_deriv
        a = _map(orc, 0, params, _dg15(u=0.5, v=0.1, dudx=10, dudy=12, dvdx=-1, dvdy=0))
        dx, dy = f32(1.0), f32(-0.4)
        b = _map(orc, 0, params, _dg15(u=f32(0.5) + dx * f32(10) + dy * f32(12), v=f32(0.1) + dx * f32(-1) + dy * f32(0),
                                      dudx=10, dudy=12, dvdx=-1, dvdy=0))
        assert a[2:].tolist() == b[2:].tolist()
        assert abs(a[0] + dx * a[2] + dy * a[4] - b[0]) < 1e-3 and abs(a[1] + dx * a[3] + dy * a[5] - b[1]) < 1e-3


def test_planar_mapping_differentials(orc):
    """texture/mapping2d.rs:449-463 via test_positional_differentials (:248-284)"""
    for params in (_PLANAR_NEW, (1.0, 2.0, 3.0, -3.0, 0.0, -1.2, 3.2, -1000.0)):
        base = _map(orc, 1, params, _dg15())
        assert _map(orc, 1, params, _dg15(u=0.5, v=0.1, dudx=10, dudy=12, dvdx=-1, dvdy=0)).tolist() == base.tolist()
        p, dpdx, dpdy = np.array([0.3, 1.2, -4.0], np.float32), np.array([0.2, 0.0, -0.3], np.float32), np.array([-0.5, 0.1, 1.3], np.float32)
        a = _map(orc, 1, params, _dg15(p=p, dpdx=dpdx, dpdy=dpdy))
        dx, dy = f32(0.1), f32(-0.1)
        b = _map(orc, 1, params, _dg15(p=p + dx * dpdx + dy * dpdy, dpdx=dpdx, dpdy=dpdy))
        assert np.abs(a[2:] - b[2:]).max() < 0.01
        assert abs(a[0] + dx * a[2] + dy * a[4] - b[0]) < 1e-3 and abs(a[1] + dx * a[3] + dy * a[5] - b[1]) < 1e-3


# ---- spherical / cylindrical / 3D mappings, scale / mix / bilerp textures, noise ------------------

def _w2t(orc, *ops):
    """rows 0..2 of a product of transforms: ops = (kind, args) as for _xf"""
    m = np.eye(4, dtype=np.float32)
    for kind, a in ops:
        mm, _ = _xf(orc, kind, a)
        m = (m.astype(np.float32) @ np.asarray(mm, np.float32).reshape(4, 4)).astype(np.float32)
    return m.reshape(-1)[:12]


_IDENT12 = (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)


def _positional_differentials(orc, kind, params):
    """texture/mapping2d.rs:248-284 test_positional_differentials"""
    base = _map(orc, kind, params, _dg15())
    assert _map(orc, kind, params, _dg15(u=0.5, v=0.1, dudx=10, dudy=12, dvdx=-1, dvdy=0)).tolist() == base.tolist()
    p, dpdx, dpdy = np.array([0.3, 1.2, -4.0], np.float32), np.array([0.2, 0.0, -0.3], np.float32), np.array([-0.5, 0.1, 1.3], np.float32)
    a = _map(orc, kind, params, _dg15(p=p, dpdx=dpdx, dpdy=dpdy))
    dx, dy = f32(0.1), f32(-0.1)
    b = _map(orc, kind, params, _dg15(p=p + dx * dpdx + dy * dpdy, dpdx=dpdx, dpdy=dpdy))
    assert np.abs(a[2:] - b[2:]).max() < 0.01
    assert abs(a[0] + dx * a[2] + dy * a[4] - b[0]) < 1e-3 and abs(a[1] + dx * a[3] + dy * a[5] - b[1]) < 1e-3


@pytest.mark.parametrize("kind", [2, 3])
def test_spherical_and_cylindrical_mapping_can_map_coords(orc, kind):
    """texture/mapping2d.rs:335-375 (spherical) and :392-432 (cylindrical)"""
    assert _map(orc, kind, _IDENT12, _dg15()).tolist() == [0.5, 0.0, 0.0, 0.0, 0.0, 0.0]
    assert _map(orc, kind, _IDENT12, _dg15(u=0.5, v=0.2)).tolist() == [0.5, 0.0, 0.0, 0.0, 0.0, 0.0]
    s, t = _map(orc, kind, _IDENT12, _dg15(p=(0.1, 0.2, 0.6)))[:2]
    assert s != 0.0 and t != 0.0
    for i in range(1, 9):
        di = f32(i) / f32(10.0)
        ns, nt = _map(orc, kind, _IDENT12, _dg15(p=(f32(0.1) * di, f32(0.2) * di, f32(0.6) * di)))[:2]
        assert abs(s - ns) < 1e-5 and abs(t - nt) < 1e-5
    # transformed_*_mapping_can_map_coords: translate(1,2,3) at the origin == identity at (1,2,3)
    tr = _w2t(orc, (0, [1.0, 2.0, 3.0]))
    assert _map(orc, kind, tr, _dg15()).tolist() == _map(orc, kind, _IDENT12, _dg15(p=(1.0, 2.0, 3.0))).tolist()


@pytest.mark.parametrize("kind", [2, 3])
def test_spherical_and_cylindrical_mapping_differentials(orc, kind):
    """texture/mapping2d.rs:377-390, 434-447: identity and translate(1,2,3) * rotate_x(45)"""
    _positional_differentials(orc, kind, _IDENT12)
    _positional_differentials(orc, kind, _w2t(orc, (0, [1.0, 2.0, 3.0]), (2, [45.0, 0, 0])))


def test_identity_mapping_3d(orc):
    """texture/mapping3d.rs:70-116 test_positional_differentials for IdentityMapping3D"""
    L = orc.lib()

    def m3(params, dg):
        out, m = np.zeros(9, np.float32), _pad12(params)
        L.orc_mapping3d_map(_p(m), _p(dg), _p(out))
        return out

    for params in (_IDENT12, _w2t(orc, (0, [1.0, 2.0, 3.0]), (2, [45.0, 0, 0]))):
        base = m3(params, _dg15())
        assert m3(params, _dg15(u=0.5, v=0.1, dudx=10, dudy=12, dvdx=-1, dvdy=0)).tolist() == base.tolist()
        p, dpdx, dpdy = np.array([0.3, 1.2, -4.0], np.float32), np.array([0.2, 0.0, -0.3], np.float32), np.array([-0.5, 0.1, 1.3], np.float32)
        a = m3(params, _dg15(p=p, dpdx=dpdx, dpdy=dpdy))
        dx, dy = f32(0.1), f32(-0.1)
        b = m3(params, _dg15(p=p + dx * dpdx + dy * dpdy, dpdx=dpdx, dpdy=dpdy))
        assert a[3:].tolist() == b[3:].tolist()
        ep = a[0:3] + dx * a[3:6] + dy * a[6:9]
        assert float(np.sum((ep - b[0:3]) ** 2)) < 1e-3


def test_scale_mix_and_bilerp_textures(orc):
    """texture/mod.rs:99-105 scale_texture_works, mix.rs:35-45 mix_texture_works,
    bilerp.rs:45-61 bilerp_texture_works"""
    ts = _TexScene(orc)
    two, vec = ts.add(0, value=(2, 2, 2)), ts.add(0, value=(1, 2, 3))
    assert ts.eval(ts.add(4, t1=two, t2=vec), _dg15()).tolist() == [2.0, 4.0, 6.0]
    a, b, amt = ts.add(0, value=(1, -2, 15)), ts.add(0, value=(1, 2, 3)), ts.add(0, value=(0.75, 0.75, 0.75))
    assert ts.eval(ts.add(5, t1=a, t2=b, t3=amt), _dg15()).tolist() == [1.0, 1.0, 6.0]
    bil = ts.add(6, value=(2, 2, 2, 3, 3, 3, 1, 1, 1, 4, 4, 4), map_kind=0, params=(1, 1, 0, 0))
    assert ts.eval(bil, _dg15())[0] == 2.0
    assert ts.eval(bil, _dg15(u=0.5, v=0.5))[0] == 2.5
    assert ts.eval(bil, _dg15(u=1.0, v=0.5))[0] == 2.5
    assert ts.eval(bil, _dg15(u=1.0, v=1.0))[0] == 4.0


def test_noise(orc):
    """texture/noise.rs:151-175 noise_is_zero_at_integers / noise_is_nonzero_at_nonintegers"""
    L = orc.lib()
    for i in range(-10, 10):
        for j in range(-10, 10):
            for k in range(-10, 10):
                assert L.orc_noise(float(i), float(j), float(k)) == 0.0
                v = abs(L.orc_noise(f32(i) + f32(0.3), f32(j) + f32(0.2), f32(k) + f32(0.1)))
                assert 0.0 < v <= 1.0


@pytest.mark.parametrize("turb,p", [(0, (0.3, -0.4, 10.2)), (1, (0.3, -0.3, 10.2))])
def test_fbm_and_turbulence_are_more_or_less_continuous(orc, turb, p):
    """texture/noise.rs:177-213"""
    L = orc.lib()
    dpdx, dpdy = np.array([0.1, 0, 0], np.float32), np.array([0, 0.1, 0], np.float32)
    p = np.array(p, np.float32)
    v = L.orc_fbm(turb, _p(p), _p(dpdx), _p(dpdy), 1.0, 10)
    q = p + np.array([0.01, -0.01, 0.005], np.float32)
    v2 = L.orc_fbm(turb, _p(q), _p(dpdx), _p(dpdy), 1.0, 10)
    assert abs(v) > 0.0 and abs(v2) > 0.0 and v != v2 and abs(v - v2) < 0.06


def test_dots_and_noise_textures_through_the_table(orc):
    """DotsTexture (dots.rs:23-46) picks inside/outside by the noise lattice; FBm / Wrinkled
    textures (fbm.rs) equal fbm / turbulence at the mapped point."""
    ts = _TexScene(orc)
    one, zero = ts.add(0, value=(1, 1, 1)), ts.add(0, value=(0, 0, 0))
    dots = ts.add(7, map_kind=1, params=_PLANAR_NEW, t1=one, t2=zero)
    L = orc.lib()
    seen = set()
    for ix in range(-6, 7):
        for iy in range(-6, 7):
            cell_has_dot = L.orc_noise(f32(ix) + f32(0.5), f32(iy) + f32(0.5), 0.5) > 0.0
            centre = ts.eval(dots, _dg15(p=(ix, iy, 0)))[0]
            corner = ts.eval(dots, _dg15(p=(ix + 0.49, iy + 0.49, 0)))[0]
            assert corner == 0.0  # |offset| >= 0.49 - 0.15 per axis: always outside the 0.35 radius
            if not cell_has_dot:
                assert centre == 0.0
            else:
                assert centre == 1.0  # the centre shifts by at most 0.15 * sqrt(2) < 0.35
            seen.add(bool(cell_has_dot))
    assert seen == {True, False}
    p, dpdx, dpdy = np.array([0.3, -0.4, 10.2], np.float32), np.array([0.1, 0, 0], np.float32), np.array([0, 0.1, 0], np.float32)
    for kind, turb in ((8, 0), (9, 1)):
        t = ts.add(kind, value=(0.5, 0, 0), map_kind=4, params=_IDENT12, aa=6)
        want = L.orc_fbm(turb, _p(p), _p(dpdx), _p(dpdy), 0.5, 6)
        got = ts.eval(t, _dg15(p=p, dpdx=dpdx, dpdy=dpdy))
        assert got.tolist() == [want, want, want]


def _bump(orc, ts, tex, dgs, ng):
    out, g, n = np.zeros(9, np.float32), np.array(dgs, np.float32), np.array(ng, np.float32)
    orc.lib().orc_bump(ts.h, tex, _p(g), _p(n), _p(out))
    return out


def test_bump_mapping(orc):
    """material::bump (material/mod.rs:23-77).  The reference's own test is an `unimplemented!()` stub
    (:146-155), so these are properties of the function as written: a constant displacement d leaves
    dpdu + d * dndu, a displacement linear in u tilts dpdu along the normal by its slope, and the
    result is flipped to the side of the geometric normal."""
    ts = _TexScene(orc)
    #        p        dpdu     dpdv     dndu         dndv         nn       u    v    dudx dudy dvdx dvdy  flip      dpdx, dpdy
    dgs = [0, 0, 0, 1, 0, 0, 0, 1, 0, 0.5, 0, 0, 0, 0.25, 0, 0, 0, 1, 0.3, 0.6, 0.2, 0.0, 0.0, 0.4, 0, 0, 0] + [0] * 6
    const = ts.add(0, value=(0.5, 0.5, 0.5))
    b = _bump(orc, ts, const, dgs, (0, 0, 1))
    assert b[0:3].tolist() == [1.25, 0, 0] and b[3:6].tolist() == [0, 1.125, 0] and b[6:9].tolist() == [0, 0, 1]
    assert _bump(orc, ts, const, dgs, (0, 0, -1))[6:9].tolist() == [0, 0, -1]  # face_forward
    dgs_flip = list(dgs)
    dgs_flip[24] = 1.0
    assert _bump(orc, ts, const, dgs_flip, (0, 0, -1))[6:9].tolist() == [0, 0, -1]
    # displacement = 2 * u (bilerp over UV: v00 = v01 = 0, v10 = v11 = 2)
    lin = ts.add(6, value=(0, 0, 0, 0, 0, 0, 2, 2, 2, 2, 2, 2), map_kind=0, params=(1, 1, 0, 0))
    dgs0 = list(dgs)
    dgs0[9:15] = [0] * 6  # no dndu / dndv
    b = _bump(orc, ts, lin, dgs0, (0, 0, 1))
    assert abs(b[0] - 1.0) < 1e-6 and abs(b[2] - 2.0) < 1e-5 and b[1] == 0.0      # dpdu = (1, 0, slope)
    assert np.abs(b[3:6] - np.array([0, 1, 0])).max() < 1e-6                       # no v dependence
    n = np.array([-2.0, 0.0, 1.0]) / math.sqrt(5.0)
    assert np.abs(b[6:9] - n).max() < 1e-6


# ---- Matrix4x4 / Transform: transform/matrix4x4.rs and transform/transform.rs tests ---------------

def _m(rows):
    return np.array(rows, np.float32).reshape(4, 4)


def _mat_mul(orc, a, b):
    out = np.zeros((4, 4), np.float32)
    orc.lib().orc_mat_mul(_p(_m(a)), _p(_m(b)), _p(out))
    return out


def _invert(orc, a):
    out = np.zeros((4, 4), np.float32)
    rc = orc.lib().orc_invert(_p(_m(a)), _p(out))
    return rc, out


class _X:
    """Transform = (m, m_inv) with the oracle's algebra"""

    def __init__(self, orc, m, mi):
        self.orc, self.m, self.mi = orc, np.ascontiguousarray(m, np.float32), np.ascontiguousarray(mi, np.float32)

    @staticmethod
    def mk(orc, kind, a3):
        return _X(orc, *_xf(orc, kind, a3))

    @staticmethod
    def from_matrix(orc, m):
        rc, mi = _invert(orc, m)
        assert rc == 0
        return _X(orc, _m(m), mi)

    def __mul__(self, o):
        m, mi = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
        self.orc.lib().orc_xf_mul(_p(self.m), _p(self.mi), _p(o.m), _p(o.mi), _p(m), _p(mi))
        return _X(self.orc, m, mi)

    def inverse(self):
        return _X(self.orc, self.mi, self.m)

    def apply(self, kind, v):
        out, v = np.zeros(3, np.float32), np.array(v, np.float32)
        self.orc.lib().orc_xf_apply(_p(self.m), _p(self.mi), kind, _p(v), _p(out))
        return out

    pt = lambda self, v: self.apply(0, v)
    vec = lambda self, v: self.apply(1, v)
    nrm = lambda self, v: self.apply(2, v)

    def swaps(self):
        return bool(self.orc.lib().orc_xf_swaps_handedness(_p(self.m), _p(self.mi)))


def _nrm3(v):
    v = np.array(v, np.float32)
    return v * (f32(1.0) / np.sqrt(np.sum(v * v, dtype=np.float32)))


def test_matrix4x4_transpose_and_product(orc):
    """transform/matrix4x4.rs:277-330 it_can_be_transposed / they_can_be_multiplied"""
    a = [[1, 2, 3, 4], [4, 3, 2, 1], [-1, 2, -3, 4], [0, 0, 0, 0]]
    out = np.zeros((4, 4), np.float32)
    orc.lib().orc_mat_transpose(_p(_m(a)), _p(out))
    assert out.tolist() == [[1, 4, -1, 0], [2, 3, 2, 0], [3, 2, -3, 0], [4, 1, 4, 0]]
    ident = np.eye(4, dtype=np.float32)
    m1 = [[1, 4, -1, 0], [2, 3, 2, 0], [3, 2, -3, 0], [4, 1, 4, 0]]
    m2 = [[3, -2, -1, 0], [0, 0.1, -2, 3], [2, 6, 3, 0], [6, 6, 1, 1]]
    assert np.array_equal(_mat_mul(orc, ident, ident), ident)
    assert np.array_equal(_mat_mul(orc, m1, ident), _m(m1)) and np.array_equal(_mat_mul(orc, ident, m2), _m(m2))
    want = _m([[1, -7.6, -12, 12], [10, 8.3, -2, 9], [3, -23.8, -16, 6], [20, 16.1, 6, 3]])
    assert np.array_equal(_mat_mul(orc, m1, m2), want)
    assert not np.array_equal(_mat_mul(orc, m2, m1), want)


def test_matrix4x4_inverse(orc):
    """transform/matrix4x4.rs:332-375 it_can_be_inverted (check_mat!: |diff| < 5e-5) and
    it_cant_invert_singular_matrices (panic -> error)"""
    rc, inv = _invert(orc, np.eye(4))
    assert rc == 0 and np.abs(inv - np.eye(4)).max() < 5e-5
    m = [[1, -2, 3, 0], [2, -5, 12, 0], [0, 2, -10, 0], [0, 0, 0, 1]]
    rc, inv = _invert(orc, m)
    assert rc == 0
    assert np.abs(_mat_mul(orc, m, inv) - np.eye(4)).max() < 5e-5 and np.abs(_mat_mul(orc, inv, m) - np.eye(4)).max() < 5e-5
    n = [[2, 3, 1, 5], [1, 0, 3, 1], [0, 2, -3, 2], [0, 2, 3, 1]]
    rc, inv = _invert(orc, n)
    assert rc == 0 and np.abs(inv - _m([[18, -35, -28, 1], [9, -18, -14, 1], [-2, 4, 3, 0], [-12, 24, 19, -1]])).max() < 5e-5
    m2 = [[-0.70710677, -0.40824828, -0.57735026, 1.0], [0.0, 0.81649655, -0.57735026, 1.0],
          [0.70710677, -0.40824828, -0.57735026, 1.0], [0, 0, 0, 1]]
    rc, inv = _invert(orc, m2)
    want = _m([[-0.70710677, 0.0, 0.70710677, 0.0], [-0.40824828, 0.81649655, -0.40824828, 0.0],
               [-0.57735026, -0.57735026, -0.57735026, 1.73205], [0, 0, 0, 1]])
    assert rc == 0 and np.abs(inv - want).max() < 5e-5
    rc, _ = _invert(orc, [[32, 8, 11, 17], [8, 20, 17, 23], [11, 17, 14, 26], [17, 23, 26, 2]])
    assert rc != 0 and b"ingular" in orc.lib().orc_last_error()


def test_transform_vectors_points_normals(orc):
    """transform/transform.rs: it_can_transform_vectors, it_cannot_translate_vectors,
    it_can_transform_points, it_can_transform_normals, it_can_translate_points, it_can_scale_vectors"""
    s2 = f32(np.sqrt(f32(2.0)))
    rx45 = _X.mk(orc, 2, [45.0, 0, 0])
    v = _nrm3([1, 1, 0])
    vt = np.array([s2 / f32(2), 0.5, 0.5], np.float32)
    assert float(np.sum((rx45.vec(v) - vt) ** 2)) < 1e-6
    ident = _X(orc, np.eye(4), np.eye(4))
    assert ident.vec(vt).tolist() == vt.tolist()
    x2 = _X.mk(orc, 0, [1, 2, 3]) * _X.mk(orc, 3, [90.0, 0, 0])
    assert float(np.sum((x2.vec([1, 1, 1]) - np.array([1, 1, -1])) ** 2)) < 1e-6
    assert x2.vec([0, 0, 0]).tolist() == [0, 0, 0] and ident.vec([1, 5, -13]).tolist() == [1, 5, -13]
    assert _X.mk(orc, 0, [1, 4, -300]).vec([1, 1, 1]).tolist() == [1, 1, 1]
    assert x2.pt([1, 1, 1]).tolist() == [2, 3, 2] and x2.pt([0, 0, 0]).tolist() == [1, 2, 3]
    assert ident.pt([1, 5, -13]).tolist() == [1, 5, -13]
    assert x2.nrm(_nrm3([1, 1, 1])).tolist() == _nrm3([1, 1, -1]).tolist()
    assert _nrm3(_X.mk(orc, 1, [2, 2, 2]).nrm([1, 0, 0])).tolist() == [1, 0, 0]
    got = _nrm3(_X.mk(orc, 3, [45.0, 0, 0]).nrm(_nrm3([1, 1, 1])))
    assert got.tolist() == [f32(0.8164967), f32(0.5773503), 0.0]
    assert _X.mk(orc, 0, [1, 2, 3]).pt([1, 1, 1]).tolist() == [2, 3, 4]
    sc = _X.mk(orc, 1, [2.0, 0.5, 100.0])
    assert sc.vec([1, 2, 0]).tolist() == [2, 1, 0] and sc.vec([-1, 0, -0.01]).tolist() == [-2, 0, -1]


def test_transform_inverse_and_rotations(orc):
    """transform/transform.rs: it_can_be_inverted, it_can_rotate_about_x / y / z"""
    x = _X.from_matrix(orc, [[2, 3, 1, 5], [1, 0, 3, 1], [0, 2, -3, 2], [0, 2, 3, 1]])
    p = np.array([1, 2, 3], np.float32)
    assert float(np.sum((x.inverse().pt(x.pt(p)) - p) ** 2)) < 1e-6
    s2 = f32(np.sqrt(f32(2.0)))
    for kind, axis, src, want in ((2, [1, 0, 0], [1, 1, 0], [s2 / 2, 0.5, 0.5]),
                                  (3, [0, 1, 0], [1, 1, 0], [0.5, s2 / 2, -0.5]),
                                  (4, [0, 0, 1], [0, 1, 1], [-0.5, 0.5, s2 / 2])):
        r = _X.mk(orc, kind, [45.0, 0, 0])
        assert r.vec(axis).tolist() == axis and r.vec([0, 0, 0]).tolist() == [0, 0, 0]
        assert float(np.sum((r.vec(_nrm3(src)) - np.array(want, np.float32)) ** 2)) < 1e-6


def test_transform_look_at(orc):
    """transform/transform.rs it_can_look_at_a_point"""
    def look(pos, at, up):
        m, mi = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
        a, b, c = (np.array(v, np.float32) for v in (pos, at, up))
        orc.lib().orc_look_at(_p(a), _p(b), _p(c), _p(m), _p(mi))
        return _X(orc, m, mi)
    x = look([1, 1, 1], [0, 0, 0], [0, 1, 0])
    o = x.pt([0, 0, 0])
    assert abs(o[0]) < 1e-6 and abs(o[1]) < 1e-6 and o[2] > 1.0
    q = x.pt([-1, -1, -1])
    assert abs(q[0]) < 1e-6 and abs(q[1]) < 1e-6 and q[2] > 2.0
    x2 = look([1, 2, 3], [-1, 0, 4], [0, 1, 0])
    z = x2.pt([0, 0, -3])
    assert z[2] == 0.0 and abs(z[0]) > 0.1 and abs(z[1]) > 0.1
    mid = x2.pt([0, 1, 3.5])
    assert mid[2] > 1.0 and abs(mid[0]) < 1e-6 and abs(mid[1]) < 1e-6


def test_transform_handedness_rays_and_bboxes(orc):
    """transform/transform.rs: it_can_detect_handedness_swap, it_can_transform_rays,
    it_can_transform_bboxes"""
    ident = _X(orc, np.eye(4), np.eye(4))
    assert not ident.swaps()
    flips = []
    for k in range(3):
        m = np.eye(4, dtype=np.float32)
        m[k, k] = -1.0
        flips.append(_X.from_matrix(orc, m))
        assert flips[-1].swaps()
    m2 = flips[2]
    x1 = m2 * _X.mk(orc, 2, [34.0, 0, 0])
    x2 = x1 * _X.mk(orc, 0, [1, -2, 4])
    la_m, la_mi = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
    a, b, c = (np.array(v, np.float32) for v in ([1, 0, -3], [15, 12, -0.0], [0, 1, 0]))
    orc.lib().orc_look_at(_p(a), _p(b), _p(c), _p(la_m), _p(la_mi))
    x3 = x2 * _X(orc, la_m, la_mi)
    assert x1.swaps() and x2.swaps() and x3.swaps() and not (x3 * m2).swaps()
    # rays: only o and d move
    ray = np.array([0, 0, 0, 0.0, 1, 0, 0, 3.4028235e38], np.float32)
    out = np.zeros(8, np.float32)
    orc.lib().orc_xf_ray(_p(ident.m), _p(ident.mi), _p(ray), _p(out))
    assert out.tolist() == ray.tolist()
    x = _X.mk(orc, 0, [1, 1, 1]) * _X.mk(orc, 3, [90.0, 0, 0])
    orc.lib().orc_xf_ray(_p(x.m), _p(x.mi), _p(ray), _p(out))
    assert out[0:3].tolist() == [1, 1, 1] and float(np.sum((out[4:7] - np.array([0, 0, -1])) ** 2)) < 1e-6
    assert out[3] == 0.0 and out[7] == ray[7]
    # bboxes
    box = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    got = np.zeros(6, np.float32)
    orc.lib().orc_xf_bbox(_p(ident.m), _p(ident.mi), _p(box), _p(got))
    assert got.tolist() == box.tolist()
    x = _X.mk(orc, 0, [-3, 0, 1]) * _X.mk(orc, 2, [45.0, 0, 0])
    orc.lib().orc_xf_bbox(_p(x.m), _p(x.mi), _p(box), _p(got))
    s2 = f32(np.sqrt(f32(2.0)))
    want = np.array([-4.0, -s2, -s2 + f32(1.0), -2.0, s2, s2 + f32(1.0)], np.float32)
    assert got.tolist() == want.tolist()  # assert_eq! in the reference: exact


# ---- geometry/{vector,point,normal}.rs ------------------------------------------------------------

def _vop(orc, op, a, b=(0, 0, 0), n=3):
    out, a, b = np.zeros(6, np.float32), np.array(a, np.float32), np.array(b, np.float32)
    orc.lib().orc_vec_op(op, _p(a), _p(b), _p(out))
    return out[:n] if n > 1 else out[0]


def test_vector_length_cross_dot_coordinate_system(orc):
    """geometry/vector.rs:243-292 (lengths, cross), :431-466 (dot), :482-490 (coordinate_system)"""
    inf, nan = float("inf"), float("nan")
    for v, want in (((0, 0, 0), 0.0), ((1, 0, 0), 1.0), ((-1, 0, 0), 1.0), ((1, 1, 1), 3.0), ((0, 2, 0), 4.0),
                    ((0, 0, 2), 4.0), ((inf, 0, 0), inf), ((0, inf, 0), inf), ((0, 0, inf), inf)):
        assert _vop(orc, 7, v, n=1) == want and _vop(orc, 2, v, n=1) == np.sqrt(f32(want))
    for k in range(3):
        v = [0.0, 0.0, 0.0]
        v[k] = nan
        assert np.isnan(_vop(orc, 7, v, n=1)) and np.isnan(_vop(orc, 2, v, n=1))
    x, y, z = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    neg = lambda v: [-c for c in v]
    assert _vop(orc, 0, x, y).tolist() == list(z) and _vop(orc, 0, y, x).tolist() == neg(z)
    assert _vop(orc, 0, y, z).tolist() == list(x) and _vop(orc, 0, z, y).tolist() == neg(x)
    assert _vop(orc, 0, z, x).tolist() == list(y) and _vop(orc, 0, x, z).tolist() == neg(y)
    assert _vop(orc, 0, (1, 1, 0), x).tolist() == neg(z) and _vop(orc, 0, x, (1, 1, 0)).tolist() == list(z)
    assert _vop(orc, 1, x, y, 1) == 0 and _vop(orc, 1, y, z, 1) == 0 and _vop(orc, 1, z, x, 1) == 0
    assert _vop(orc, 1, x, x, 1) == 1 and _vop(orc, 1, y, y, 1) == 1 and _vop(orc, 1, z, z, 1) == 1
    v = (2, 3, 4)
    assert [_vop(orc, 1, e, v, 1) for e in (x, y, z)] == [2, 3, 4]
    s = np.sqrt(f32(113.0))
    u = np.array([f32(3) / s, f32(10) / s, f32(2) / s], np.float32)
    rem = np.array(v, np.float32) - _vop(orc, 1, u, v, 1) * u
    assert abs(_vop(orc, 1, rem, u, 1)) < 1e-6
    for k in range(3):
        w = [0.0, 0.0, 0.0]
        w[k] = nan
        assert np.isnan(_vop(orc, 1, u, w, 1))
        w[k] = inf
        assert _vop(orc, 1, u, w, 1) == inf
    v = (3.0, -1.0, 0.0003)
    cs = _vop(orc, 4, v, n=6)
    a, b = cs[:3], cs[3:]
    assert abs(_vop(orc, 1, v, a, 1)) < 1e-6 and abs(_vop(orc, 1, v, b, 1)) < 1e-6 and abs(_vop(orc, 1, a, b, 1)) < 1e-6


def test_point_distances_and_face_forward(orc):
    """geometry/point.rs:196-226 (distance, distance_squared), geometry/normal.rs:415-430 (face_forward)"""
    inf, nan = float("inf"), float("nan")
    o = (0, 0, 0)
    for v, want in (((0, 0, 0), 0.0), ((1, 0, 0), 1.0), ((-1, 0, 0), 1.0), ((1, 1, 1), 3.0), ((0, 2, 0), 4.0),
                    ((0, 0, 2), 4.0), ((inf, 0, 0), inf), ((0, inf, 0), inf), ((0, 0, inf), inf)):
        assert _vop(orc, 8, v, o, 1) == want and _vop(orc, 5, v, o, 1) == np.sqrt(f32(want))
    for k in range(3):
        v = [0.0, 0.0, 0.0]
        v[k] = nan
        assert np.isnan(_vop(orc, 8, v, o, 1)) and np.isnan(_vop(orc, 5, v, o, 1))
    n, minus = [1.0, 0.0, 0.0], [-1.0, -0.0, -0.0]
    ff = lambda v: _vop(orc, 6, n, v).tolist()
    assert ff((-1, -1, -1)) == minus and ff((1, 1, 1)) == n
    for v in ((nan, 1, 1), (1, nan, 1), (1, 1, nan), (inf, 1, 1), (1, inf, 1), (1, 1, inf), (1, -inf, 1), (1, 1, -inf)):
        assert ff(v) == n
    assert ff((-inf, 1, 1)) == minus


# ---- HaltonSampler: montecarlo.rs radical_inverse tests + sampler/halton.rs as written -------------

def test_radical_inverse(orc):
    """montecarlo.rs:167-199 it_can_compute_radical_inverses"""
    ri = orc.lib().orc_radical_inverse
    assert [ri(n, 2) for n in range(8)] == [0.0, 0.5, 0.25, 0.75, 0.125, 0.625, 0.375, 0.875]
    assert [ri(n, 4) for n in range(12)] == [0.0, 0.25, 0.5, 0.75, 1 / 16, 5 / 16, 9 / 16, 13 / 16, 2 / 16, 6 / 16,
                                             10 / 16, 14 / 16]
    for n, want in enumerate([0.0, 1 / 3, 2 / 3, 1 / 9, 4 / 9, 7 / 9, 2 / 9, 5 / 9, 8 / 9]):
        assert abs(ri(n, 3) - want) < 1e-6


def test_halton_sampler_as_written(orc):
    """sampler/halton.rs:17-108 (the reference has no test for it: `mod tests {}`), checked against a
    direct transcription in numpy for one task window, plus the layout invariants the GPU relies on."""
    from pbrt_rust_b200 import scenes
    cfg = scenes.config1(xres=40, yres=24, sampler="halton")     # 2 x 2 -> 4 samples per pixel
    c = orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0)
    cs, _, counts = orc.halton_samples(c)
    lay = orc.layout(c)
    se, nt = lay["sample_ext"], lay["num_tasks"]
    h, w, cap, _ = cs.shape
    valid = ~np.isnan(cs[..., 0])
    assert np.array_equal(valid.sum(axis=2), counts) and cap == counts.max()
    assert (valid[..., :-1] >= valid[..., 1:]).all()            # a pixel's real samples come first
    assert abs(counts.mean() - 4.0) < 0.2 and counts.sum() == valid.sum()
    yy, xx = np.meshgrid(np.arange(se[2], se[3]), np.arange(se[0], se[1]), indexing="ij")
    fx, fy = np.floor(cs[..., 0]), np.floor(cs[..., 1])
    assert (fx[valid] == np.broadcast_to(xx[..., None], fx.shape)[valid]).all()   # binned by home pixel
    assert (fy[valid] == np.broadcast_to(yy[..., None], fy.shape)[valid]).all()
    # transcription of get_more_samples for task 0's window
    ext = np.array([se[0], se[1], se[2], se[3]], np.int32)
    win = np.zeros(4, np.int32)
    orc.lib().orc_compute_sub_window(_p(ext), 0, nt, _p(win))
    x0, x1, y0, y1 = (int(v) for v in win)
    dx, dy = x1 - x0, y1 - y0
    wanted = max(dx, dy) ** 2 * 4
    delta = f32(max(float(dy), float(dx)))
    ri = orc.lib().orc_radical_inverse
    got = {}
    for i in range(wanted):
        u, v = f32(ri(i, 3)), f32(ri(i, 2))
        ix = f32(x0) * (f32(1.0) - u) + (f32(x0) + delta) * u
        iy = f32(y0) * (f32(1.0) - v) + (f32(y0) + delta) * v
        if ix >= f32(x1) or iy >= f32(y1):
            continue
        got.setdefault((int(np.floor(iy)), int(np.floor(ix))), []).append(
            (ix, iy, f32(ri(i + 1, 5)), f32(ri(i + 1, 7)), f32(0.0)))
    assert len(got) > 10
    for (py, px), lst in got.items():
        row = cs[py - se[2], px - se[0]]
        assert counts[py - se[2], px - se[0]] == len(lst)
        assert np.array_equal(row[:len(lst)], np.array(lst, np.float32))


def test_halton_render_modes_agree(orc):
    """default (pixel raster, generation order) vs strict (per-task sub-films) accumulation: the same
    samples and radiance, sums differing by rounding only; every real sample is traced once."""
    import pbrt_rust_b200 as pb
    from pbrt_rust_b200 import scenes
    cfg = scenes.config3(nx=24, nz=12, xres=48, yres=32, xs=2, ys=2)
    e = cfg["film"].get_sample_extent()
    smp = pb.Sampler.halton(e[0], e[1], e[2], e[3], 4, 0.0, 0.0)
    osc = orc.OracleScene(cfg["scene"])
    r0 = orc.render(osc, orc.render_config(cfg["camera"], smp, num_cpus=8, mode=0), want_hits=True)
    r1 = orc.render(osc, orc.render_config(cfg["camera"], smp, num_cpus=8, mode=1))
    cap, counts = orc.halton_cap(orc.render_config(cfg["camera"], smp, num_cpus=8, mode=0))
    assert r0["stats"]["camera_rays"] == counts.sum() == r1["stats"]["camera_rays"]
    assert r0["hit_ids"].size == counts.size * cap
    # strict mode clips samples to the task's sub-film: only task-border pixels may differ
    same = np.abs(r0["rgb"] - r1["rgb"]).max(axis=-1) < 1e-5
    assert same.mean() > 0.9 and r0["rgb"].max() > 0.05


def _film_rgb_of_grey(g):
    """What the film reports for a grey radiance g: rgb_to_xyz (spectrum.rs:37-41) when the sample is
    added, xyz_to_rgb (spectrum.rs:31-35) when the image is read.  As written the two are not
    inverses — the first row of xyz_to_rgb has 1.37150 where pbrt-v2 has 1.537150 — so a grey value
    comes back with R scaled by 1.1657; kept as written (a computed value, SURVEY §0.2 rule)."""
    x, y, z = (0.412453 + 0.357580 + 0.180423) * g, (0.212671 + 0.715160 + 0.072169) * g, (0.019334 + 0.119193 + 0.950227) * g
    return np.array([3.240479 * x - 1.37150 * y - 0.498535 * z, -0.969256 * x + 1.875991 * y + 0.041556 * z,
                     0.055648 * x - 0.204043 * y + 1.057311 * z])


# ---- A13: the oracle-DEFINED area light (the reference has a stub, src/area_light.rs:9-24) pinned to
# physics.  Uniform-area sampling of the emitter with pdf = d^2 / (|cos| A) must converge to the
# closed-form irradiance of a polygonal Lambertian emitter.
@pytest.mark.parametrize("target,light_samples", [((0.0, 0.0, 0.0), 1), ((7.0, 0.0, 3.0), 1), ((-11.0, 0.0, -6.5), 4),
                                                  ((1.5, 0.0, -2.5), 2)])
def test_area_light_estimator_converges_to_lamberts_polygon_formula(orc, target, light_samples):
    from pbrt_rust_b200 import scenes
    cfg = scenes.irradiance_probe(target=target, light_samples=light_samples, res=16, spp=8)   # 2^14 camera samples
    ref = orc.render(orc.OracleScene(cfg["scene"]), orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0))
    got = ref["rgb"].reshape(-1, 3).mean(axis=0)
    e = scenes.polygon_irradiance(target, (0.0, 1.0, 0.0), cfg["light_quad"], 15.0)
    want = _film_rgb_of_grey(0.5 / np.pi * e)
    assert e > 0.05
    assert np.allclose(got, want, rtol=1e-2), (got, want)          # stated tolerance: 1 %


def test_area_light_is_one_sided_and_emits_its_radiance(orc):
    from pbrt_rust_b200 import scenes
    # the emitter faces away from the receiver: no light arrives (cos_l <= 0 -> Li = 0, pdf = 0)
    cfg = scenes.irradiance_probe(emit_down=False, res=8, spp=4)
    ref = orc.render(orc.OracleScene(cfg["scene"]), orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0))
    assert ref["rgb"].max() == 0.0
    # looking up at the emitting side: every pixel is Le = L (plus nothing: an emitter does not light itself)
    cfg = scenes.irradiance_probe(target=(0.0, 8.0, 0.0), eye=(0.5, 1.0, 0.3), res=8, spp=2)
    ref = orc.render(orc.OracleScene(cfg["scene"]), orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0))
    assert np.allclose(ref["rgb"], _film_rgb_of_grey(15.0), rtol=2e-6), (ref["rgb"].min(), ref["rgb"].max())
    # ... and at its back side: black
    cfg = scenes.irradiance_probe(target=(0.0, 8.0, 0.0), eye=(0.5, 1.0, 0.3), emit_down=False, res=8, spp=2)
    ref = orc.render(orc.OracleScene(cfg["scene"]), orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0))
    assert ref["rgb"].max() == 0.0


def test_reference_arm_scene_equals_the_api_scene_and_needs_no_product_code(orc):
    """bench.py --impl reference builds config 3 from oracle/refscene.py (liborc.so + numpy only).  It
    must describe the very same frame as pbrt_rust_b200.scenes.config3, and loading it must not pull
    libpbrtb200.so into the process (the two arms share no code)."""
    import subprocess
    import sys
    from oracle import refscene
    from pbrt_rust_b200 import scenes
    h, c = refscene.config3(nx=60, nz=30, xres=96, yres=64, xs=2, ys=2, mode=0)
    a = orc.render(h, c)
    cfg = scenes.config3(nx=60, nz=30, xres=96, yres=64, xs=2, ys=2)
    b = orc.render(orc.OracleScene(cfg["scene"]), orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8, mode=0))
    assert np.array_equal(a["film"].view(np.uint32), b["film"].view(np.uint32))
    code = ("import sys; sys.path.insert(0, %r)\nfrom oracle import orc, refscene\n"
            "h, c = refscene.config3(nx=8, nz=4, xres=16, yres=16, xs=1, ys=1)\norc.render(h, c)\n"
            "maps = open('/proc/self/maps').read()\nassert 'liborc' in maps and 'libpbrtb200' not in maps and 'pbrt_rust_b200' not in sys.modules\n"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    subprocess.check_call([sys.executable, "-c", code])
