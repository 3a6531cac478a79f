"""Lock-step warp model of the traversal kernels on config 3 (analysis tool; CPU only).

  python scripts/simt_cost.py [tiles]

Samples 8x4-pixel tiles of the 1080p frame (16 spp), feeds their camera rays — in the kernels' packet
order — through scripts/micro/simt_cost.cpp, then does the same for the shadow rays those hits generate,
and prints rounds / active lanes per 32-ray packet plus the bottom-up any-hit estimate."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "scripts", "micro", "simt_cost.cpp")
LIB = os.path.join(ROOT, "scripts", "micro", "libsimt_cost.so")


def lib():
    deps = [SRC] + [os.path.join(ROOT, "pbrt_rust_b200", "csrc", f) for f in ("trace_core.cuh", "trace_math.cuh", "host_logic.hpp")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.simt_tree.restype = C.c_void_p
    L.simt_tree.argtypes = [C.c_void_p]
    L.simt_run.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.simt_bottom_up.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.simt_postponed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


NAMES = ["node_rounds", "node_lanes", "leaf_rounds", "leaf_lanes", "push_rounds", "pop_rounds", "pop_iters", "box_tests",
         "tri_tests", "packets", "rays", "iters"]


def report(tag, c):
    d = dict(zip(NAMES, c))
    pk = d["packets"]
    print(f"{tag}: per packet: node rounds {d['node_rounds'] / pk:.1f} (lanes {d['node_lanes'] / max(1, d['node_rounds']):.1f}), "
          f"leaf rounds {d['leaf_rounds'] / pk:.1f} (lanes {d['leaf_lanes'] / max(1, d['leaf_rounds']):.1f}), push rounds "
          f"{d['push_rounds'] / pk:.1f}, pop rounds {d['pop_rounds'] / pk:.1f}, pop iters {d['pop_iters'] / pk:.1f}; per ray: "
          f"pair steps {d['box_tests'] / 2 / d['rays']:.1f}, tri tests {d['tri_tests'] / d['rays']:.2f}")
    return d


def main():
    n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    import bench
    from pbrt_rust_b200.api import HostScene
    cfg = bench.make_cfg()
    hs = HostScene(cfg["scene"])
    L = lib()
    tree = L.simt_tree(C.byref(hs.flat.contents))
    cam = cfg["camera"].desc
    r2c = np.array(list(cam.raster_to_camera), np.float64).reshape(4, 4)
    c2w = np.array(list(cam.camera_to_world), np.float64).reshape(4, 4)
    rng = np.random.default_rng(7)
    tx = rng.integers(0, 1920 // 8, n_tiles)
    ty = rng.integers(0, 1080 // 4, n_tiles)
    rays = []
    for a, b in zip(tx, ty):
        for yy in range(4):
            for xx in range(8):
                sx, sy = np.meshgrid(np.arange(4), np.arange(4), indexing="xy")
                u = rng.random((2, 16))
                ix = a * 8 + xx + (sx.reshape(-1) + u[0]) / 4
                iy = b * 4 + yy + (sy.reshape(-1) + u[1]) / 4
                pr = np.stack([ix, iy, np.zeros(16), np.ones(16)], 0)
                pc = r2c @ pr
                pc = pc[:3] / pc[3]
                d = pc / np.linalg.norm(pc, axis=0)
                dw = (c2w[:3, :3] @ d).T
                ow = np.broadcast_to(c2w[:3, 3], dw.shape)
                r = np.zeros((16, 8), np.float32)
                r[:, 0:3], r[:, 3], r[:, 4:7], r[:, 7] = ow, 0.0, dw, 3.4028235e38
                rays.append(r)
    rays = np.ascontiguousarray(np.concatenate(rays))
    n = rays.shape[0]
    prim = np.zeros(n, np.uint32)
    t = np.zeros(n, np.float32)
    c = np.zeros(12)
    L.simt_run(tree, _p(rays), n, 0, _p(prim), _p(t), _p(c))
    report("closest (while-while)", c)
    hit = prim != 0xFFFFFFFF
    print(f"  hit fraction {hit.mean():.3f}")
    # shadow rays of the hits, queue order = sample order
    o = rays[hit, 0:3].astype(np.float64) + rays[hit, 4:7].astype(np.float64) * t[hit, None]
    m = o.shape[0]
    ps = np.stack([rng.random(m) * 4 - 2, np.full(m, 8.0), rng.random(m) * 4 - 2], 1)
    dist = np.linalg.norm(ps - o, axis=1)
    sh = np.zeros((m, 8), np.float32)
    sh[:, 0:3], sh[:, 3], sh[:, 4:7], sh[:, 7] = o, t[hit] * 5e-4, (ps - o) / dist[:, None], (1 - 1e-3) * dist
    sh = np.ascontiguousarray(sh)
    c = np.zeros(12)
    sprim = np.zeros(m, np.uint32)
    L.simt_run(tree, _p(sh), m, 1, _p(sprim), None, _p(c))
    report("any-hit (if-if, unordered)", c)
    occ_td = sprim != 0xFFFFFFFF
    print(f"  occluded fraction {occ_td.mean():.3f}")
    inv = 1.0 / sh[:, 4:7]
    oct_ = (inv[:, 0] < 0).astype(int) | ((inv[:, 1] < 0).astype(int) << 1) | ((inv[:, 2] < 0).astype(int) << 2)
    w = oct_[: (m // 32) * 32].reshape(-1, 32)
    print(f"  shadow warps with one octant: {(w.min(1) == w.max(1)).mean():.3f};  x uniform "
          f"{((inv[:(m // 32) * 32, 0] < 0).reshape(-1, 32).std(1) == 0).mean():.3f}, z uniform "
          f"{((inv[:(m // 32) * 32, 2] < 0).reshape(-1, 32).std(1) == 0).mean():.3f}")
    for slots in (1, 2):
        for thr in (8, 16, 24, 32):
            c8 = np.zeros(8)
            occ2 = np.zeros(m, np.uint8)
            L.simt_postponed(tree, _p(sh), m, thr, slots, _p(occ2), _p(c8))
            pk = c8[4]
            print(f"any-hit postponed leaves (slots {slots}, vote threshold {thr}): per packet: node rounds {c8[0] / pk:.1f} (lanes "
                  f"{c8[1] / max(1, c8[0]):.1f}), leaf rounds {c8[2] / pk:.1f} (lanes {c8[3] / max(1, c8[2]):.1f}), tri tests per ray "
                  f"{c8[7] / c8[5]:.2f}; same answers: {np.array_equal(occ2.astype(bool), occ_td)}")
    c4 = np.zeros(4)
    occ = np.zeros(m, np.uint8)
    fp = np.ascontiguousarray(prim[hit])
    L.simt_bottom_up(tree, _p(sh), _p(fp), m, _p(occ), _p(c4))
    print(f"bottom-up any-hit: per ray: box tests {c4[0] / m:.1f}, tri tests {c4[1] / m:.2f}, dependent node loads {c4[2] / m:.1f}, "
          f"fallbacks {int(c4[3])}; agrees with top-down: {np.array_equal(occ.astype(bool), occ_td)}")


if __name__ == "__main__":
    main()
