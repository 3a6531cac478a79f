"""Procedural geometry of the benchmark scenes: plain numpy, no package-relative imports,
so that the CPU reference arm of bench.py can load this FILE by path without importing the product
package (whose import loads libpbrtb200.so).  The generator PRNG is SplitMix64 (ours; unrelated to the
renderer's StdRng)."""
import numpy as np


def splitmix64(seed, n):
    """n uniform floats in [0,1) from SplitMix64(seed), vectorised."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def random_triangles(n=100_000, seed=1):
    u = splitmix64(seed, n * 12).reshape(n, 12)
    c = u[:, 0:3] * 20.0 - 10.0
    off = (u[:, 3:12] * 0.6 - 0.3).reshape(n, 3, 3)
    P = (c[:, None, :] + off).astype(np.float32).reshape(-1, 3)
    vi = np.arange(3 * n, dtype=np.uint32)
    return vi, P


def heightfield(nx, nz, x0=-20.0, x1=20.0, z0=-10.0, z1=10.0):
    """(nx x nz) cells, 2 triangles per cell: y = 0.6 sin(0.9x) cos(1.1z) + 0.15 hash(i,j)."""
    i = np.arange(nx + 1, dtype=np.float64)
    j = np.arange(nz + 1, dtype=np.float64)
    X, Z = np.meshgrid(x0 + (x1 - x0) * i / nx, z0 + (z1 - z0) * j / nz, indexing="ij")
    ii, jj = np.meshgrid(np.arange(nx + 1, dtype=np.uint64), np.arange(nz + 1, dtype=np.uint64), indexing="ij")
    with np.errstate(over="ignore"):
        h = (ii * np.uint64(73856093)) ^ (jj * np.uint64(19349663))
        h = (h ^ (h >> np.uint64(13))) * np.uint64(0x9E3779B97F4A7C15)
        hv = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    Y = 0.6 * np.sin(0.9 * X) * np.cos(1.1 * Z) + 0.15 * hv * (40.0 / nx)
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32).reshape(-1, 3)
    a = (np.arange(nx)[:, None] * (nz + 1) + np.arange(nz)[None, :]).astype(np.uint32)
    b, c, d = a + (nz + 1), a + 1, a + (nz + 1) + 1
    vi = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], axis=2).reshape(-1).astype(np.uint32)
    return vi, P


def quad_light(y=8.0, half=2.0, cx=0.0, cz=0.0):
    P = np.array([[cx - half, y, cz - half], [cx + half, y, cz - half], [cx + half, y, cz + half],
                  [cx - half, y, cz + half]], np.float32)
    # winding chosen so that dg.nn (= normalize((p2-p1) x (p3-p2)) after the refine reversal,
    # mesh.rs:220-262 with default uvs) points down (-y): the quad emits towards the ground.
    vi = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    return vi, P


