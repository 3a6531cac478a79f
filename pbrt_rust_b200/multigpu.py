"""Tile partition + film gather for 1/2/4/8 GPUs.

The scene is replicated per GPU; the film is cut into fixed 64x64 tiles assigned cyclically
(tile -> gpu = tile_id mod G).  Every GPU evaluates all camera samples whose filter footprint touches
its tiles (halo recompute — free with a counter-addressable RNG), so a film pixel is produced by
exactly one GPU with a summation order that does not depend on G.  The gather is a single
reduce(SUM) of full-size film buffers whose non-owned pixels are exactly zero (x + 0 == x), issued
on NCCL over NVLink/NVSwitch (gloo in the CPU tests)."""
import numpy as np


def partition_tiles(pixel_extent, rank, world, tile=64):
    """Film pixel rects (x0, y0, x1, y1) owned by `rank`."""
    x0, x1, y0, y1 = pixel_extent
    rects, tid = [], 0
    for ty in range(y0, y1, tile):
        for tx in range(x0, x1, tile):
            if tid % world == rank:
                rects.append((tx, ty, min(tx + tile, x1), min(ty + tile, y1)))
            tid += 1
    return rects


def coverage(pixel_extent, world, tile=64):
    """How many ranks own each film pixel (must be exactly 1 everywhere)."""
    x0, x1, y0, y1 = pixel_extent
    cov = np.zeros((y1 - y0, x1 - x0), np.int32)
    for r in range(world):
        for (a, b, c, d) in partition_tiles(pixel_extent, r, world, tile):
            cov[b - y0:d - y0, a - x0:c - x0] += 1
    return cov


def gather_film(film_tensor, dist, dst=0):
    """In-place reduce(SUM) of the per-rank films onto `dst`; exact because ownership is disjoint."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor
