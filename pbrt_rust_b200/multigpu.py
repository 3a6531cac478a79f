"""Tile partition + film gather for 1/2/4/8 GPUs.

The scene is replicated per GPU; the film is cut into fixed 64x64 tiles assigned cyclically
(tile -> gpu = tile_id mod G).  Every GPU evaluates all camera samples whose filter footprint touches
its tiles (halo recompute — free with a counter-addressable RNG), so a film pixel is produced by
exactly one GPU with a summation order that does not depend on G.

Two gathers exist.  `PeerFilm` (the default on GPUs): rank 0 owns the film buffer and shares it by
CUDA IPC; every rank's film kernel stores its owned pixels straight into that buffer over
NVLink/NVSwitch, so the only per-frame communication besides those stores is a barrier.
`gather_film`: a single reduce(SUM) of full-size film buffers whose non-owned pixels are exactly
zero (x + 0 == x), on NCCL (gloo in the CPU tests) — the portable fallback and the CPU-testable
statement of the same result."""
import ctypes as C

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


class PeerFilm:
    """Film buffer on rank 0's GPU that every rank writes its owned tiles into (CUDA IPC + NVLink
    stores from k_film; include/pbrtb200.h pbrtb200_peer_film_*).  One process per GPU."""

    def __init__(self, ctx, n_pixels, dist, device):
        """Collective: every rank must call it.  Never raises between the collectives it issues;
        `self.ok` is False on every rank if any rank could not create or map the buffer."""
        import torch
        from ._ffi import lib
        from .api import DevicePtr
        self.ctx, self.dist = ctx, dist
        self.rank = dist.get_rank()
        self.n_pixels = n_pixels
        self.opened = False
        msg = torch.zeros(65, dtype=torch.uint8, device=device)  # 64 handle bytes + "created" flag
        p = C.c_void_p()
        if self.rank == 0:
            buf = C.create_string_buffer(64)
            if lib().pbrtb200_peer_film_create(ctx.h, n_pixels, C.byref(p), buf) == 0:
                msg[:64].copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
                msg[64] = 1
        dist.broadcast(msg, src=0)
        host = msg.cpu().numpy()
        good = int(host[64]) == 1
        if good and self.rank != 0:
            good = lib().pbrtb200_peer_film_open(ctx.h, bytes(host[:64].tobytes()), C.byref(p)) == 0
            self.opened = good
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = int(flag.item()) == 1
        self.ptr = DevicePtr(p.value or 0)
        if not self.ok:
            self.close()

    def tensor(self):
        """Rank 0: the gathered film as a CUDA tensor view (n_pixels * 4 floats)."""
        import torch
        assert self.rank == 0

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (self.n_pixels * 4,), "typestr": "<f4", "data": (self.ptr.addr, False),
                                      "version": 3, "strides": None}
        return torch.as_tensor(h, device="cuda")

    def close(self):
        from ._ffi import lib
        if self.opened:
            lib().pbrtb200_peer_film_close(self.ctx.h, C.c_void_p(self.ptr.addr))
            self.opened = False


class BandBalancer:
    """Cost-balanced contiguous partition: rank k owns the film rows [b[k], b[k+1]).  Contiguous
    bands keep each GPU's rays in one part of the scene (the cyclic tiles make every GPU touch the
    whole BVH with 1/G of the rays, which costs the closest-hit kernel ~20 %), and the boundaries
    follow the measured per-rank frame time of the previous frame(s), so sky / horizon / dense
    regions end up with equal work.  Any partition gives the same film (ownership is disjoint and
    the per-pixel summation order does not depend on it), so rebalancing between frames is free of
    side effects; only the pixel work list is rebuilt when the boundaries move."""

    def __init__(self, pixel_extent, world, quantum=4):
        self.x0, self.x1, self.y0, self.y1 = pixel_extent
        self.world, self.q = world, quantum
        h = self.y1 - self.y0
        self.b = [self.y0 + self._snap(round(k * h / world)) for k in range(world)] + [self.y1]
        self._fix()

    def _snap(self, v):
        return int(round(v / self.q)) * self.q

    def _fix(self):
        """Boundaries non-decreasing inside [y0, y1]; at least one quantum per rank when the film has
        that many rows (a film with fewer rows than ranks leaves the last ranks an EMPTY band: they
        get an empty tile set and render nothing)."""
        self.b[0], self.b[-1] = self.y0, self.y1
        for k in range(1, self.world):
            self.b[k] = max(self.b[k], self.b[k - 1] + self.q)
        for k in range(self.world - 1, 0, -1):
            self.b[k] = min(self.b[k], self.b[k + 1] - 1)
        for k in range(1, self.world):  # short films: clamp instead of running past either end
            self.b[k] = min(max(self.b[k], self.b[k - 1], self.y0), self.y1)

    def tiles_for(self, rank):
        a, c = self.b[rank], self.b[rank + 1]
        return [(self.x0, a, self.x1, c)] if c > a else []

    def imbalance(self, times):
        return max(times) / (sum(times) / len(times))

    def update(self, times, damping=0.8):
        """times[k] = last frame time of rank k.  Moves the boundaries assuming a piecewise-constant
        cost density per band; returns True if they changed."""
        dens = [t / max(1, self.b[k + 1] - self.b[k]) for k, t in enumerate(times)]
        total = sum(times)
        target = total / self.world
        new_b, k, acc, y = [self.y0], 0, 0.0, float(self.y0)
        for r in range(1, self.world):
            need = target * r
            while k < self.world and acc + dens[k] * (self.b[k + 1] - y) < need:
                acc += dens[k] * (self.b[k + 1] - y)
                y = float(self.b[k + 1])
                k += 1
            if k >= self.world:
                new_b.append(self.y1)
                continue
            yy = y + (need - acc) / max(dens[k], 1e-12)
            new_b.append(yy)
        new_b.append(self.y1)
        old = list(self.b)
        self.b = [self.y0] + [self.y0 + self._snap(old[i] + damping * (new_b[i] - old[i]) - self.y0)
                              for i in range(1, self.world)] + [self.y1]
        self._fix()
        return self.b != old
