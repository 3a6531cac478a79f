"""N>1 host logic on CPU: world_size-2 gloo run of the tile partition + film gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pbrt_rust_b200 import multigpu


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("ext", [(0, 1920, 0, 1080), (48, 95, 4, 8), (0, 70, 0, 130)])
def test_tiles_partition_the_film_exactly_once(world, ext):
    cov = multigpu.coverage(ext, world)
    assert cov.min() == 1 and cov.max() == 1
    sizes = [sum((c - a) * (d - b) for a, b, c, d in multigpu.partition_tiles(ext, r, world)) for r in range(world)]
    assert sum(sizes) == (ext[1] - ext[0]) * (ext[3] - ext[2])
    if ext[1] - ext[0] >= 1024:
        assert max(sizes) / (sum(sizes) / world) < 1.1      # near-even static split


def _worker(rank, world, port, ext, ref_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = torch.from_numpy(np.load(ref_path))
    x0, x1, y0, y1 = ext
    mine = torch.zeros_like(ref)
    for (a, b, c, d) in multigpu.partition_tiles(ext, rank, world, tile=16):
        mine[b - y0:d - y0, a - x0:c - x0] = ref[b - y0:d - y0, a - x0:c - x0]
    flat = mine.reshape(-1).contiguous()
    multigpu.gather_film(flat, dist, dst=0)
    if rank == 0:
        np.save(out_path, flat.reshape(ref.shape).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_is_bit_exact(tmp_path, orc):
    """Each rank holds only its tiles of an (oracle-rendered) film; the reduce must rebuild the
    whole film bit for bit — the property the NCCL gather in bench.py relies on."""
    from pbrt_rust_b200 import scenes
    cfg = scenes.config1(xres=80, yres=60)
    osc = orc.OracleScene(cfg["scene"])
    ref = orc.render(osc, orc.render_config(cfg["camera"], cfg["sampler"], num_cpus=8))["film"]
    ext = cfg["film"].get_pixel_extent()
    ref_path, out_path = str(tmp_path / "ref.npy"), str(tmp_path / "out.npy")
    np.save(ref_path, ref)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, ext, ref_path, out_path), nprocs=2, join=True)
    out = np.load(out_path)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_band_balancer_covers_rows_and_converges(world):
    """multigpu.BandBalancer: bands partition the film rows exactly once for any boundaries, and
    the feedback rule equalises a skewed cost profile within a few frames."""
    ext = (0, 1920, 0, 1080)
    b = multigpu.BandBalancer(ext, world)
    dens = np.where(np.arange(1080) < 200, 2.5, 1.0) + 0.3 * np.sin(np.arange(1080) / 90.0)
    for _ in range(10):
        cov = np.zeros(1080, np.int32)
        for k in range(world):
            for (x0, y0, x1, y1) in b.tiles_for(k):
                assert (x0, x1) == (0, 1920) and y1 > y0
                cov[y0:y1] += 1
        assert cov.min() == 1 and cov.max() == 1
        times = [float(dens[b.b[k]:b.b[k + 1]].sum()) for k in range(world)]
        if b.imbalance(times) < 1.02 or not b.update(times):
            break
    assert b.imbalance(times) < 1.05
    # films shorter than the number of ranks still give every rank a non-empty band where possible
    tiny = multigpu.BandBalancer((3, 9, 5, 5 + world + 1), world)
    rows = sorted(r for k in range(world) for (_, y0, _, y1) in tiny.tiles_for(k) for r in range(y0, y1))
    assert rows == list(range(5, 5 + world + 1))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("ext", [(0, 1920, 0, 1080), (0, 64, 0, 5), (3, 40, 7, 8), (0, 16, 0, 33)])
def test_band_balancer_always_covers_the_film_exactly_once(world, ext):
    """Row bands: boundaries stay inside the film and non-decreasing, whatever the measured times
    do; a film with fewer rows than ranks leaves some ranks an EMPTY tile list (which the C ABI
    renders as "nothing", never as "the whole film")."""
    bal = multigpu.BandBalancer(ext, world)
    rng = np.random.default_rng(world * 1000 + ext[3])
    for step in range(12):
        b = bal.b
        assert b[0] == ext[2] and b[-1] == ext[3] and all(b[i] <= b[i + 1] for i in range(world))
        cov = np.zeros((ext[3] - ext[2], ext[1] - ext[0]), np.int32)
        for r in range(world):
            for (a, y0, c, y1) in bal.tiles_for(r):
                assert ext[0] <= a < c <= ext[1] and ext[2] <= y0 < y1 <= ext[3]
                cov[y0 - ext[2]:y1 - ext[2], a - ext[0]:c - ext[0]] += 1
        assert cov.min() == 1 and cov.max() == 1
        bal.update(list(rng.random(world) * 3 + 0.01))


@pytest.mark.parametrize("n_bands", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("rows,y0", [(1080, 0), (150, 7), (5, 0), (33, -2)])
def test_group_band_cut_covers_the_film_and_balances_cost(n_bands, rows, y0):
    """pbrtb200_cut_bands (the host arithmetic inside pbrtb200_group_render): contiguous bands that
    cover every row once, boundaries on multiples of 4 rows, summed cost per band within one quantum
    of the ideal share — for flat, ramped and spiky cost profiles, more bands than rows included."""
    import ctypes as C
    from pbrt_rust_b200 import _ffi
    L = _ffi.lib()
    rng = np.random.default_rng(rows * 10 + n_bands)
    for cost in (np.ones(rows), np.linspace(1.0, 9.0, rows), rng.random(rows) ** 4 + 0.01, np.zeros(rows)):
        cost = np.ascontiguousarray(cost, np.float32)
        b = (C.c_int32 * (n_bands + 1))()
        assert L.pbrtb200_cut_bands(cost.ctypes.data_as(C.POINTER(C.c_float)), rows, y0, n_bands, b) == 0
        b = list(b)
        assert b[0] == y0 and b[-1] == y0 + rows and all(x <= y for x, y in zip(b, b[1:]))
        assert all((x - y0) % 4 == 0 for x in b[1:-1] if x != y0 + rows)
        if cost.sum() > 0 and rows >= 16 * n_bands:
            share = [cost[b[k] - y0:b[k + 1] - y0].sum() for k in range(n_bands)]
            ideal = cost.sum() / n_bands
            slack = 4 * cost.max() + 1e-6          # one quantum of the most expensive rows on either side
            assert max(abs(s_ - ideal) for s_ in share) <= 2 * slack, (share, ideal)


def _simulate_bands(n, seed, frames=40, noise=0.01, change_at=None):
    """Drives pbrtb200_bands_* with a hidden per-row cost (a horizon peak the probe only half sees), a
    fixed per-device overhead and per-frame timing noise.  Returns per frame (max / mean of the
    device times, moved?)."""
    import ctypes as C
    from pbrt_rust_b200 import _ffi
    L = _ffi.lib()
    rng = np.random.default_rng(seed)
    H, y0 = 1080, 0
    y = np.arange(H)
    true = 1 + 4 * np.exp(-((y - 300) / 60.0) ** 2) + 0.5 * (y > 300)
    true *= 7.9 / true.sum()
    probe = np.ascontiguousarray(true * (1 + 0.3 * np.sin(y / 50.0)), np.float32)
    bal = L.pbrtb200_bands_new(probe.ctypes.data_as(C.POINTER(C.c_float)), H, y0, n)
    assert bal
    try:
        b = (C.c_int32 * (n + 1))()
        hist = []
        for f in range(frames):
            assert L.pbrtb200_bands_get(bal, b) == 0
            bb = list(b)
            assert bb[0] == y0 and bb[-1] == y0 + H and all(p <= q for p, q in zip(bb, bb[1:]))
            extra = np.zeros(n)
            if change_at is not None and f >= change_at:   # e.g. host film: half the devices pay 15 % more
                extra[: n // 2] = 0.15
            t = np.array([0.15 + true[bb[k]:bb[k + 1]].sum() for k in range(n)]) * (1 + extra)
            t = np.ascontiguousarray(t * (1 + noise * rng.standard_normal(n)), np.float32)
            moved = L.pbrtb200_bands_update(bal, t.ctypes.data_as(C.POINTER(C.c_float)))
            assert moved in (0, 1)
            hist.append((float(t.max() / t.mean()), moved))
        return hist
    finally:
        L.pbrtb200_bands_free(bal)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_band_balancer_converges_and_then_holds_still(n):
    """The balancer of pbrtb200_group_render (pbrtb200_bands_*): from a half-wrong cost probe it gets
    the slowest device within ~3 % of the mean in a few frames, and once settled it does not move
    again under 1 % timing noise (a move costs the devices a pixel-list rebuild)."""
    for seed in range(4):
        hist = _simulate_bands(n, seed)
        assert max(r for r, _ in hist[8:]) <= 1.06, (seed, hist)     # settled within 3 %, plus noise: below the 6 % that would move them
        assert sum(m for _, m in hist) <= 8, hist                    # a handful of moves in total
        assert sum(m for _, m in hist[12:]) == 0, hist               # none once settled


def test_band_balancer_follows_a_real_change_but_not_noise():
    """Settled bands start a new round only when the devices stay more than 6 % apart for three frames
    (here: half the devices become 15 % slower at frame 20) — and settle again."""
    hist = _simulate_bands(8, 11, frames=60, change_at=20)
    assert sum(m for _, m in hist[12:20]) == 0
    assert hist[20][0] > 1.06 and sum(m for _, m in hist[20:30]) >= 1
    assert max(r for r, _ in hist[34:]) <= 1.06 and sum(m for _, m in hist[40:]) == 0, hist
    # bad arguments
    import ctypes as C
    from pbrt_rust_b200 import _ffi
    L = _ffi.lib()
    assert not L.pbrtb200_bands_new(None, 0, 0, 2)
    assert L.pbrtb200_bands_update(None, None) < 0
    one = L.pbrtb200_bands_new(None, 10, 5, 1)    # uniform cost, one band: nothing to balance
    b = (C.c_int32 * 2)()
    t = (C.c_float * 1)(1.0)
    assert L.pbrtb200_bands_update(one, t) == 0 and L.pbrtb200_bands_get(one, b) == 0 and list(b) == [5, 15]
    L.pbrtb200_bands_free(one)
