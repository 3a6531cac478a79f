"""N-rank check of the P2P film gather (run under torchrun on a multi-GPU box): the film gathered
in rank 0's HBM by every rank's k_film storing its owned tiles over NVLink must be bit-identical to
(a) the NCCL reduce(SUM) gather and (b) a single-GPU whole-film render."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pbrt_rust_b200 as pb
from pbrt_rust_b200 import scenes, multigpu

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = scenes.config3(nx=200, nz=100, xres=640, yres=360, xs=2, ys=2)
film = cfg["film"]; h, w = film.shape
r = pb.GpuRenderer(cfg["sampler"], cfg["camera"], cfg["integrator"], num_cpus=8, device=local)
tiles = multigpu.partition_tiles(film.get_pixel_extent(), rank, world)
peer = multigpu.PeerFilm(r.ctx, h * w, dist, torch.device("cuda", local))
assert peer.ok, "CUDA IPC film sharing unavailable"
for _ in range(3):  # repeated frames: ownership is disjoint, so re-writing is idempotent
    r.render(cfg["scene"], tiles=tiles, out=peer.ptr, keep_others=True)
    dist.barrier()
d = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
r.render(cfg["scene"], tiles=tiles, out=d)
dist.reduce(d, dst=0, op=dist.ReduceOp.SUM)
if rank == 0:
    p2p = peer.tensor().cpu().numpy()
    nccl = d.cpu().numpy()
    whole = r.render(cfg["scene"]).reshape(-1)
    ok1, ok2 = np.array_equal(p2p.view(np.uint32), nccl.view(np.uint32)), np.array_equal(p2p.view(np.uint32), whole.view(np.uint32))
    print(f"multigpu_check world={world}: p2p==nccl {ok1}  p2p==single-GPU {ok2}  weight-sum min {p2p[3::4].min():.3f}")
    assert ok1 and ok2
dist.barrier()
peer.close()
dist.destroy_process_group()
