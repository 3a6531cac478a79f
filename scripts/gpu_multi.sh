#!/bin/bash
# usage: gpu_multi.sh N [N ...]   (on a box with >= max N GPUs): P2P gather check + bench in both gather modes
mkdir -p gpurun_out
for N in "$@"; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5
for mode in p2p nccl; do
  PBRTB200_GATHER=$mode timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu 2> gpurun_out/bench_n${N}_$mode.err | tail -1 > gpurun_out/bench_n${N}_$mode.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}_$mode.json"))
print("N=${N} $mode: ms/frame %.3f  Mrays/s %.0f  e2e %.0f (%.3f ms)  gather=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["film_gather"][:40]))
PY
  grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_$mode.err | tail -3
done
done
