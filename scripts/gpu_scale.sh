#!/bin/bash
# scaling run on one multi-GPU box: P2P gather check + bench at each N given
mkdir -p gpurun_out
for N in "$@"; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/multigpu_check.py 2>&1 | grep "multigpu_check"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu 2> gpurun_out/bench_n${N}.err | tail -1 > gpurun_out/bench_n${N}.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}.json"))
print("N=${N}: ms/frame %.3f  Mrays/s %.0f  e2e %.0f (%.3f ms)  stages %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms_per_frame"].items()}))
PY
done
