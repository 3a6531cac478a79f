#!/bin/bash
# partition A/B on one multi-GPU box: bands (balanced) vs cyclic tiles at each N given
mkdir -p gpurun_out
for N in "$@"; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for part in bands cyclic; do
PBRTB200_PARTITION=$part timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu 2> gpurun_out/bench_n${N}_$part.err | tail -1 > gpurun_out/bench_n${N}_$part.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_$part.json"))
    print("N=${N} $part: ms/frame %.3f  Mrays/s %.0f  e2e %.0f (%.3f ms)  per-rank device ms %s  bands %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], [round(x, 3) for x in d["config"]["per_rank_device_ms"]], d["config"]["band_rows"]))
except Exception as e:
    print("N=${N} $part FAILED", e); print(open("gpurun_out/bench_n${N}_$part.err").read()[-1500:])
PY
done
done
