#!/bin/bash
# N-GPU validation of pbrtb200_group_render: parity test (film bit-identical to one GPU), then bench at N (torchrun) and N=1.
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -k "group_render or cost_profile" > gpurun_out/r2_group_tests_n$N.log 2>&1
tail -5 gpurun_out/r2_group_tests_n$N.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_scale_$n.json').read().strip().splitlines()[-1])
    print('N=$n value %.0f Mrays/s  ms %.3f | e2e %.0f ms %.3f | bands %s | per-device ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['band_rows'], d['config']['per_device_ms']))
except Exception as e:
    print('N=$n failed', e); print(open('gpurun_out/r2_scale_$n.err').read()[-1500:])
PY
  fi
done
