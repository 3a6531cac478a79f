#!/bin/bash
# BASELINE configs 2, 4, 5 at named size on one GPU (throughput + parity), then the sanitizer pass.
mkdir -p gpurun_out
for c in c2 c4 c5; do
  timeout 900 python scripts/run_configs.py $c > gpurun_out/r2_config_$c.log 2>&1
  tail -c 1800 gpurun_out/r2_config_$c.log; echo
done
for c in c1 c2 c4 c5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_$c.json').read().strip().splitlines()[-1])
    print('$c value %.0f Mrays/s ms %.3f | e2e %.0f ms %.3f | launches %d | %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], {k: round(v,2) for k,v in d['stage_ms_per_frame'].items()}))
except Exception as e:
    print('$c failed', e); print(open('gpurun_out/r2_bench_$c.err').read()[-1200:])
PY
done
bash scripts/gpu_sanitize.sh
