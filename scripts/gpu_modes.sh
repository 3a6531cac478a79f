#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for tm in 0 1 2; do for sm in 0 1 2; do
  echo "== TRACE_MODE=$tm SHADOW_MODE=$sm"
  PBRTB200_TRACE_MODE=$tm PBRTB200_SHADOW_MODE=$sm python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stage_ms_per_frame']
print('ms/frame %.3f trace %.3f shadow %.3f' % (d['ms_per_step'], s['ms_trace'], s['ms_shadow']))"
done; done
PBRTB200_TRACE_MODE=2 PBRTB200_SHADOW_MODE=2 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
