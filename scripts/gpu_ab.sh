#!/bin/bash
# A/B: work-counter claim size of the persistent trace kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for b in 1 2 4 8; do
echo "== TRACE_BATCH=$b"
PBRTB200_TRACE_BATCH=$b python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('ms/frame %.3f | raygen %.2f trace %.2f shade %.2f shadow %.2f film %.2f' % (d['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film']))"
done
