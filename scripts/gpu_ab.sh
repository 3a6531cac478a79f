#!/bin/bash
for lg in 23 24 25 23 24; do
echo "== CHUNK_LOG2=$lg"
PBRTB200_CHUNK_LOG2=$lg python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('ms/frame %.3f e2e %.3f ms | raygen %.2f trace %.2f shade %.2f shadow %.2f film %.2f | launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['gpu_launches']))"
done
