#!/bin/bash
# A/B: k_shade occupancy target after the per-block atomic aggregation
for occ in 8 10 12; do
echo "== SHADE_OCC=$occ"
PBRTB200_SHADE_OCC=$occ python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('ms/frame %.3f e2e %.0f | raygen %.2f trace %.2f shade %.2f shadow %.2f film %.2f' % (d['ms_per_step'], d['e2e']['value'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film']))"
done
