#!/bin/bash
PBRTB200_SHADOW_MODE=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('ms/frame %.3f | raygen %.2f trace %.2f shade %.2f shadow %.2f film %.2f | launches/frame %d' % (d['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['gpu_launches']/d['steps']))"; }
run PBRTB200_SHADOW_MODE=0
run PBRTB200_SHADOW_MODE=2
run PBRTB200_SHADOW_MODE=0
run PBRTB200_SHADOW_MODE=2
