#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
PBRTB200_TRACE_MODE=3 PBRTB200_SHADOW_MODE=3 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('ms/frame %.3f | raygen %.2f trace %.2f shade %.2f shadow %.2f film %.2f | launches/frame %d' % (d['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['gpu_launches']/d['steps']))"; }
run A=1
run A=1
run PBRTB200_SHADOW_MODE=3
run PBRTB200_TRACE_MODE=3
run PBRTB200_TRACE_MODE=3 PBRTB200_SHADOW_MODE=3
run PBRTB200_CHUNK_LOG2=21
run PBRTB200_CHUNK_LOG2=23
run PBRTB200_CHUNK_LOG2=24
run PBRTB200_CHUNK_LOG2=26
