#!/bin/bash
for lib in libpbrtb200 lib_shade256 lib_film12 lib_film16 libpbrtb200; do
echo "== $lib"
PBRTB200_LIB_PATH=$PWD/pbrt_rust_b200/$lib.so python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('ms/frame %.3f e2e %.3f ms | raygen %.2f trace %.2f shade %.3f shadow %.2f film %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film']))"
done
