#!/bin/bash
# Round 2, run 4: per-kernel stack depth / dispatch variants (built from the traversal snapshot), box 2.
mkdir -p gpurun_out
O=gpurun_out/r2_ab4.txt
: > $O
line() {
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
for v in base4 tri_all tri_none any_s8 clo_s0 lb10 lb14; do
  for box in 2 3; do
    PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so PBRTB200_BOX=$box line "lib=$v box=$box"
  done
done
cat $O
