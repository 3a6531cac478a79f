#!/bin/bash
# Round 2, run 5: full GPU suite on the radiance-record / pixel-ring / group build, then loop-shape and block-size variants.
mkdir -p gpurun_out
O=gpurun_out/r2_ab5.txt
: > $O
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1
echo "gpu tests: $(tail -1 gpurun_out/r2_tests5.log)" >> $O
tail -30 gpurun_out/r2_tests5.log
line() {
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
line "lib=default"
for v in any_m3 any_m0 clo_m0 t64 t256; do
  PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so line "lib=$v"
done
cat $O
