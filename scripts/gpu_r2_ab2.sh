#!/bin/bash
# Round 2, run 2: shared-stack depth x register cap x box family; ncu --set full of the default build.
mkdir -p gpurun_out
O=gpurun_out/r2_ab2.txt
: > $O
line() {
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
for box in 2 3; do PBRTB200_BOX=$box line "lib=default(s24,lb9/10) box=$box"; done
for v in s8_lb9 s8_lb10 s8_lb12 s12_lb9 s12_lb10 s16_lb9 s16_lb10; do
  for box in 2 3; do
    PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so PBRTB200_BOX=$box line "lib=$v box=$box"
  done
done
cat $O
for box in 2 3; do
PBRTB200_BOX=$box ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 2 -f -o gpurun_out/r2_prof_trace_box$box \
    python scripts/prof_frame.py 2 > gpurun_out/r2_prof_trace_box$box.log 2>&1
done
ls -la gpurun_out | tail -8
