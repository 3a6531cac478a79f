#!/bin/bash
# Round 2, run 6: any-hit with postponed leaves, loop shapes, block sizes.
mkdir -p gpurun_out
O=gpurun_out/r2_ab6.txt
: > $O
line() {
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
line "lib=default"
for v in any_m4 any_m3 clo_m0 t64 t256; do
  PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so line "lib=$v"
done
PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_any_m4.so python -m pytest tests/test_gpu_parity.py -x -q -k "trace_any or config3 or fuzz or area_light" > gpurun_out/r2_tests6.log 2>&1
echo "gpu tests with any_m4: $(tail -1 gpurun_out/r2_tests6.log)" >> $O
cat $O
PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_any_m4.so ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:k_trace -s 2 -c 2 --csv --log-file gpurun_out/r2_ncu6_m4.csv python scripts/prof_frame.py 2 > gpurun_out/r2_ncu6_m4.log 2>&1
grep -v "^==" gpurun_out/r2_ncu6_m4.csv | cut -d, -f5,13- | tail -14
