#!/bin/bash
# Round 2, run 1: parity gate for the new traversal families, then A/B of box family x register cap.
mkdir -p gpurun_out
O=gpurun_out/r2_ab1.txt
: > $O
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1
echo "gpu tests (BOX=3 default): $(tail -1 gpurun_out/r2_tests1.log)" >> $O
line() {  # label, env...
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
for rep in 1 2; do
  for box in 1 2 3; do
    PBRTB200_BOX=$box line "lib=default(lb9/10) box=$box"
  done
  for v in lb8 nolb; do
    for box in 2 3; do
      PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so PBRTB200_BOX=$box line "lib=$v box=$box"
    done
  done
done
# instruction counts and lane utilisation of the trace kernels (second frame), box 1 vs 3
for box in 1 3; do
  PBRTB200_BOX=$box ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none -k regex:k_trace -s 2 -c 2 --csv --log-file gpurun_out/r2_ncu_box$box.csv python scripts/prof_frame.py 2 > gpurun_out/r2_ncu_box$box.log 2>&1
done
cat $O
