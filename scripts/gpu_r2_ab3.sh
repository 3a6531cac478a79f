#!/bin/bash
# Round 2, run 3: per-axis tri-state specialisation; shared-stack depth (incl. all-local) x register cap x box family.
mkdir -p gpurun_out
O=gpurun_out/r2_ab3.txt
: > $O
python -m pytest tests/test_gpu_parity.py -x -q -k "trace or primary or config3 or fuzz or area_light or empty_tile or intersect_with" > gpurun_out/r2_tests3.log 2>&1
echo "gpu tests (subset, default lib): $(tail -1 gpurun_out/r2_tests3.log)" >> $O
line() {
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
for box in 2 3; do PBRTB200_BOX=$box line "lib=default(s24,lb9/10) box=$box"; done
for v in s8_lb12 s8_lb10 s0_lb12 s0_lb10 s4_lb12; do
  for box in 2 3; do
    PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so PBRTB200_BOX=$box line "lib=$v box=$box"
  done
done
cat $O
for box in 2 3; do
PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_s8_lb12.so PBRTB200_BOX=$box ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio \
    --clock-control none -k regex:k_trace -s 2 -c 2 --csv --log-file gpurun_out/r2_ncu3_box$box.csv python scripts/prof_frame.py 2 > gpurun_out/r2_ncu3_box$box.log 2>&1
done
