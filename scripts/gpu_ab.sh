#!/bin/bash
# A/B of library variants (built with __graft_entry__.build_variant) on one GPU: bench stage times per variant,
# the GPU suite and the traversal kernels' counters with the first variant.  usage: gpu_ab.sh <variant> [variant ...]
mkdir -p gpurun_out
O=gpurun_out/r2_ab.txt
: > $O
line() {
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
line "lib=default"
for v in "$@"; do
  PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$v.so line "lib=$v"
done
line "lib=default(again)"
V=$1
PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$V.so python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab_tests.log 2>&1
echo "gpu tests with $V: $(tail -1 gpurun_out/r2_ab_tests.log)" >> $O
cat $O
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
for lib in default $V; do
  if [ $lib = default ]; then unset PBRTB200_LIB; else export PBRTB200_LIB=$PWD/pbrt_rust_b200/libpbrtb200_$lib.so; fi
  ncu --metrics $M --clock-control none -k regex:k_trace -s 2 -c 2 --csv --log-file gpurun_out/r2_ab_ncu_$lib.csv python scripts/prof_frame.py 2 > gpurun_out/r2_ab_ncu_$lib.log 2>&1
  echo "== $lib"; grep -v "^==" gpurun_out/r2_ab_ncu_$lib.csv | cut -d, -f5,13- | tail -19
done
