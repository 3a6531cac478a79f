#!/bin/bash
# A/B of the streaming (evict-first) hints on the per-sample buffers: default build vs libpbrtb200_nostream.so
# (__graft_entry__.build_variant('nostream', ['PB_STREAM_HINTS=0'])) on configs 3, 4, 5; GPU suite with the default.
mkdir -p gpurun_out
O=gpurun_out/r2_stream_ab.txt
: > $O
line() {  # $1 label, $2 config, $3 steps
  python bench.py --config $2 --steps $3 --warmup 3 --no-cpu 2>gpurun_out/r2_last.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$2 $1 ms/frame %.3f e2e %.3f | raygen %.3f trace %.3f shade %.3f shadow %.3f film %.3f | sm %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film'], d['clocks']['sm_mhz']))" >> $O
}
NS=$PWD/pbrt_rust_b200/libpbrtb200_nostream.so
line default c3 20
PBRTB200_LIB=$NS line nostream c3 20
line default c3 20
PBRTB200_LIB=$NS line nostream c3 20
line default c4 3
PBRTB200_LIB=$NS line nostream c4 3
line default c5 3
PBRTB200_LIB=$NS line nostream c5 3
python -m pytest tests -m gpu -q > gpurun_out/r2_tests.log 2>&1
echo "gpu tests (default build): $(tail -1 gpurun_out/r2_tests.log)" >> $O
cat $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__issue_active.avg.pct_of_peak_sustained_elapsed
for lib in default nostream; do
  if [ $lib = default ]; then unset PBRTB200_LIB; else export PBRTB200_LIB=$NS; fi
  ncu --metrics $M --clock-control none -s 7 -c 7 --csv --log-file gpurun_out/r2_stream_ncu_$lib.csv python scripts/prof_frame.py 2 > gpurun_out/r2_stream_ncu_$lib.log 2>&1
  echo "== $lib"; grep -v "^==" gpurun_out/r2_stream_ncu_$lib.csv | cut -d, -f5,13- | sed 's/"Command line profiler metrics",//' | tail -42
done
