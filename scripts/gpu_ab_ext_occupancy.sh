#!/bin/bash
# A/B on the B200 box: occupancy target of the EXT shading kernel (libs built with
# -DPB_SHADE_EXT_MIN_BLOCKS=4/6/8 as pbrt_rust_b200/lib_ext{4,6,8}.so), forced onto config 3 / 4.
mkdir -p gpurun_out
cp pbrt_rust_b200/libpbrtb200.so /tmp/lib_default.so
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f | trace %.3f shade %.3f shadow %.3f' % (d['ms_per_step'], s['ms_trace'], s['ms_shade'], s['ms_shadow']))"; }
for mb in 4 6 8; do
  cp pbrt_rust_b200/lib_ext$mb.so pbrt_rust_b200/libpbrtb200.so
  PBRTB200_FORCE_EXT=1 python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/ab.err | tail -1 | line "c3 ext min_blocks=$mb"
  PBRTB200_FORCE_EXT=1 python scripts/run_configs.py c4 --frames 2 --no-oracle 2>>gpurun_out/ab.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 ext min_blocks=$mb ms/frame %.2f shade %.2f' % (d['ms_per_frame'], d['device_stage_ms']['ms_shade']))"
done | tee gpurun_out/ab_ext_occupancy.txt
cp /tmp/lib_default.so pbrt_rust_b200/libpbrtb200.so
python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/ab.err | tail -1 | line "c3 default" | tee -a gpurun_out/ab_ext_occupancy.txt
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/t_all2.log
