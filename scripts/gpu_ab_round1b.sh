#!/bin/bash
# A/B on the B200 box: any-hit child ordering (PBRTB200_SHADOW_MODE 0..3) on config 3, and the
# general texture evaluator (PBRTB200_FORCE_EXT) on config 3 (constant textures) and config 4
# (checkerboard / uv textures over 4K x 64 spp).
mkdir -p gpurun_out
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f e2e %.3f | raygen %.2f trace %.3f shade %.3f shadow %.3f film %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], s['ms_raygen'], s['ms_trace'], s['ms_shade'], s['ms_shadow'], s['ms_film']))"; }
for rep in 1 2; do
for m in 0 2 3 1; do
  PBRTB200_SHADOW_MODE=$m python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/ab.err | tail -1 | line "shadow_mode=$m"
done
done | tee gpurun_out/ab_shadow_mode.txt
for e in 0 1; do
  PBRTB200_FORCE_EXT=$e python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/ab.err | tail -1 | line "c3 force_ext=$e"
done | tee gpurun_out/ab_force_ext.txt
for e in 0 1; do
  echo "c4 force_ext=$e"; PBRTB200_FORCE_EXT=$e python scripts/run_configs.py c4 --frames 3 --no-oracle 2>>gpurun_out/ab.err | tail -3
done | tee -a gpurun_out/ab_force_ext.txt
