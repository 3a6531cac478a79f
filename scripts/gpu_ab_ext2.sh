#!/bin/bash
mkdir -p gpurun_out
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_frame']
print('$1 ms/frame %.3f | trace %.3f shade %.3f shadow %.3f' % (d['ms_per_step'], s['ms_trace'], s['ms_shade'], s['ms_shadow']))"; }
{
PBRTB200_FORCE_EXT=1 python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/ab.err | tail -1 | line "c3 ext"
PBRTB200_FORCE_EXT=1 python scripts/run_configs.py c4 --frames 2 --no-oracle 2>>gpurun_out/ab.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 ext ms/frame %.2f shade %.2f' % (d['ms_per_frame'], d['device_stage_ms']['ms_shade']))"
python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/ab.err | tail -1 | line "c3 default"
} | tee gpurun_out/ab_ext2.txt
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/t_all3.log
