#!/bin/bash
# 8-GPU box: group parity tests, bench.py at N = 1, 2, 4, 8 (torchrun as the driver launches it), then configs 5 and 4 at N = 1..8.
bash scripts/gpu_multi.sh 8
python scripts/scale_configs.py c5 c4 --frames 3 > gpurun_out/r2_scale_configs.log 2>&1
tail -12 gpurun_out/r2_scale_configs.log | cut -c1-420
