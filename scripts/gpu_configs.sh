#!/bin/bash
# BASELINE.json configs 2, 4, 5 at full size on one GPU (parity on crops), results to gpurun_out/
mkdir -p gpurun_out
for c in c2 c4 c5; do
  python scripts/run_configs.py $c > gpurun_out/config_${c}_n1.log 2>&1
  tail -2 gpurun_out/config_${c}_n1.log
done
