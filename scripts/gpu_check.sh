#!/bin/bash
# GPU test-suite + one bench line (no CPU leg); log to gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_tests.log 2>&1
tail -25 gpurun_out/r2_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
tail -c 1500 gpurun_out/r2_bench_quick.json; tail -3 gpurun_out/r2_bench_quick.err
