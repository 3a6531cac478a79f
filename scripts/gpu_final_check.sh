#!/bin/bash
# Last GPU gate of a round: the newest GPU tests first, then the whole GPU suite, then a short
# bench line.  Everything is logged under gpurun_out/ as it goes (the call may be cut short).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "extended or bump" > gpurun_out/t_new.log 2>&1
echo "rc=$?" >> gpurun_out/t_new.log
tail -5 gpurun_out/t_new.log
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -c 1500 gpurun_out/bench_quick.json
timeout 600 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/t_all.log 2>&1
echo "rc=$?" >> gpurun_out/t_all.log
tail -25 gpurun_out/t_all.log
