#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halton" 2>&1 | tail -40 | tee gpurun_out/t_halton.log
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/t_all4.log
