#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/t_all5.log
python scripts/fuzz_parity.py 400 200 ext 2>&1 | tail -1 | tee gpurun_out/fuzz_ext2.txt
python scripts/halton_probe.py 2>&1 | tail -40 | tee gpurun_out/halton_probe.json
