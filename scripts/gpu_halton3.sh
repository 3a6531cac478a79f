#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/t_all6.log
python scripts/fuzz_parity.py 0 250 ext 2>&1 | grep -v " ok " | tail -8 | tee gpurun_out/fuzz_ext3.txt
python scripts/halton_probe.py 2>&1 | tail -40 | tee gpurun_out/halton_probe.json
