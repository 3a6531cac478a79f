#!/bin/bash
mkdir -p gpurun_out
python scripts/fuzz_parity.py 0 400 ext 2>&1 | grep -v " ok " | tail -15 | tee gpurun_out/fuzz_ext.txt
python scripts/fuzz_parity.py 0 300 2>&1 | tail -2 | tee gpurun_out/fuzz_base.txt
python scripts/halton_probe.py 2>&1 | tail -40 | tee gpurun_out/halton_probe.json
