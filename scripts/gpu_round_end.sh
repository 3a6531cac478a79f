#!/bin/bash
# Round-end GPU pass: the whole GPU suite, a slice of the extended fuzz, the Halton probe, then the
# profile bundle (bench lines, ncu launch list, ncu --set full of the hot kernels) that
# scripts/make_profiles.py turns into profiles/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/t_final.log
python scripts/fuzz_parity.py 600 120 ext 2>&1 | tail -1 | tee gpurun_out/fuzz_ext_final.txt
python scripts/halton_probe.py > gpurun_out/halton_probe.json 2>&1; tail -22 gpurun_out/halton_probe.json
bash scripts/gpu_profile.sh 2>&1 | tail -6
