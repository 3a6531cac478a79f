#!/bin/bash
# Round-end GPU pass: newest GPU test, then the profile bundle (bench lines, ncu launch list,
# ncu --set full of the hot kernels) that scripts/make_profiles.py turns into profiles/.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "upload_validates or bump or extended" 2>&1 | tail -4 | tee gpurun_out/t_new2.log
bash scripts/gpu_profile.sh 2>&1 | tail -30
