#!/bin/bash
# Run on the B200 box (gpurun): bench lines (both arms), ncu launch list, ncu --set full of the hot kernels
# (config 3), and of the traversal kernels on config 5 (7.6 GB of BVH: does it stay issue / L1-bound?).
# usage: gpu_profile.sh a   bench (both arms) + launch list + full sets of k_trace, k_shade
#        gpu_profile.sh b   full sets of k_raygen, k_film and of k_trace on config 5
# (two calls: gpurun brings back at most 64 MiB of gpurun_out/ per call)
mkdir -p gpurun_out
PART=${1:-a}
if [ "$PART" = "a" ]; then
  python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -c 3000 gpurun_out/bench.json
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_ref.json
  # launch list (every launch, serialised, cold caches: compare shares only)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/prof_frame.py 2 > gpurun_out/launches.log 2>&1
  # full sets: launches from the second frame (config 3 = one chunk per frame: 2 trace launches per frame)
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 2 -f -o gpurun_out/prof_k_trace \
      python scripts/prof_frame.py 3 > gpurun_out/prof_k_trace.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 1 -f -o gpurun_out/prof_k_shade \
      python scripts/prof_frame.py 3 > gpurun_out/prof_k_shade.log 2>&1
else
  for k in k_raygen k_film; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
        python scripts/prof_frame.py 3 > gpurun_out/prof_$k.log 2>&1
  done
  ncu --set full --clock-control none -k regex:k_trace -s 2 -c 2 -f -o gpurun_out/prof_c5_k_trace \
      python scripts/prof_frame.py 1 c5 > gpurun_out/prof_c5_k_trace.log 2>&1
fi
ls -la gpurun_out | grep ncu-rep
