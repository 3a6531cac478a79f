#!/bin/bash
# Run on the B200 box (gpurun): bench line, ncu launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2500 gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
# launch list (every launch, serialised, cold caches: compare shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/prof_frame.py 2 > gpurun_out/launches.log 2>&1
# full sets: launches from the second frame
# (config 3 = one chunk per frame: 1 raygen, 1 shade, 2 trace launches; skip the first two frames)
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 2 -f -o gpurun_out/prof_k_trace \
    python scripts/prof_frame.py 3 > gpurun_out/prof_k_trace.log 2>&1
for k in k_shade k_raygen; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k \
      python scripts/prof_frame.py 3 > gpurun_out/prof_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_film -s 2 -c 1 -f -o gpurun_out/prof_k_film \
      python scripts/prof_frame.py 3 > gpurun_out/prof_k_film.log 2>&1
ls -la gpurun_out | head -30
