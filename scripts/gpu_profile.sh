#!/bin/bash
# Run on the B200 box (gpurun): bench line, ncu launch list, ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
# launch list: skip frame 1 (cold), list frame 2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/prof_frame.py 2 > gpurun_out/launches.log 2>&1
# full sets: 2 launches of each hot kernel from the second frame
for k in k_trace k_shade k_film k_raygen k_resolve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 2 -f -o gpurun_out/prof_$k \
      python scripts/prof_frame.py 2 > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
