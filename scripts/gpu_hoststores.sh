#!/bin/bash
# A/B of PBRTB200_HOST_FILM_STORES (k_film storing a band's rows straight into the page-locked host film) on N GPUs
N=${1:-8}
mkdir -p gpurun_out
for v in 0 1; do
  PBRTB200_HOST_FILM_STORES=$v python scripts/group_probe.py $N c3 40 > gpurun_out/hoststores_n${N}_$v.log 2>&1
  grep SUMMARY gpurun_out/hoststores_n${N}_$v.log
done
for v in 0 2; do
  PBRTB200_HOST_FILM_STORES=$v python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=1 host_film_stores=$v value %.0f e2e %.0f (%.3f ms)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
