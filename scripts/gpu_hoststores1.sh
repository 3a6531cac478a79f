#!/bin/bash
# One GPU: e2e of config 3 with the film staged + copied (0) and stored by k_film into the pinned host film (2), alternating.
nvidia-smi topo -m 2>/dev/null | head -6
for i in 1 2 3; do
  for v in 0 2; do
    PBRTB200_HOST_FILM_STORES=$v python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('host_film_stores=$v value %.0f (%.3f ms) e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
  done
done
