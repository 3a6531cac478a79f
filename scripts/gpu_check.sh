#!/bin/bash
# quick GPU gate: parity tests + bench line (no CPU leg)
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/frame %.3f  Mrays/s %.0f  e2e %.0f  launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches']))
print(d['stage_ms_per_frame']); print(d['clocks'])"
