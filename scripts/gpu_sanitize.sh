#!/bin/bash
# compute-sanitizer over the render path: memcheck + racecheck (+ initcheck), summaries to gpurun_out/.
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_frames.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1) | $(grep -c 'sanitize_frames: done' gpurun_out/r2_sanitizer_$tool.log) run(s) completed"
done
