#!/bin/bash
# compute-sanitizer synccheck / initcheck / racecheck over the same tiny coverage driver (1 GPU).
mkdir -p gpurun_out
for tool in synccheck initcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool --print-limit 5 python benchmark/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Uninitialized|Barrier error|Divergent" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
