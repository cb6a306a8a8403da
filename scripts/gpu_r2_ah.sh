#!/bin/bash
# round 2, call AH: racecheck / synccheck / initcheck on the small-mesh driver
set -u
mkdir -p gpurun_out
for tool in racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_ah_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_ah_$tool.log | sort | uniq -c | sort -rn | head -8
done
