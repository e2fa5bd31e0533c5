#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck of the whole head at a small configuration (test plan item 5)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tests/tools/gpu_stage_check.py head 20 27 3 1 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|== stage|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
