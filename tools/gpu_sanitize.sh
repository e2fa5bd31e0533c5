#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck of the whole head at a small configuration (test plan
# item 5) and of the round-2 kernels (fused decode + NMS incl. the chunked path, gather, channels-last pack, K1 next to conv1)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tests/tools/gpu_stage_check.py head 20 27 3 1 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|== stage|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/tools/gpu_sanitize_r02.py > gpurun_out/sanitizer_r02_$tool.log 2>&1
  echo "r02 $tool rc=$?"; grep -E "ERROR SUMMARY|== stage|Error|error|decode|concurrent" gpurun_out/sanitizer_r02_$tool.log | head -10
done
