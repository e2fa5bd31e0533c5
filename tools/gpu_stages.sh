#!/bin/bash
# Runs every stage check under its own timeout and collects the logs in gpurun_out/stages.log
mkdir -p gpurun_out
LOG=gpurun_out/stages.log
: > $LOG
{
nvidia-smi -L
nproc
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
ls /root/reference 2>&1 | head -3
} >> $LOG 2>&1
for st in env pack corr conv1 conv2 conv3 resample head; do
  timeout 120 python tests/tools/gpu_stage_check.py $st ${STAGE_ARGS:-20 27 3 2} >> $LOG 2>&1
  echo "exit($st)=$?" >> $LOG
done
# a second geometry: multi-tile in y (H > 32) and odd widths
for st in corr conv1 conv2 conv3 head; do
  timeout 180 python tests/tools/gpu_stage_check.py $st 45 37 5 1 >> $LOG 2>&1
  echo "exit($st big)=$?" >> $LOG
done
tail -150 $LOG
