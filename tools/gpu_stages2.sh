#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/stages2.log
: > $LOG
for st in pack corr head; do
  timeout 120 python tools/gpu_stage_check.py $st 20 27 3 2 >> $LOG 2>&1
  echo "exit($st)=$?" >> $LOG
done
for st in corr head; do
  timeout 180 python tools/gpu_stage_check.py $st 45 37 5 1 >> $LOG 2>&1
  echo "exit($st big)=$?" >> $LOG
done
tail -150 $LOG
