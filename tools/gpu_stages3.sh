#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/stages3.log
: > $LOG
for st in conv2 conv3 head; do
  timeout 120 python tools/gpu_stage_check.py $st 20 27 3 2 >> $LOG 2>&1
  echo "exit($st)=$?" >> $LOG
done
for st in conv2 conv3 head; do
  timeout 180 python tools/gpu_stage_check.py $st 45 37 5 1 >> $LOG 2>&1
  echo "exit($st big)=$?" >> $LOG
done
grep -v "^  head forward" $LOG | tail -70
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['clocks'], d['e2e'])"
tail -3 gpurun_out/bench.err
