#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/stag.log; : > $LOG
for st in conv1 conv2; do for geo in "20 27 3 2" "45 37 5 1" "33 24 2 1"; do
  timeout 120 python tools/gpu_stage_check.py $st $geo 2>&1 | grep -E "OK|FAIL|stage|Error|error" >> $LOG; echo "exit($st $geo)=$?" >> $LOG
done; done
cat $LOG | grep -v "^== stage.*H" | tail -30
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
