#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for st in conv2 head; do timeout 180 python tools/gpu_stage_check.py $st 45 37 5 1 2>&1 | grep -v "head forward" | tail -14; done
for V in nosplit default; do
  if [ "$V" = "nosplit" ]; then export OS2D_B200_CONV_NO_TAIL_SPLIT=1; else unset OS2D_B200_CONV_NO_TAIL_SPLIT; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$V.json')); print('$V', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
cp gpurun_out/bench_default.json gpurun_out/bench.json
