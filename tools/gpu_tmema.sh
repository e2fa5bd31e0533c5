#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('SMEM-A', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
export OS2D_B200_CONV_TMEM_A=1
for st in conv1 conv2; do for geo in "20 27 3 2" "28 24 2 1" "45 37 5 1"; do
  timeout 120 python tools/gpu_stage_check.py $st $geo 2>&1 | grep -E "OK|FAIL|rror" | cut -c1-100; echo "exit($st $geo)=$?"
done; done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmema.json 2> gpurun_out/bench_tmema.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tmema.json')); print('TMEM-A', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
