#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --classes 4 --size 512 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1.json')); print('cfg1 512px C4', round(d['value']), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('cfg2', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
