#!/bin/bash
# other BASELINE configs as bench workloads (not the headline line): C=1000 on one GPU, batch 8 x 960 px x 200 classes
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --classes 1000 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_c1000.json 2> gpurun_out/bench_c1000.err; echo "rc=$?"; tail -2 gpurun_out/bench_c1000.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c1000.json')); print('C=1000', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['pipeline'])"
timeout 600 python bench.py --steps 5 --warmup 3 --classes 200 --size 960 --batch 8 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "rc=$?"; tail -2 gpurun_out/bench_cfg5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5.json')); print('cfg5 B8 960px C200', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
timeout 600 python bench.py --steps 20 --warmup 3 --classes 4 --size 512 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1.json')); print('cfg1 512px C4', round(d['value']), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})"
