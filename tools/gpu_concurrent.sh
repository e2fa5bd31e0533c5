#!/bin/bash
# stage-concurrent K1 || conv1: parity test (under timeout) then a sweep of the SM split
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "stage_concurrent" --timeout 120 2>&1 | tail -5
for CS in 0 14 16 18 20 22 24 28; do
  timeout 200 python bench.py --steps 20 --warmup 3 --strong-classes 0 --sustained-seconds 0 --no-pipeline --no-cpu-baseline --concurrent-corr $CS > gpurun_out/cc_$CS.json 2> gpurun_out/cc_$CS.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/cc_$CS.json') if l.startswith('{')][-1]); print('corr_sms $CS: value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), {k: round(v,3) for k,v in d['stage_ms'].items()})" || tail -3 gpurun_out/cc_$CS.err
done
