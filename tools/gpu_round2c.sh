#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_head.py tests/test_gpu_dist.py -q -m gpu --timeout 200 2>&1 | tail -5
for A in 0 1 0 1; do
  timeout 200 python bench.py --steps 20 --warmup 3 --strong-classes 0 --no-pipeline --no-cpu-baseline --async-resample $A > gpurun_out/ak_$A.json 2> gpurun_out/ak_$A.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/ak_$A.json') if l.startswith('{')][-1]); print('async $A: value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'sustained', round(d['sustained']['value']), d['sustained']['clocks']['sm_mhz'], {k: round(v,3) for k,v in d['stage_ms'].items()})" || tail -3 gpurun_out/ak_$A.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for A in 0 1; do
timeout 600 $TR --master-port 2954$A bench.py --gpus 2 --steps 20 --warmup 3 --strong-classes 0 --no-pipeline --async-resample $A > gpurun_out/ak2_$A.json 2> gpurun_out/ak2_$A.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/ak2_$A.json') if l.startswith('{')][-1]); print('N=2 async $A: value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'parity', d['parity']['gathered_equals_recomputed'], 'sustained', round(d['sustained']['value']), {k: round(v,3) for k,v in d['stage_ms'].items()})" || tail -5 gpurun_out/ak2_$A.err
done
