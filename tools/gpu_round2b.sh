#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_concurrent.sh
timeout 400 python -m pytest tests/test_gpu_dist.py -q --timeout 200 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 3 --sustained-seconds 0 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err
echo "N=2 rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^W1017\|^$" gpurun_out/bench_n2b.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2b.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'parity', d['parity'])
print(' stage', {k: round(v,3) for k,v in d['stage_ms'].items()})
print(' e2e_det', d['e2e_detections']); print(' host_link', d['host_link']); print(' pipeline', d['pipeline'])
s=d['strong_c1000']; print(' strong', round(s['value']), s['ms_per_step'], s['parity'])
PY
