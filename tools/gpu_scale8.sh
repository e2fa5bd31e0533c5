#!/bin/bash
mkdir -p gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "N=$N rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/bench_n$N.err | tail -3
  python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'])"
done
