#!/bin/bash
mkdir -p gpurun_out
NG=${NG:-2}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
LIST="1 $NG"; if [ "$NG" = "8" ]; then LIST="1 2 4 8"; fi; if [ "$NG" = "4" ]; then LIST="1 2 4"; fi
for N in $LIST; do
  if [ $N = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; fi
  echo "N=$N rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/bench_n$N.err | tail -3
  python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'])"
done
