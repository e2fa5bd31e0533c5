#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -5
for N in 1 2; do
  if [ $N = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; fi
  echo "N=$N rc=$?"; tail -2 gpurun_out/bench_n$N.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['value']), round(d['ms_per_step'],3), d['e2e']['value'], d['clocks'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-300
