#!/bin/bash
# 8-GPU validation: multi-GPU tests (world 2 inside), bench at N=8 with the copy-engine gather (full line) and NCCL (short)
mkdir -p gpurun_out
NG=${NG:-8}
nvidia-smi topo -m > gpurun_out/topo_n$NG.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_dist.py -q --timeout 200 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for MODE in copy_engine nccl fused; do
  EXTRA=""; if [ "$MODE" != "copy_engine" ]; then EXTRA="--sustained-seconds 0 --no-pipeline"; fi
  timeout 600 $TR --master-port 2953$NG bench.py --gpus $NG --steps 20 --warmup 3 --gather $MODE $EXTRA > gpurun_out/bench_n${NG}_$MODE.json 2> gpurun_out/bench_n${NG}_$MODE.err
  echo "N=$NG $MODE rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^W1017\|^$" gpurun_out/bench_n${NG}_$MODE.err | tail -5
  python - gpurun_out/bench_n${NG}_$MODE.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
except Exception as e:
    print('NO JSON', e); sys.exit(0)
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'parity', d['parity'] and d['parity']['gathered_equals_recomputed'])
print(' stage', {k: round(v,3) for k,v in d['stage_ms'].items()})
s=d.get('strong_c1000') or {}
print(' strong', s.get('value'), s.get('ms_per_step'), (s.get('e2e') or {}).get('value'), (s.get('parity') or {}).get('gathered_equals_recomputed'), s.get('error'), s.get('stage_ms'))
print(' pipeline', d.get('pipeline'))
print(' sustained', (d.get('sustained') or {}).get('value'), (d.get('sustained') or {}).get('clocks'))
PY
done
