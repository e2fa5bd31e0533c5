#!/bin/bash
# NG-GPU bench round: detector test, 1-GPU bench (full line), NG-GPU bench per gather mode, reference arms
mkdir -p gpurun_out
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
except Exception as e:
    print(sys.argv[1], 'NO JSON', e); sys.exit(0)
def g(o,*k):
    for x in k:
        if o is None: return None
        o=o.get(x)
    return o
print(sys.argv[1], 'value', round(d.get('value',0)), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(g(d,'e2e','value') or 0), 'e2e_ms', g(d,'e2e','ms_per_step'),
      'launches', d.get('gpu_launches'), 'parity', g(d,'parity','gathered_equals_recomputed'))
print('  stage', {k: round(v,3) for k,v in (d.get('stage_ms') or {}).items()}, 'clocks', d.get('clocks'))
s=d.get('strong_c1000') or {}
print('  strong', round(s.get('value',0)), 'ms', s.get('ms_per_step'), 'e2e', round(g(s,'e2e','value') or 0), 'parity', g(s,'parity','gathered_equals_recomputed'), s.get('error'), {k: round(v,3) for k,v in (s.get('stage_ms') or {}).items()})
print('  pipeline', d.get('pipeline'))
su=d.get('sustained') or {}
print('  sustained', round(su.get('value',0)), su.get('ms_per_step'), su.get('clocks'), g(su,'roofline','frac'))
print('  roofline', g(d,'roofline','frac'), 'corr', g(d,'roofline_corr','frac'), 'cpu', d.get('cpu_baseline'))
PY
}
if [ "$NG" != "1" ]; then
timeout 300 python -m pytest tests/test_gpu_dist.py -q -k "detector" --timeout 200 2>&1 | tail -5
fi
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "N=1 rc=$?"; tail -3 gpurun_out/bench_n1.err; show gpurun_out/bench_n1.json
if [ "$NG" != "1" ]; then
for MODE in copy_engine fused nccl; do
  EXTRA=""; if [ "$MODE" != "copy_engine" ]; then EXTRA="--sustained-seconds 0 --no-pipeline"; fi
  timeout 600 $TR --master-port 2952$NG bench.py --gpus $NG --steps 20 --warmup 3 --gather $MODE $EXTRA > gpurun_out/bench_n${NG}_$MODE.json 2> gpurun_out/bench_n${NG}_$MODE.err
  echo "N=$NG $MODE rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^W1017\|^$" gpurun_out/bench_n${NG}_$MODE.err | tail -5; show gpurun_out/bench_n${NG}_$MODE.json
done
fi
if [ -n "$WITH_REF" ]; then
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_cpu.json 2> gpurun_out/bench_ref_cpu.err; echo "ref rc=$?"; show gpurun_out/bench_ref_cpu.json
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/bench_refgpu_cfg2.json 2> gpurun_out/bench_refgpu_cfg2.err; echo "refgpu rc=$?"; tail -2 gpurun_out/bench_refgpu_cfg2.err; cat gpurun_out/bench_refgpu_cfg2.json | cut -c1-600
timeout 600 python bench.py --impl reference-gpu --steps 3 --warmup 3 --batch 8 --size 960 --classes 200 > gpurun_out/bench_refgpu_cfg5.json 2> gpurun_out/bench_refgpu_cfg5.err; echo "refgpu5 rc=$?"; tail -2 gpurun_out/bench_refgpu_cfg5.err; cat gpurun_out/bench_refgpu_cfg5.json | cut -c1-600
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --batch 8 --size 960 --classes 200 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"; show gpurun_out/bench_cfg5.json
fi
