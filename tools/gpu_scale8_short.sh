#!/bin/bash
# trimmed 8-GPU validation of the default configuration (copy-engine gather): one full bench line
mkdir -p gpurun_out
NG=${NG:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n${NG}_final.json 2> gpurun_out/bench_n${NG}_final.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^W1017\|^$" gpurun_out/bench_n${NG}_final.err | tail -4
python - gpurun_out/bench_n${NG}_final.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'parity', d['parity']['gathered_equals_recomputed'])
print(' stage', {k: round(v,3) for k,v in d['stage_ms'].items()})
s=d['strong_c1000']; print(' strong', round(s['value']), round(s['ms_per_step'],3), round(s['e2e']['value']), s['parity']['gathered_equals_recomputed'])
print(' pipeline', d['pipeline']['value'], d['pipeline']['ms_per_step'], d['pipeline']['decode_nms_ms_per_image'])
print(' e2e_det', d['e2e_detections']['value'], d['e2e_detections']['ms_per_step']); print(' host_link', d['host_link'])
print(' sustained', d['sustained']['value'], d['sustained']['clocks'])
PY
