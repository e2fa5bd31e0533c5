#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 300 python bench.py --steps 20 --warmup 3 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/b1.json') if l.startswith('{')][-1]); print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'])); print(d['pipeline'])"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'label_nms|gather_det' -o gpurun_out/r02_prof_nms -f python tools/gpu_profile_driver.py > gpurun_out/r02_prof_nms.log 2>&1
echo "ncu rc=$?"
