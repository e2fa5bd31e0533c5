#!/bin/bash
# 1-GPU round: all tests, PDL A/B, wave experiment, ncu profiles
mkdir -p gpurun_out
bash tools/gpu_tests.sh
for V in pdl nopdl; do
  if [ $V = nopdl ]; then export OS2D_B200_NO_PDL=1; else unset OS2D_B200_NO_PDL; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --strong-classes 0 --no-pipeline --no-cpu-baseline > gpurun_out/ab_$V.json 2> gpurun_out/ab_$V.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/ab_$V.json') if l.startswith('{')][-1]); print('$V: value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'sustained', round(d['sustained']['value']), {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
unset OS2D_B200_NO_PDL
bash tools/gpu_waves.sh
bash tools/gpu_profile_r02.sh
