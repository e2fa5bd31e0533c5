#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for T in 28 20 22 24 default; do
  if [ "$T" = "default" ]; then unset OS2D_B200_CONV_TILE_ROWS; else export OS2D_B200_CONV_TILE_ROWS=$T; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_T$T.json 2> gpurun_out/bench_T$T.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_T$T.json')); print('T=$T', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
cp gpurun_out/bench_Tdefault.json gpurun_out/bench.json
