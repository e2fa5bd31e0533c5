#!/bin/bash
# class-wave experiment (row g): step time and DRAM bytes per step with the five kernels driven wave by wave
mkdir -p gpurun_out
for WV in 0 8 10 12 16 24; do
  timeout 300 python bench.py --steps 20 --warmup 3 --strong-classes 0 --sustained-seconds 0 --no-pipeline --no-cpu-baseline --wave-classes $WV > gpurun_out/wave_$WV.json 2> gpurun_out/wave_$WV.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/wave_$WV.json') if l.startswith('{')][-1]); print('wave $WV: value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
for WV in 0 10 16; do
  timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/wave_dram_$WV.csv \
    python bench.py --steps 2 --warmup 1 --strong-classes 0 --sustained-seconds 0 --no-pipeline --no-cpu-baseline --wave-classes $WV > gpurun_out/wave_dram_$WV.log 2>&1
  echo "ncu wave $WV rc=$?"
done
