#!/bin/bash
# pytest -m gpu, bench, ncu launch list of the bench command + one --set full capture of the dominant kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()}); print(d['clocks'], d['cpu_baseline'], d['postproc'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 18 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_kernel|corr_kernel|resample_kernel|conv3s_kernel' -s 12 -c 6 -o gpurun_out/prof_r01 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
echo "ncu full rc=$?"
