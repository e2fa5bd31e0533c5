#!/bin/bash
# final 1-GPU round: every GPU test, smoke, the default bench line, the ncu launch list of the same command
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final_bench_n1.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print(' stage', {k: round(v,4) for k,v in d['stage_ms'].items()})
print(' roofline', d['roofline']['frac'], 'corr', d['roofline_corr']['frac'], 'whole', d['tensor_frac_whole_step'])
print(' sustained', round(d['sustained']['value']), d['sustained']['clocks'], d['sustained']['roofline']['frac'])
print(' pipeline', d['pipeline']); print(' e2e_det', d['e2e_detections']); print(' cpu', d['cpu_baseline'])
s=d['strong_c1000']; print(' strong', round(s['value']), s['ms_per_step'], round(s['e2e']['value']))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/final_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
