#!/bin/bash
# ncu launch list of the bench command + one --set full capture of the dominant kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits > gpurun_out/smi_query.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown,clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap --format=csv,noheader,nounits >> gpurun_out/smi_query.txt 2>&1
cat gpurun_out/smi_query.txt
ncu --version | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 21 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"; tail -25 gpurun_out/launches.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_kernel|corr_kernel|resample_kernel' -s 15 -c 5 -o gpurun_out/prof_r01 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/
