#!/bin/bash
# round-2 profiles: ncu launch list of the bench command + one --set full capture of every kernel on the path
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --strong-classes 0 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${TAG}_prof -f \
    python tools/gpu_profile_driver.py > gpurun_out/${TAG}_prof.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${TAG}_prof.log
