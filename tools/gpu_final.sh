#!/bin/bash
# round-end style validation: GPU parity suite, smoke, default bench line, ncu launch list of the same command
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -2 gpurun_out/bench_final2.err; cut -c1-400 gpurun_out/bench_final2.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu3.log 2>&1; tail -3 gpurun_out/launches_final2.csv | cut -c1-200
