#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/corr2.log; : > $LOG
for geo in "20 27 3 2" "45 37 5 1" "80 80 7 1"; do
  timeout 120 python tools/gpu_stage_check.py corr $geo >> $LOG 2>&1; echo "exit(corr 2cta $geo)=$?" >> $LOG
done
timeout 180 python tools/gpu_stage_check.py head 20 27 3 2 2>&1 | grep -v "head forward" >> $LOG; echo "exit(head)=$?" >> $LOG
tail -40 $LOG
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for V in 1cta 2cta; do
  if [ "$V" = "1cta" ]; then export OS2D_B200_CORR_1CTA=1; else unset OS2D_B200_CORR_1CTA; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$V.json')); print('$V', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()}, 'corr frac', round(d['roofline_corr']['frac'],3))"
done
cp gpurun_out/bench_2cta.json gpurun_out/bench.json
