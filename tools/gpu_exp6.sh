#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/gpu_class_pipeline_bench.py 100 > gpurun_out/class_pipeline.json 2> gpurun_out/class_pipeline.err; tail -3 gpurun_out/class_pipeline.err; cat gpurun_out/class_pipeline.json
