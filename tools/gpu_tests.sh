#!/bin/bash
# all GPU tests (no -x: show every failure), per-test timeout
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
