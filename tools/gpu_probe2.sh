#!/bin/bash
# 2-GPU probe: symmetric memory / copy-engine push / NCCL timing, then the existing dist tests incl. the fused gather
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -30 >> gpurun_out/host.txt; numactl -H >> gpurun_out/host.txt 2>&1
ls /root/reference > gpurun_out/ref_present.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/gpu_probe_p2p.py > gpurun_out/probe_p2p.log 2>&1
echo "probe rc=$?"; grep '^{' gpurun_out/probe_p2p.log
OS2D_B200_TEST_FUSED_GATHER=1 timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -15 gpurun_out/pytest_dist.log
