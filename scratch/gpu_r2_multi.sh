#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_tests.log
NCCL_DEBUG=WARN timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -rs -s >> gpurun_out/multi_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/multi_tests.log
tail -n 25 gpurun_out/multi_tests.log
