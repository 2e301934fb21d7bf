#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 120 python scratch/small_bench.py bf16 2>&1 | tail -n 1
timeout 120 python scratch/small_bench.py bf16x3 2>&1 | tail -n 1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/z_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -n 3 gpurun_out/z_pytest_gpu.log
