#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 90 python scratch/dg_time.py bf16 2>&1 | tail -n 1
timeout 90 python scratch/dg_time.py bf16 2>&1 | tail -n 1
timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -n 2
