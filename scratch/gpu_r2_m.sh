#!/bin/bash
cd /root/repo
timeout 120 python scratch/small_bench.py bf16 2>&1 | tail -n 1
timeout 120 python scratch/small_time.py frappe bf16 2>&1 | grep -i "colsum\|total"
timeout 300 python -m pytest tests/test_gpu_branches.py tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_gpu_bf16x3.py -q -x 2>&1 | tail -n 2
