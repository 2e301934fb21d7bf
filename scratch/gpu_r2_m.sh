#!/bin/bash
cd /root/repo
for rep in 1 2; do for v in 1 0; do echo -n "side3=$v "; CFFM_SIDE3=$v timeout 120 python scratch/small_bench.py bf16 2>&1 | tail -n 1; done; done
CFFM_SIDE3=1 timeout 200 python -m pytest tests/test_gpu_branches.py tests/test_gpu_parity.py -q -x 2>&1 | tail -n 2
