#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py tests/test_gpu_bf16x3.py -q --timeout 600 -s > gpurun_out/o_pytest.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/o_pytest.log; grep -E "^(bf16|bf16x3|fp32|gelu) " gpurun_out/o_pytest.log | head -12
