#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_bf16.py tests/test_gpu_fullsize.py tests/test_gpu_trajectory.py -x -q > gpurun_out/s_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/s_tests.log
python bench.py --steps 10 --warmup 3 --workloads none --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
tail -n 5 gpurun_out/s_tests.log; cut -c1-300 gpurun_out/s_bench.json
