#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python scratch/fwd_time.py 8192 bf16x3 > gpurun_out/r_fwd.log 2>&1
timeout 900 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_fullsize.py tests/test_gpu_trajectory.py -x -q > gpurun_out/r_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r_tests.log
python bench.py --steps 10 --warmup 3 --workloads none --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -3 gpurun_out/r_fwd.log gpurun_out/r_tests.log; cut -c1-600 gpurun_out/r_bench.json
