#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/u_ablate.log
timeout 90 python scratch/dg_time.py bf16x3 2>&1 | tail -n 2 >> gpurun_out/u_ablate.log
echo "rc $?" >> gpurun_out/u_ablate.log
cat gpurun_out/u_ablate.log
if grep -q conv_dgrad_l0 gpurun_out/u_ablate.log; then
  timeout 400 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_fullsize.py -x -q > gpurun_out/s_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/s_tests.log
  tail -n 4 gpurun_out/s_tests.log
  timeout 300 python bench.py --steps 10 --warmup 3 --workloads none --modes none --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
  cut -c1-200 gpurun_out/s_bench.json
fi
