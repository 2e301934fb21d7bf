#!/bin/bash
# round 2, call B (2 GPUs): data-parallel + row-sharded equivalence, re-run of the tests fixed after call A, bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 300 > gpurun_out/b_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -25 gpurun_out/b_pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_bf16x3.py -q --timeout 300 -s > gpurun_out/b_pytest_x3.log 2>&1; echo "x3 rc=$?"; tail -8 gpurun_out/b_pytest_x3.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checkpoint.py -q --timeout 300 > gpurun_out/b_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -8 gpurun_out/b_pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_trajectory.py -q --timeout 600 > gpurun_out/b_pytest_traj.log 2>&1; echo "traj rc=$?"; tail -8 gpurun_out/b_pytest_traj.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 --precision bf16 --modes none > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/b_bench_n2.json; tail -3 gpurun_out/b_bench_n2.err
