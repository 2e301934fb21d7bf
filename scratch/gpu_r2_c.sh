#!/bin/bash
# round 2, call C (2 GPUs): sharded equivalence, split-mode tests (direct + factorised forward), trajectory + its diagnostic, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 300 -k sharded > gpurun_out/c_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -12 gpurun_out/c_pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_bf16x3.py -q --timeout 300 -s > gpurun_out/c_pytest_x3.log 2>&1; echo "x3 rc=$?"; tail -8 gpurun_out/c_pytest_x3.log
timeout 300 python scratch/diag_traj.py > gpurun_out/c_diag_traj.log 2>&1; echo "diag rc=$?"
timeout 600 python -m pytest tests/test_gpu_trajectory.py -q --timeout 600 > gpurun_out/c_pytest_traj.log 2>&1; echo "traj rc=$?"; tail -8 gpurun_out/c_pytest_traj.log
timeout 600 python bench.py --steps 10 --warmup 3 --workloads none --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/c_bench.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 --precision bf16 --modes none > gpurun_out/c_bench_n2.json 2> gpurun_out/c_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/c_bench_n2.json; tail -3 gpurun_out/c_bench_n2.err
