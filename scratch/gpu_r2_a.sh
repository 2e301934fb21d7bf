#!/bin/bash
# round 2, call A: build check, smoke, the whole GPU suite (new modes in their own processes with hard timeouts), bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
free -g > gpurun_out/a_mem.txt; nproc >> gpurun_out/a_mem.txt
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_gpu_bf16x3.py -q -x --timeout 300 > gpurun_out/a_pytest_x3.log 2>&1; echo "x3 rc=$?"; tail -5 gpurun_out/a_pytest_x3.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_bf16x3.py --deselect tests/test_gpu_trajectory.py > gpurun_out/a_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -15 gpurun_out/a_pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_trajectory.py -q --timeout 600 > gpurun_out/a_pytest_traj.log 2>&1; echo "traj rc=$?"; tail -15 gpurun_out/a_pytest_traj.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/a_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/a_bench_ref.json
