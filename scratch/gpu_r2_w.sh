#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/z_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -n 3 gpurun_out/z_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/z_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/z_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/z_bench_ref.json
