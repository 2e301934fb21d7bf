#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --workloads none --no-cpu-baseline > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "rc=$?"; cut -c1-300 gpurun_out/n2_bench.json; tail -n 3 gpurun_out/n2_bench.err
