#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python scratch/prof_step.py 8192 bf16x3 > gpurun_out/t_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dgrad0_fact -s 1 -c 1 -o gpurun_out/t_x3_dgrad0 python scratch/prof_step.py 8192 bf16x3 > gpurun_out/t_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/t_*.ncu-rep; tail -n 3 gpurun_out/t_plain.log
