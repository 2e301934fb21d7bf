#!/bin/bash
mkdir -p gpurun_out
python scratch/prof_step.py 8192 bf16x3 > gpurun_out/j_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fwd0_fact -s 1 -c 1 -o gpurun_out/j_x3_fwd0 python scratch/prof_step.py 8192 bf16x3 > gpurun_out/j_ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'^k_tc$' -s 20 -c 2 -o gpurun_out/j_x3_bwd0 python scratch/prof_step.py 8192 bf16x3 > gpurun_out/j_ncu2.log 2>&1
echo "ncu2 rc=$?"; ls -la gpurun_out/j_*.ncu-rep
