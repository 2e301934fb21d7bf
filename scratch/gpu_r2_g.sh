#!/bin/bash
mkdir -p gpurun_out
python scratch/prof_small.py frappe bf16 3 > gpurun_out/g_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_seg_rows|k_seg_long|k_wgrad_reduce|k_head_grads|k_sum_chunks_layers|k_small_sort" -s 6 -c 6 -o gpurun_out/g_small python scratch/prof_small.py frappe bf16 3 > gpurun_out/g_ncu1.log 2>&1
echo "ncu1 rc=$?"
python scratch/prof_step.py 8192 bf16 > gpurun_out/g_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_seg_chunks|k_seg_fixup|k_inner_linear_bwd|k_inner_linear_fwd|k_inner_dense_grad" -s 5 -c 5 -o gpurun_out/g_criteo python scratch/prof_step.py 8192 bf16 > gpurun_out/g_ncu2.log 2>&1
echo "ncu2 rc=$?"; ls -la gpurun_out/g_*.ncu-rep
