#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/u_ablate.log
for a in 0; do CFFM_DFACT_ABLATE=$a timeout 120 python scratch/dg_time.py bf16x3 2>&1 | tail -n 1 >> gpurun_out/u_ablate.log; done
CFFM_DFACT_ABLATE=0 python scratch/dg_time.py bf16 2>&1 | tail -n 1 >> gpurun_out/u_ablate.log
CFFM_DFACT_ABLATE=4 python scratch/dg_time.py bf16 2>&1 | tail -n 1 >> gpurun_out/u_ablate.log
cat gpurun_out/u_ablate.log
