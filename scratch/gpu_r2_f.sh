#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "sparse" > gpurun_out/f_pytest.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/f_pytest.log
python scratch/prof_small.py frappe bf16 3 > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_frappe.csv python scratch/prof_small.py frappe bf16 3 > gpurun_out/f_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/f_plain.log
