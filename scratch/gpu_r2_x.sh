#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python scratch/prof_small.py frappe bf16 4 > gpurun_out/x_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/x_launches_frappe.csv python scratch/prof_small.py frappe bf16 4 > gpurun_out/x_ncu.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/x_plain.log
