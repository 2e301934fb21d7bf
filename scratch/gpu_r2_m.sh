#!/bin/bash
cd /root/repo
for p in bf16 bf16x3; do
echo -n "default     "; timeout 120 python scratch/small_bench.py $p 2>&1 | tail -n 1
echo -n "factorised  "; CFFM_FACT_MIN_FIELDS=1 CFFM_FACT_MIN_BATCH=1 timeout 120 python scratch/small_bench.py $p 2>&1 | tail -n 1
done
