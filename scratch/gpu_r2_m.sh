#!/bin/bash
cd /root/repo
for rep in 1 2; do for mk in 0 1 3 7 11 15; do CFFM_SIDE_MASK=$mk timeout 120 python scratch/small_bench.py bf16 2>&1 | tail -n 1; done; done
