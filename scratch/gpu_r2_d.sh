#!/bin/bash
# round 2, call D (1 GPU): whole GPU suite after the K=16/64 scoring change, scoring sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/d_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -15 gpurun_out/d_pytest_gpu.log
timeout 900 python bench_score.py --batches 64,1024,16384,65536 --reps 5 > gpurun_out/d_score_sweep.jsonl 2> gpurun_out/d_score_sweep.err; echo "sweep rc=$?"; tail -5 gpurun_out/d_score_sweep.jsonl; tail -3 gpurun_out/d_score_sweep.err
