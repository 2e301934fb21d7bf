#!/bin/bash
# round-2 artefacts on one GPU: smoke, whole GPU suite, both bench arms, ncu launch list of the bench command,
# ncu --set full of the split-mode layer-0 kernels
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/z_gpu.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/z_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/z_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/z_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/z_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/z_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --modes none --workloads none --no-cpu-baseline"
$CMD > gpurun_out/z_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/z_launches_bench.csv $CMD > gpurun_out/z_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scratch/prof_step.py 8192 bf16x3 > gpurun_out/z_plain_step.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fwd0_fact|k_deinterleave_x1|k_wgrad0_fact|k_dgrad0_fact" -s 4 -c 4 -o gpurun_out/z_layer0_split python scratch/prof_step.py 8192 bf16x3 > gpurun_out/z_ncu_layer0.log 2>&1
echo "ncu layer0 rc=$?"; ls -la gpurun_out/z_*
