#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_gpu_optimizers.py tests/test_gpu_checkpoint.py -q --timeout 300 > gpurun_out/e_pytest.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/e_pytest.log
for wl in frappe ml-tag book-crossing; do
  timeout 300 python bench.py --workload $wl --precision bf16 --modes none --workloads none --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_${wl}.json 2> gpurun_out/e_bench_${wl}.err; echo "$wl rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/e_bench_${wl}.json"))
print(d["value"], d["ms_per_step"], d["gpu_launches"])
for k,v in list(d["kernels"].items())[:14]: print("  ", k, v["launches_per_step"], v["avg_ms"], v["share"])
PY
done
