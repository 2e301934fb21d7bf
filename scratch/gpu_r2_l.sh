#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/l_pytest.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/l_pytest.log
for wl in frappe ml-tag book-crossing criteo; do
  timeout 300 python bench.py --workload $wl --precision bf16 --modes none --workloads none --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/l_bench_${wl}.json 2> gpurun_out/l_bench_${wl}.err; echo "$wl rc=$?"; python - <<PY
import json
s=open("gpurun_out/l_bench_${wl}.json").read(); d=json.loads(s[s.index('{"metric'):])
print(d["value"], d["ms_per_step"], d["gpu_launches"]/30)
for k,v in list(d["kernels"].items())[:12]: print("  ", k, v["launches_per_step"], v["avg_ms"], v["share"])
PY
done
