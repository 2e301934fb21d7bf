#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_fullsize.py tests/test_gpu_checkpoint.py -q --timeout 600 > gpurun_out/q_pytest.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/q_pytest.log
timeout 200 python scratch/fwd_time.py 8192 bf16x3 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --workloads none --modes none --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
s=open("gpurun_out/q_bench.json").read(); d=json.loads(s[s.index('{"metric'):])
print("bf16x3", d["value"], d["ms_per_step"])
for k,v in list(d["kernels"].items())[:8]: print("   ",k,v)
PY
