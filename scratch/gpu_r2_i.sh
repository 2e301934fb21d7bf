#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_bf16x3.py tests/test_gpu_fullsize.py -q --timeout 600 > gpurun_out/i_pytest.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --workloads none --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
s=open("gpurun_out/i_bench.json").read(); d=json.loads(s[s.index('{"metric'):])
print("bf16x3", d["value"], d["ms_per_step"])
for k,v in list(d["kernels"].items())[:8]: print("   ",k,v)
m=d["modes"]["bf16"]; print("bf16", m["value"], m["ms_per_step"], m["roofline"]["kernel"], m["roofline"]["frac"])
PY
timeout 300 python bench.py --precision bf16 --modes none --steps 10 --warmup 3 --workloads none --no-cpu-baseline > gpurun_out/i_bench_bf16.json 2> gpurun_out/i_bench_bf16.err
python - <<'PY'
import json
s=open("gpurun_out/i_bench_bf16.json").read(); d=json.loads(s[s.index('{"metric'):])
print("bf16", d["value"], d["ms_per_step"])
for k,v in list(d["kernels"].items())[:10]: print("   ",k,v)
PY
