#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimizers.py tests/test_gpu_bf16.py tests/test_gpu_fullsize.py tests/test_gpu_checkpoint.py tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/h_pytest.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/h_pytest.log
timeout 300 python bench.py --workload frappe --precision bf16 --modes none --workloads none --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_frappe.json 2> gpurun_out/h_bench_frappe.err; echo "frappe rc=$?"
timeout 600 python bench.py --precision bf16 --modes none --workloads none --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_criteo.json 2> gpurun_out/h_bench_criteo.err; echo "criteo rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 --precision bf16 --modes none > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ("h_bench_frappe","h_bench_criteo","h_bench_n2"):
    s=open("gpurun_out/%s.json"%f).read(); d=json.loads(s[s.index('{"metric'):])
    print(f, d["value"], d["ms_per_step"], d["gpu_launches"], d.get("other_table_layout") and {k:v for k,v in d["other_table_layout"].items() if k!="kernels"})
    for k,v in d["kernels"].items():
        if k.startswith(("sparse","sort","shard","dp_","dense_adagrad","inner")): print("   ",k,v)
PY
