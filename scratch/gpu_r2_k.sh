#!/bin/bash
# N GPUs: the driver's scaling command in the default (bf16x3, sharded) configuration and in bf16
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/k_bench_n${N}.json 2> gpurun_out/k_bench_n${N}.err; echo "bench n$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 10 --warmup 3 --precision bf16 --modes none > gpurun_out/k_bench_bf16_n${N}.json 2> gpurun_out/k_bench_bf16_n${N}.err; echo "bench bf16 n$N rc=$?"
python - <<PY
import json
for f in ("k_bench_n$N","k_bench_bf16_n$N"):
    s=open("gpurun_out/%s.json"%f).read(); d=json.loads(s[s.index('{"metric'):])
    print(f, d["dtype"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "strong", d["strong"] and (d["strong"]["value"], d["strong"]["ms_per_step"]), "other", d["other_table_layout"] and (d["other_table_layout"]["tables"], d["other_table_layout"]["value"], d["other_table_layout"]["ms_per_step"]))
    print("   modes", {k:(v["value"],v["ms_per_step"]) for k,v in d["modes"].items()})
PY
tail -3 gpurun_out/k_bench_n${N}.err
