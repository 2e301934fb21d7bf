#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for ss in 1; do
  echo "== CFFM_SIDE_STREAM=$ss"
  CFFM_SIDE_STREAM=$ss timeout 300 python bench.py --steps 10 --warmup 3 --modes bf16 --no-cpu-baseline > gpurun_out/zz_bench_$ss.json 2> gpurun_out/zz_bench_$ss.err
  python - <<PY
import json
s=open("gpurun_out/zz_bench_$ss.json").read(); d=json.loads(s[s.index('{"metric'):])
print("criteo bf16x3", d["value"], d["ms_per_step"], "bf16", d["modes"]["bf16"]["value"])
for w,v in d["workloads"].items(): print(w, {p:(x["value"], x["ms_per_step"]) for p,x in v.items() if isinstance(x,dict) and "ms_per_step" in x})
PY
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -n 3
