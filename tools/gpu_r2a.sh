#!/bin/bash
# round 2, visit A: the GPU suite (incl. the config-scale parity tests) + baseline bench lines of the indirect workloads
O=gpurun_out/r2a; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
cp gpurun_out/parity_scale.json $O/ 2>/dev/null
for w in indirect12 indirect14 direct6_fixed; do
  timeout 120 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"))
except Exception as e: print("$w failed", e)
PY
done
