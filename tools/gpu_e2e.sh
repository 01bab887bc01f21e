#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_e2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e2e.log
tail -3 gpurun_out/pytest_e2e.log
for w in direct7_fixed indirect12; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
tail -2 gpurun_out/q_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$w.json"))
print("$w", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
done
