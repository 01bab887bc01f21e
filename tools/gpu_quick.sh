#!/bin/bash
# quick GPU check: direct parity tests + direct bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "direct or smoke or closures" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -5 gpurun_out/pytest_quick.log
for w in direct7_fixed direct6_fixed; do
timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$w.json"))
print("$w", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["clocks"])
PY
done
