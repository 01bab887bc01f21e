#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "indirect or closures" > gpurun_out/pytest_ind.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ind.log
tail -25 gpurun_out/pytest_ind.log
for w in indirect12; do
timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
tail -3 gpurun_out/q_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$w.json"))
print("$w", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "att %.2f"%d["roofline"]["attempted_steps_per_segment"], "e2e %.3e"%d["e2e"]["value"], d["clocks"])
PY
done
