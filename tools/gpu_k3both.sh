#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "indirect" > gpurun_out/pytest_k3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k3.log
tail -4 gpurun_out/pytest_k3.log
for w in indirect12 indirect14; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
tail -2 gpurun_out/q_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$w.json"))
print("$w", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["attempted_steps_per_segment"])
PY
done
