#!/bin/bash
mkdir -p gpurun_out
for v in joint state; do
timeout 300 python bench.py --workload indirect12 --err-norm $v --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/v_$v.json"))
print("$v", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "att %.2f"%d["roofline"]["attempted_steps_per_segment"])
PY
done
