#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "indirect or closures or solvers or sumsq" 2>&1 | tail -4
for v in cw q3; do
LTO_K3=$v timeout 300 python bench.py --workload indirect12 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b2_$v.json 2>gpurun_out/b2.err; tail -3 gpurun_out/b2.err
python - <<PY
import json
d=json.load(open("gpurun_out/b2_$v.json")); print("$v value %.3e ms %.4f frac %.3f e2e %.3e att %.2f"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["roofline"]["attempted_steps_per_segment"]))
PY
done
if [ "$1" = "prof" ]; then bash tools/gpu_prof.sh k_indirect_$2 prof_$2 --workload indirect12; fi
