#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "indirect_vs_oracle" > gpurun_out/pytest_i14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_i14.log
tail -30 gpurun_out/pytest_i14.log
timeout 300 python bench.py --workload indirect14 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_indirect14.json 2> gpurun_out/q_indirect14.err
tail -3 gpurun_out/q_indirect14.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_indirect14.json"))
print("indirect14", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["attempted_steps_per_segment"])
PY
