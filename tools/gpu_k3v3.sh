#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solvers.py tests/test_gpu_newton.py -q -x -k "indirect or solve or newton" > gpurun_out/pytest_k3v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k3v3.log
tail -15 gpurun_out/pytest_k3v3.log
for v in v2; do
LTO_K3=$v timeout 300 python bench.py --workload indirect12 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_indirect12_$v.json 2> gpurun_out/q_indirect12_$v.err
tail -2 gpurun_out/q_indirect12_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_indirect12_$v.json"))
print("indirect12 $v", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["attempted_steps_per_segment"])
PY
done
timeout 300 python tools/icw_prof.py 2>&1 | tail -6
