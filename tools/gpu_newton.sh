#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newton.py -q > gpurun_out/pytest_newton.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_newton.log
tail -30 gpurun_out/pytest_newton.log
timeout 600 python tools/try_batch.py 2>&1 | tail -8
timeout 400 python bench.py --workload continuation_solve --steps 5 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_continuation_solve.json 2> gpurun_out/bench_cs.err; tail -2 gpurun_out/bench_cs.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_continuation_solve.json"))
print("continuation_solve dev ms %.2f  wall ms %.2f  traj/s %.0f  iters %d conv %d launches %d value %.3e e2e %.3e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["trajectories_per_s"], d["iterations_max"], d["trajectories_converged"], d["gpu_launches"], d["value"], d["e2e"]["value"]))
PY
