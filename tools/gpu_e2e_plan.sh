#!/bin/bash
# after the chunk-schedule change of the host-buffer pipeline: full GPU test-suite + e2e of the throughput workloads
mkdir -p gpurun_out/e2e_plan
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/e2e_plan/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e2e_plan/pytest_gpu.log
tail -4 gpurun_out/e2e_plan/pytest_gpu.log
for w in direct7_fixed direct6_fixed indirect12 indirect14 indirect12_1m; do
  timeout 60 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_plan/bench_$w.json 2> gpurun_out/e2e_plan/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/e2e_plan/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("%-16s value %.4e  ms %.4f  e2e %.4e  launches %s" % (d["config"]["workload"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")))
PY
