#!/bin/bash
# end-of-session validation on one B200: the whole GPU test-suite, smoke(), and one bench line per workload
mkdir -p gpurun_out/final
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log
tail -4 gpurun_out/final/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
timeout 300 python bench.py > gpurun_out/final/bench_direct7_fixed.json 2> gpurun_out/final/bench_direct7_fixed.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
for w in direct6_fixed direct7_adaptive indirect12 indirect14 indirect12_1m continuation continuation_solve; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 6 > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    r = d.get("roofline") or {}
    print("%-28s value %.4e  ms %.4f  frac %s  e2e %.4e  launches %s  cpu %s" % (d["config"]["workload"] + ("/ref" if d.get("impl") else ""), d["value"], d.get("ms_per_step", 0),
          ("%.3f" % r["frac"]) if r.get("frac") else "-", d["e2e"]["value"], d.get("gpu_launches"), ("%.3e" % d["cpu_baseline"]["value"]) if d.get("cpu_baseline") else "-"))
PY
