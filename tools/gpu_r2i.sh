#!/bin/bash
# round 2, visit I: whole GPU suite + default bench line (pageable legs) + smoke
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
cp gpurun_out/parity_scale.json $O/ 2>/dev/null
timeout 200 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_default.json").read().strip().splitlines()[-1])
    print("default", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "pageable", {k:v["value"] for k,v in d["e2e"]["pageable"].items() if isinstance(v,dict)}, "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("variational",{}).get("value"))
except Exception as e: print("default failed", e)
PY
timeout 100 python bench.py --workload indirect12 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_indirect12.json 2> $O/bench_indirect12.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_indirect12.json").read().strip().splitlines()[-1])
    print("indirect12", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "pageable", {k:v["value"] for k,v in d["e2e"]["pageable"].items() if isinstance(v,dict)})
except Exception as e: print("indirect12 failed", e)
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
