#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_default.json").read().strip().splitlines()[-1])
    print("default", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "pageable", {k:v["value"] for k,v in d["e2e"]["pageable"].items() if isinstance(v,dict)})
except Exception as e: print("default failed", e); print(open("$O/bench_default.err").read()[-800:])
PY
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solvers.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
