#!/bin/bash
# K3 with slot-major stash + bulk STM stores: smoke, K3 parity, bench, ncu capture
O=gpurun_out/r2n; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 0; }
tail -1 $O/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_scale.py tests/test_gpu_newton.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 300 python bench.py --workload indirect12 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_indirect12.json 2> $O/bench_indirect12.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_indirect12.json").read().splitlines() if l.startswith("{")][-1])
print("indirect12", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
PY
timeout 300 python bench.py --workload indirect14 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_indirect14.json 2> $O/bench_indirect14.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_indirect14.json").read().splitlines() if l.startswith("{")][-1])
print("indirect14", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
PY
bash tools/gpu_prof.sh "^k_indirect_cw$" r02_k_indirect_cw --workload indirect12
bash tools/gpu_prof.sh "^k_indirect_cw14$" r02_k_indirect_cw14 --workload indirect14
