#!/bin/bash
# GPU visit: full parity suite + bench lines of every workload at N=1.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_direct7_fixed.json 2> gpurun_out/bench_direct7_fixed.err
for w in indirect12 indirect12_1m continuation; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
tail -3 gpurun_out/bench_$w.err
done
python - <<PY
import json
for w in ["direct7_fixed","indirect12","indirect12_1m","continuation"]:
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%w))
        print(w, "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["clocks"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(w, "FAILED", e)
PY
