#!/bin/bash
# completion counters in the host-buffer indirect pipeline (LTO_HOST_PROGRESS=1): correctness + e2e against the default
O=gpurun_out/progress
mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_chunk" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for w in indirect12 indirect14; do
  LTO_HOST_PROGRESS=1 timeout 40 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_prog.json 2> $O/bench_${w}_prog.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/progress/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("%-32s value %.4e  ms %.4f  e2e %.4e  launches %s" % (f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")))
PY
