#!/bin/bash
# two kernel streams in the host-buffer indirect pipeline: correctness (multi-chunk test) and e2e with 1 vs 2 streams
mkdir -p gpurun_out/streams
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_chunk or multi_device or full_size" > gpurun_out/streams/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/streams/pytest.log
tail -6 gpurun_out/streams/pytest.log
for s in 1 2; do
  for w in indirect12 indirect14 indirect12_1m; do
    LTO_HOST_STREAMS=$s timeout 60 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/streams/bench_${w}_s$s.json 2> gpurun_out/streams/bench_${w}_s$s.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/streams/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("%-40s value %.4e  ms %.4f  e2e %.4e  launches %s" % (f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")))
PY
