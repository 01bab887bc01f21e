#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 60 python bench.py --workload indirect12 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_hc.json 2> $O/bench_hc.err || { echo "bench failed"; tail -5 $O/bench_hc.err; exit 0; }
python -c "
import json; d=json.loads(open('$O/bench_hc.json').read().strip().splitlines()[-1]); print('hc', d['value'], d['ms_per_step'], d['roofline']['frac'])"
bash tools/gpu_prof.sh k_indirect_hc r2c_k_indirect_hc --workload indirect12
