#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
for v in 0 136 144 152; do
  LTO_HC_REGS=$v timeout 60 python bench.py --workload indirect12 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_hc_$v.json 2> $O/bench_hc_$v.err || { echo "bench $v failed"; tail -5 $O/bench_hc_$v.err; exit 0; }
  python -c "
import json; d=json.loads(open('$O/bench_hc_$v.json').read().strip().splitlines()[-1]); print('hc regs $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
LTO_ICW_PROF=1 timeout 60 python tools/ihc_prof.py > $O/prof.log 2>&1; head -3 $O/prof.log
