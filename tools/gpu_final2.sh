#!/bin/bash
# end-of-round validation on one B200: the whole GPU test-suite, smoke(), one bench line per workload, launch list of the default bench
O=gpurun_out/final2
mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 120 python bench.py > $O/bench_direct7_fixed.json 2> $O/bench_direct7_fixed.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in direct6_fixed direct7_adaptive indirect12 indirect14 indirect12_1m continuation continuation_solve; do
  timeout 120 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final2/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    r = d.get("roofline") or {}
    print("%-28s value %.4e  ms %.4f  frac %s  e2e %.4e  launches %s  cpu %s" % (d["config"]["workload"] + ("/ref" if d.get("impl") else ""), d["value"], d.get("ms_per_step", 0),
          ("%.3f" % r["frac"]) if r.get("frac") else "-", d["e2e"]["value"], d.get("gpu_launches"), ("%.3e" % d["cpu_baseline"]["value"]) if d.get("cpu_baseline") else "-"))
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01c_launches_direct7_fixed.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r01c_launches_direct7_fixed.log 2>&1
tail -2 $O/r01c_launches_direct7_fixed.log | cut -c1-300
