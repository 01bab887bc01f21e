#!/bin/bash
# round 2, closing run on one B200: smoke, whole GPU suite, bench lines, fresh ncu captures of K1 (nstate 7 / 6) and the launch list
O=gpurun_out/r2final; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $O/smoke.log; exit 0; }
tail -1 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
cp gpurun_out/parity_scale.json $O/ 2>/dev/null
line() {
  local name=$1; shift
  timeout 400 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_$name.json").read().splitlines() if l.startswith("{")][-1])
    r=d.get("roofline") or {}
    print("$name", "%.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "frac", r.get("frac"), "exec", (r.get("executed") or {}).get("frac"), "e2e %.4g" % d["e2e"]["value"],
          "pageable", {k: round(v["value"] / 1e6, 1) for k, v in (d["e2e"].get("pageable") or {}).items() if isinstance(v, dict)},
          "cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("variational") or {}).get("value"), "launches", d.get("gpu_launches"))
except Exception as e: print("$name failed", e); print(open("$O/bench_$name.err").read()[-500:])
PY
}
line direct7_fixed --steps 10 --warmup 3
line reference --impl reference --steps 3 --warmup 1
line direct6_fixed --workload direct6_fixed --steps 10 --warmup 3
line direct7_adaptive --workload direct7_adaptive --steps 10 --warmup 3 --no-cpu-baseline
bash tools/gpu_prof.sh "^k_direct_cw$" r02_k_direct_cw --workload direct7_fixed
bash tools/gpu_prof.sh "^k_direct_cw$" r02_k_direct_cw_n6 --workload direct6_fixed
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_direct7_fixed.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
