#!/bin/bash
# Round-1 evidence run: bench lines of every workload, launch lists, full ncu captures of the four throughput kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_direct7_fixed.json 2> gpurun_out/bench_direct7_fixed.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in direct6_fixed direct7_adaptive indirect12 indirect14 indirect12_1m continuation; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_direct7.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_indirect12.csv python bench.py --workload indirect12 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_continuation.csv python bench.py --workload continuation --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1
bash tools/gpu_prof.sh k_direct_cw prof_direct_cw
bash tools/gpu_prof.sh k_indirect_cw prof_indirect_cw --workload indirect12
bash tools/gpu_prof.sh k_indirect_state prof_indirect_state --workload continuation
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_direct_state -s 3 -c 1 -o gpurun_out/prof_direct_state -f python tools/k4_time.py > gpurun_out/prof_direct_state.log 2>&1
python - <<PY
import json
for w in ["direct7_fixed","direct6_fixed","direct7_adaptive","indirect12","indirect14","indirect12_1m","continuation","reference"]:
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%w) if l.startswith("{")][-1])
        r=d.get("roofline",{})
        print(w, "value %.4e ms %.4f frac %s e2e %.4e cpu %s"%(d["value"], d["ms_per_step"], r.get("frac"), d["e2e"]["value"], d.get("cpu_baseline",{}).get("value")))
    except Exception as e:
        print(w, "FAILED", e)
PY
