#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch lists + full captures of the two throughput kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_direct7_fixed.json 2> gpurun_out/bench_direct7_fixed.err
timeout 600 python bench.py --workload direct6_fixed --no-cpu-baseline > gpurun_out/bench_direct6_fixed.json 2> gpurun_out/bench_direct6_fixed.err
timeout 600 python bench.py --workload indirect12 --steps 5 --warmup 3 > gpurun_out/bench_indirect12.json 2> gpurun_out/bench_indirect12.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_direct7.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_indirect12.csv python bench.py --workload indirect12 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_ind.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_direct_cw -s 3 -c 1 -o gpurun_out/prof_direct_cw -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_indirect_cw -s 3 -c 1 -o gpurun_out/prof_indirect_cw -f python bench.py --workload indirect12 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ind.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_direct7_fixed.json; cat gpurun_out/bench_indirect12.json
