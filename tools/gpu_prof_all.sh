#!/bin/bash
# round-2 profiles: one `ncu --set full` capture per throughput kernel (final build) + the launch list of the default bench command
mkdir -p gpurun_out
bash tools/gpu_prof.sh "k_direct_cw" r02_k_direct_cw --workload direct7_fixed
bash tools/gpu_prof.sh "k_direct_cw" r02_k_direct_cw_n6 --workload direct6_fixed
bash tools/gpu_prof.sh "k_indirect_cw<" r02_k_indirect_cw --workload indirect12
bash tools/gpu_prof.sh "k_indirect_cw14" r02_k_indirect_cw14 --workload indirect14
LTO_K3=hc bash tools/gpu_prof.sh "k_indirect_hc" r02_k_indirect_hc --workload indirect12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_direct7_fixed.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_indirect12.csv python bench.py --workload indirect12 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_i12.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_indirect14.csv python bench.py --workload indirect14 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_i14.log 2>&1
ls -la gpurun_out/r02_*
