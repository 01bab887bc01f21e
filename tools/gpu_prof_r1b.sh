#!/bin/bash
# ncu evidence for the kernels added/changed in this session (one GPU): full captures + launch lists
mkdir -p gpurun_out
bash tools/gpu_prof.sh k_indirect_cw r01b_k_indirect_cw --workload indirect12
bash tools/gpu_prof.sh k_indirect_cw14 r01b_k_indirect_cw14 --workload indirect14
bash tools/gpu_prof.sh k_indirect_newton r01b_k_indirect_newton --workload continuation_solve
for w in direct7_fixed indirect14 continuation_solve; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_launches_$w.log 2>&1
  tail -1 gpurun_out/r01b_launches_$w.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
