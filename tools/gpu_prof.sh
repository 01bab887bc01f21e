#!/bin/bash
# tools/gpu_prof.sh <kernel-regex> <name> <bench args...> : one `ncu --set full` capture of a kernel of a bench.py run (for gpurun);
# the report lands in gpurun_out/<name>.ncu-rep and is summarised offline with tools/ncu_summary.py
K=$1; N=$2; shift 2
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$K" -s 3 -c 1 -o gpurun_out/$N -f python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/$N.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/$N.ncu-rep
