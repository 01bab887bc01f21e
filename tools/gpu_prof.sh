#!/bin/bash
# ncu capture of one kernel: usage gpu_prof.sh <kernel-regex> <out-name> <bench args...>
mkdir -p gpurun_out
K=$1; O=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/$O -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$O.log 2>&1
tail -2 gpurun_out/$O.log
