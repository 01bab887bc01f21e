#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
for v in iso; do
  IHC_NCW=9 LTO_K3=hc LTO_B200_LIB=$PWD/tools/experiments/lib/liblto_$v.so LTO_ICW_PROF=1 timeout 60 python tools/ihc_prof.py > $O/prof_$v.log 2>&1; echo "== $v rc=$?"; cat $O/prof_$v.log | head -16
done
