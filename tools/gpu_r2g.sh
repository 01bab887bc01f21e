#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
LTO_ICW_PROF=1 timeout 60 python tools/ihc_prof.py > $O/prof.log 2>&1; head -6 $O/prof.log
