#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
LTO_TRACE=1 timeout 100 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct" -o faulthandler_timeout=40 -s > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; grep -v "^  File" $O/pytest.log | tail -30 | cut -c1-160
