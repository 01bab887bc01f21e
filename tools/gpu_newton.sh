#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newton.py -q > gpurun_out/pytest_newton.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_newton.log
tail -30 gpurun_out/pytest_newton.log
timeout 600 python tools/try_batch.py 2>&1 | tail -8
