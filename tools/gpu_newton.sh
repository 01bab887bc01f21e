#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newton.py -q > gpurun_out/pytest_newton.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_newton.log
tail -40 gpurun_out/pytest_newton.log
