#!/bin/bash
# round-1 session-4 GPU check: new Newton/solver tests first, then the whole GPU suite, smoke and the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_newton.py -x -q > gpurun_out/pytest_newton.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_newton.log
tail -25 gpurun_out/pytest_newton.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_newton.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-600
