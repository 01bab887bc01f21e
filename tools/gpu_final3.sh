#!/bin/bash
# last call of the round: the whole GPU suite on the final build (completion-counter variants apart), then the opt-in
# completion-counter pipeline (batched counting): its tests and its e2e
O=gpurun_out/final3
mkdir -p $O
timeout 75 python -m pytest tests -m gpu -q --deselect "tests/test_gpu_parity.py::test_indirect_multi_chunk_host_pipeline[12-40000-3]" --deselect "tests/test_gpu_parity.py::test_indirect_multi_chunk_host_pipeline[14-40000-3]" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 30 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_chunk" > $O/pytest_progress.log 2>&1; echo "pytest rc=$?" >> $O/pytest_progress.log
tail -3 $O/pytest_progress.log
for w in indirect12 indirect14; do
  LTO_HOST_PROGRESS=1 timeout 30 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_prog.json 2> $O/bench_${w}_prog.err
  python -c "
import json,sys
d=json.loads(open('$O/bench_${w}_prog.json').read().strip().splitlines()[-1]); print('$w prog: value %.4e e2e %.4e' % (d['value'], d['e2e']['value']))" 2>&1 | tail -1
done
