#!/bin/bash
# Multi-GPU visit: usage gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
for w in direct7_fixed indirect12 indirect12_1m continuation; do
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_$N.json 2> gpurun_out/scale_${w}_$N.err
tail -2 gpurun_out/scale_${w}_$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${w}_$N.json") if l.startswith("{")][-1])
    print("$w N=$N value %.3e ms %.3f"%(d["value"], d["ms_per_step"]), "e2e %.3e"%d["e2e"]["value"], d.get("allgather",{}))
except Exception as e:
    print("$w FAILED", e)
PY
done
