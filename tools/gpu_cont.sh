#!/bin/bash
# usage: gpu_cont.sh N [gather]
N=${1:-1}; G=${2:-peer}
mkdir -p gpurun_out
run() { if [ "$N" = "1" ]; then timeout 600 python bench.py "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; fi; }
for w in continuation indirect12_1m direct7_fixed indirect12; do
run --workload $w --gather $G --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_${N}_$G.json 2> gpurun_out/scale_${w}_${N}_$G.err; tail -4 gpurun_out/scale_${w}_${N}_$G.err | cut -c1-300
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*_${N}_$G.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        ag=d.get("allgather",{})
        print(f, "N=%d value %.3e ms %.3f"%(d["n_gpus"], d["value"], d["ms_per_step"]), "e2e %.3e"%d["e2e"]["value"], "launches", d["gpu_launches"], "| allgather", ag.get("value"), "peer", ag.get("peer_gather",{}).get("value"), ag.get("peer_gather",{}).get("matches_allgather"))
    except Exception as e:
        print(f, "FAILED", e)
PY
