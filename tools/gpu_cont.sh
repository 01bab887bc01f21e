#!/bin/bash
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
timeout 600 python -m pytest tests -m gpu -x -q -k "sumsq or multi_device" 2>&1 | tail -5
for w in continuation indirect12_1m; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -3 gpurun_out/bench_$w.err
done
else
for w in continuation indirect12_1m direct7_fixed; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_$N.json 2> gpurun_out/scale_${w}_$N.err; tail -3 gpurun_out/scale_${w}_$N.err
done
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_cont*.json")+glob.glob("gpurun_out/bench_indirect12_1m.json")+glob.glob("gpurun_out/scale_*_$N.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "N=%d value %.3e ms %.3f"%(d["n_gpus"], d["value"], d["ms_per_step"]), "e2e %.3e"%d["e2e"]["value"], "launches", d["gpu_launches"], d.get("allgather",{}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
