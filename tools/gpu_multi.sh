#!/bin/bash
# tools/gpu_multi.sh <N> <tag> [workloads...] : multi-GPU evidence on one N-GPU box (gpurun --gpus N)
N=$1; TAG=$2; shift 2
O=gpurun_out/$TAG; mkdir -p $O
run() {  # run <n> <name> <bench args...>
  local n=$1 name=$2; shift 2
  if [ $n -eq 1 ]; then timeout 200 python bench.py --gpus 1 "$@" > $O/$name.json 2> $O/$name.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/$name.json").read().strip().splitlines() if l.startswith("{")][-1])
    x={k:d.get(k) for k in ("value","ms_per_step","n_gpus")}
    if "no_delivery" in d: x["no_delivery"]=d["no_delivery"]["value"]; x["ingest_gbs"]=d["delivery"]["solver_rank_ingest_gbs"]; x["nccl"]=d["allgather_nccl"]["value"]
    if "e2e" in d and d["e2e"]: x["e2e"]=d["e2e"].get("value")
    if "speedup_over_one_device" in d: x["speedup"]=d["speedup_over_one_device"]; x["identical"]=d["identical_results"]
    print("$name", x)
except Exception as e: print("$name failed", e); print(open("$O/$name.err").read()[-600:])
PY
}
for w in "$@"; do
  case $w in
    test) timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" > $O/pytest_multi.log 2>&1; tail -2 $O/pytest_multi.log ;;
    default) run $N default_n$N --steps 10 --warmup 3 --no-cpu-baseline ;;
    single) timeout 300 python bench.py --gpus $N --single-process --steps 10 --warmup 3 > $O/single_n$N.json 2> $O/single_n$N.err; run_dummy=1
            python -c "
import json; d=json.loads(open('$O/single_n$N.json').read().strip().splitlines()[-1]); print('single_n$N', d['value'], 'one device', d['same_batch_on_one_device']['value'], 'speedup', d['speedup_over_one_device'], 'identical', d['identical_results'])" ;;
    single_indirect) timeout 300 python bench.py --gpus $N --single-process --workload indirect12 --steps 5 --warmup 3 > $O/single_indirect12_n$N.json 2> $O/single_indirect12_n$N.err
            python -c "
import json; d=json.loads(open('$O/single_indirect12_n$N.json').read().strip().splitlines()[-1]); print('single_indirect12_n$N', d['value'], 'one device', d['same_batch_on_one_device']['value'], 'speedup', d['speedup_over_one_device'], 'identical', d['identical_results'])" ;;
    incast) timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/nvlink_incast.py > $O/incast_n$N.json 2> $O/incast_n$N.err; tail -1 $O/incast_n$N.json ;;
    *) run $N ${w}_n$N --workload $w --steps 5 --warmup 3 --no-cpu-baseline ;;
  esac
done
