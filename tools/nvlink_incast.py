#!/usr/bin/env python
"""NVLink incast: every rank but the solver rank pushes a buffer into the solver rank's HBM at the same time (copy engines,
lto_push_async over CUDA-IPC mapped memory) -- the ingest rate one GPU sustains from N-1 peers, i.e. the ceiling of
"results delivered to the solver rank" (DESIGN.md section 7).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/nvlink_incast.py"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from lowthrustopt_b200 import capi

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = capi.Handle(local)
MB = int(os.environ.get("INCAST_MB", 256)); reps = 8
nbytes = MB << 20
if rank == 0:
    dst = h.dev_alloc(nbytes * world)
    payload = [h.ipc_export(dst)]
else:
    payload = [None]
dist.broadcast_object_list(payload, src=0)
base = dst if rank == 0 else h.ipc_open(payload[0])
src = torch.full((nbytes // 8,), float(rank), dtype=torch.float64, device="cuda")
res = {}
for senders in sorted({1, max(1, (world - 1) // 2), world - 1}):
    active = 1 <= rank <= senders
    for it in range(2):                                           # 0 = warm-up
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if active:
            for _ in range(reps):
                h.push_async(base + rank * nbytes, src.data_ptr(), nbytes)
            h.sync_copies()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0 if active else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res["%d_senders" % senders] = {"ingest_gbs": senders * reps * nbytes / dt.item() / 1e9, "per_sender_gbs": reps * nbytes / dt.item() / 1e9}
if rank == 0:
    print(json.dumps({"what": "copy-engine pushes into rank 0's HBM over NVLink, %d MiB x %d per sender, wall clock max over senders" % (MB, reps),
                      "world": world, **res}), flush=True)
dist.barrier()
if rank != 0:
    h.ipc_close(base)
dist.barrier()
if rank == 0:
    h.dev_free(dst)
h.close()
dist.destroy_process_group()
