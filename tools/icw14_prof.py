"""LTO_ICW_PROF=1 python tools/icw14_prof.py : per-warp cycle split of the 14-dim indirect throughput kernel."""
import os, sys
os.environ["LTO_ICW_PROF"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from lowthrustopt_b200 import capi, synthetic as S
h = capi.Handle(0)
n = 131072
b = S.indirect_batch(n, ndim=14, seed=20180002)
p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
dev = torch.device("cuda", 0)
dx0 = torch.from_numpy(b["x0"]).to(dev); dt0 = torch.from_numpy(b["t0"]).to(dev); dt1 = torch.from_numpy(b["t1"]).to(dev)
d_def = torch.empty((n, 14), dtype=torch.float64, device=dev); d_ns = torch.empty((n, 2), dtype=torch.int32, device=dev)
d_phi = torch.empty((n, 14, 14), dtype=torch.float64, device=dev)
st = torch.cuda.ExternalStream(h.stream, device=dev)
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        h.indirect_dev(p, n, 0, 14, dx0.data_ptr(), dt0.data_ptr(), dt1.data_ptr(), None, None, None, d_def.data_ptr(), None, d_ns.data_ptr(), d_phi.data_ptr())
        e1.record()
    h.sync()
w = h.debug_profile().astype(np.float64)
grid, NW = 148, 8
c = w[:grid * NW * 4].reshape(grid, NW, 4)
pre = w[grid * NW * 4: grid * NW * 4 + grid]
st_, co = c[:, 0], c[:, 1:]
print("kernel ms %.3f  attempts/seg %.2f" % (e0.elapsed_time(e1), d_ns.cpu().numpy()[:, 1].mean()))
print("state : alive %.0f  work/attempt %.0f  wait/attempt %.0f  pre/attempt %.0f  tile-attempts %.0f" % (
    st_[:, 3].mean(), (st_[:, 0] / st_[:, 2]).mean(), (st_[:, 1] / st_[:, 2]).mean(), (pre / st_[:, 2]).mean(), st_[:, 2].mean()))
print("column: alive %.0f  busy/half-phase %.0f  wait/visit %.0f  half-phases %.0f  busy %.1f%%" % (
    co[..., 3].mean(), (co[..., 0] / co[..., 2]).mean(), (co[..., 1] / (co[..., 2] / 2)).mean(), co[..., 2].mean(), 100 * (co[..., 0] / co[..., 3]).mean()))
