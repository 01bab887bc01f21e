"""LTO_ICW_PROF=1 python tools/ihc_prof.py [state] : per-warp cycle split of the half-column K3 (lto_indirect_hc.cu)."""
import os, sys
os.environ["LTO_ICW_PROF"] = "1"; os.environ.setdefault("LTO_K3", "hc")
os.environ.setdefault("LTO_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments", "lib", "liblto_k3x.so"))   # bash tools/experiments/build_variant.sh k3x
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from lowthrustopt_b200 import capi, synthetic as S
norm = capi.LTO_NORM_STATE if (len(sys.argv) > 1 and sys.argv[1] == "state") else capi.LTO_NORM_STATE_SENS
h = capi.Handle(0)
n = 131072
b = S.indirect_batch(n, ndim=12, seed=20180002)
p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05, err_norm=norm)
dev = torch.device("cuda", 0)
dx0 = torch.from_numpy(b["x0"]).to(dev); dt0 = torch.from_numpy(b["t0"]).to(dev); dt1 = torch.from_numpy(b["t1"]).to(dev)
d_def = torch.empty((n, 12), dtype=torch.float64, device=dev); d_ns = torch.empty((n, 2), dtype=torch.int32, device=dev)
d_phi = torch.empty((n, 12, 12), dtype=torch.float64, device=dev)
st = torch.cuda.ExternalStream(h.stream, device=dev)
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        h.indirect_dev(p, n, 0, 12, dx0.data_ptr(), dt0.data_ptr(), dt1.data_ptr(), None, None, None, d_def.data_ptr(), None, d_ns.data_ptr(), d_phi.data_ptr())
        e1.record()
    h.sync()
ns = d_ns.cpu().numpy()
kms = e0.elapsed_time(e1)
w = h.debug_profile().astype(np.float64)
grid, NW, NCW, NT = 148, 12, int(os.environ.get("IHC_NCW", "8")), 3
c = w[:grid * NW * 4].reshape(grid, NW, 4)
pre = w[grid * NW * 4: grid * NW * 4 + grid * NT].reshape(grid, NT)
co, sw = c[:, :NCW], c[:, NCW:NCW + NT]
print("kernel ms %.3f  attempts/seg %.2f  -> %.1f M seg/s" % (kms, ns[:, 1].mean(), n / kms / 1e3))
print("state : alive %.0f  work/attempt %.0f  wait/attempt %.0f  pre/attempt %.0f  attempts %.0f" % (
    sw[..., 3].mean(), (sw[..., 0] / sw[..., 2]).mean(), (sw[..., 1] / sw[..., 2]).mean(), (pre / sw[..., 2]).mean(), sw[..., 2].mean()))
print("column: alive %.0f  work/task %.0f  wait/tile-visit %.0f  tasks %.0f  busy %.1f%%" % (
    co[..., 3].mean(), (co[..., 0] / co[..., 2]).mean(), (co[..., 1] / (co[..., 2] / 3)).mean(), co[..., 2].mean(), 100 * (co[..., 0] / co[..., 3]).mean()))
att = w[grid * NW * 4 + grid * NT: grid * NW * 4 + grid * NT + grid * NCW].reshape(grid, NCW)
print("column: cycles inside col_attempt per task %.0f (the rest of a task: column load from the L2 scratch, header, candidate store, error hand-over)" % (att / co[..., 2]).mean())
for wi in range(NCW):
    print("  col warp %d: work/task %.0f busy %.1f%%" % (wi, (co[:, wi, 0] / co[:, wi, 2]).mean(), 100 * (co[:, wi, 0] / co[:, wi, 3]).mean()))
