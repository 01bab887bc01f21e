"""LTO_ICW_PROF=1 LTO_K3=wl python tools/iwl_prof.py : per-warp cycle split of the warp-local K3 (lto_indirect_wl.cu)."""
import os, sys
os.environ["LTO_ICW_PROF"] = "1"
os.environ.setdefault("LTO_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments", "lib", "liblto_k3x.so"))   # bash tools/experiments/build_variant.sh k3x
os.environ["LTO_K3"] = "wl"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from lowthrustopt_b200 import capi, synthetic as S
h = capi.Handle(0)
n = 131072
b = S.indirect_batch(n, ndim=12, seed=20180002)
p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
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
grid, NW = 148, int(os.environ.get("IWL_NWARP", "8"))
c = w[:grid * NW * 8].reshape(grid, NW, 8)
att = c[..., 4]
print("kernel ms %.3f  attempts/seg %.2f  -> %.1f M seg/s" % (kms, ns[:, 1].mean(), n / kms / 1e3))
print("per warp: alive %.0f cycles, %.0f attempts (x 8 slots), %.0f refills" % (c[..., 6].mean(), att.mean(), c[..., 5].mean()))
print("cycles per attempt: refill %.0f  state pass %.0f  6 column passes %.0f (%.0f each)  decision+output %.0f  | total %.0f" % (
    (c[..., 0] / att).mean(), (c[..., 1] / att).mean(), (c[..., 2] / att).mean(), (c[..., 2] / att).mean() / 6, (c[..., 3] / att).mean(),
    (c[..., 6] / att).mean()))
print("share: refill %.1f%%  state %.1f%%  columns %.1f%%  decision %.1f%%" % tuple(100 * (c[..., i].sum() / c[..., 6].sum()) for i in range(4)))
print("slot-attempts per SM per kcycle: %.2f   (slots busy: %.1f%% of the 8 per warp)" % (
    (att.sum(axis=1) * 8 / c[..., 6].max(axis=1)).mean() * 1e3 / 1e3, 100 * ns[:, 1].sum() / (att.sum() * 8)))
