"""LTO_ICW_PROF=1 python tools/icw_prof.py : per-warp cycle split of the indirect throughput kernel."""
import os, sys
os.environ["LTO_ICW_PROF"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from lowthrustopt_b200 import capi, synthetic as S
norm = capi.LTO_NORM_STATE if (len(sys.argv) > 1 and sys.argv[1] == "state") else capi.LTO_NORM_STATE_SENS
h = capi.Handle(0)
b = S.indirect_batch(131072, ndim=12, seed=20180002)
p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05, err_norm=norm)
for _ in range(2):
    r = h.indirect(b["x0"], b["t0"], b["t1"], params=p)
w = h.debug_profile().astype(np.float64)
grid, NW, NT = 148, 8, 2
c = w[:grid * NW * 4].reshape(grid, NW, 4)
pre = w[grid * NW * 4: grid * NW * 4 + grid * NT].reshape(grid, NT)
st, co = c[:, :NT], c[:, NT:]
print("kernel ms %.3f  attempts/seg %.2f" % (h.last_kernel_ms, r["nsteps"][:, 1].mean()))
print("state : alive %.0f  work/attempt %.0f  wait/attempt %.0f  pre/attempt %.0f  attempts %.0f" % (
    st[..., 3].mean(), (st[..., 0] / st[..., 2]).mean(), (st[..., 1] / st[..., 2]).mean(), (pre / st[..., 2]).mean(), st[..., 2].mean()))
print("column: alive %.0f  work/half-phase %.0f  wait/visit %.0f  half-phases %.0f  busy %.1f%%" % (
    co[..., 3].mean(), (co[..., 0] / co[..., 2]).mean(), (co[..., 1] / (co[..., 2] / 2)).mean(), co[..., 2].mean(), 100 * (co[..., 0] / co[..., 3]).mean()))
for wi in range(6):
    print("  col warp %d: work/half-phase %.0f busy %.1f%%" % (wi, (co[:, wi, 0] / co[:, wi, 2]).mean(), 100 * (co[:, wi, 0] / co[:, wi, 3]).mean()))
