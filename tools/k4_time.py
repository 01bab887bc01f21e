"""Device time of the defect-only (K4) paths: python tools/k4_time.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from lowthrustopt_b200 import capi, synthetic as S
h = capi.Handle(0); dev = torch.device("cuda", 0)
st = torch.cuda.ExternalStream(h.stream, device=dev)
def timeit(f, n=10):
    for _ in range(3): f()
    h.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(n): f()
        e1.record()
    h.sync()
    return e0.elapsed_time(e1) / n
n = 65536
b = S.direct_batch(n, nstate=7); d = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
o_def = torch.empty((n, 7), dtype=torch.float64, device=dev); o_err = torch.empty(n, dtype=torch.float64, device=dev); o_st = torch.empty(n, dtype=torch.int32, device=dev)
o_jac = torch.empty((n, 20, 7), dtype=torch.float64, device=dev)
p = capi.direct_params()
for jac in (False, True):
    ms = timeit(lambda: h.direct_dev(p, n, 0, 7, 10, d["Xa"].data_ptr(), d["Xb"].data_ptr(), d["ua"].data_ptr(), d["ub"].data_ptr(), d["ta"].data_ptr(), d["tb"].data_ptr(),
                                     o_def.data_ptr(), o_err.data_ptr(), o_st.data_ptr(), o_jac.data_ptr() if jac else None))
    fl = 2 * 9 * 13 * 60 + 2 * 9 * (55 + 12 + 8) * 2 * 7          # state-only flops per segment (approx)
    print("direct7 n=%d jac=%s: %.3f ms  %.1f M seg/s  (state-only ~%.1f TFLOP/s)" % (n, jac, ms, n / ms / 1e3, fl * n / ms / 1e9))
for name, n, mk in (("indirect12 config4", 131072, None), ("continuation 1024x200", 204800, "c")):
    if mk is None:
        b = S.indirect_batch(n, ndim=12); x0 = torch.from_numpy(b["x0"]).to(dev); t0 = torch.from_numpy(b["t0"]).to(dev); t1 = torch.from_numpy(b["t1"]).to(dev); nn = 0
        tl = None
    else:
        c = S.continuation_batch(); x0 = torch.from_numpy(c["XC_all"]).to(dev); t0 = torch.from_numpy(c["t_TU"]).to(dev); t1 = None; nn = 201
        tl = torch.from_numpy(c["thrustLimit"]).to(dev)
    o_def = torch.empty((n, 12), dtype=torch.float64, device=dev); o_st = torch.empty(n, dtype=torch.int32, device=dev); o_ns = torch.empty((n, 2), dtype=torch.int32, device=dev)
    o_phi = torch.empty((n, 12, 12), dtype=torch.float64, device=dev)
    q = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
    for jac in (False, True):
        ms = timeit(lambda: h.indirect_dev(q, n, nn, 12, x0.data_ptr(), t0.data_ptr(), t1.data_ptr() if t1 is not None else None, None, tl.data_ptr() if tl is not None else None, None,
                                           o_def.data_ptr(), o_st.data_ptr(), o_ns.data_ptr(), o_phi.data_ptr() if jac else None), n=5)
        att = o_ns[:, 1].double().mean().item()
        print("%s jac=%s: %.3f ms  %.1f M seg/s  attempts %.2f  (state-only ~%.2f TFLOP/s)" % (name, jac, ms, n / ms / 1e3, att, att * 13 * 250 * n / ms / 1e9))
