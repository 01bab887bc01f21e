import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from lowthrustopt_b200 import capi, synthetic
h = capi.Handle(0)
for (scale, pexp, tl, mi) in ((0.01, 2.0, 10.0, 8), (0.1, 2.0, 10.0, 8), (0.3, 2.0, 10.0, 10), (0.02, 1.0, None, 8)):
    c = synthetic.continuation_batch(n_traj=1024, n_seg_per_traj=200, ndim=12)
    XC = c["XC_all"].copy(); XC[:, :, 6:] *= scale / 0.1
    p = capi.indirect_params(p=pexp, thrustLimit=tl or 0.05, rho=1.0); p.max_attempts = 2000
    thr = None if tl else c["thrustLimit"]
    for rep in range(2):
        l0 = h.launches; t0 = time.perf_counter()
        r = h.indirect_solve_batch(XC, c["t_TU"], params=p, thrustLimit=thr, max_iter=mi)
        dt = time.perf_counter() - t0
    it = r["iters"]; fl = r["status_flag"]
    print("scale", scale, "p", pexp, "tl", tl, "wall %.1f ms" % (dt * 1e3), "dev %.1f ms" % h.last_kernel_ms, "launches", h.launches - l0,
          "iters hist", np.bincount(it), "flags", np.bincount(fl, minlength=3), "er max ok %.1e" % r["er"][fl == 0].max() if (fl == 0).any() else "")
# the Newton kernel alone
import torch
dev = torch.device("cuda", 0)
c = synthetic.continuation_batch(n_traj=1024, n_seg_per_traj=200, ndim=12)
r = h.indirect_traj(c["XC_all"], c["t_TU"], params=capi.indirect_params(p=2.0, thrustLimit=10.0))
phi = torch.from_numpy(r["phi"]).to(dev); d = torch.from_numpy(r["defect"]).to(dev); upd = torch.empty((1024, 201, 12), dtype=torch.float64, device=dev)
st = torch.cuda.ExternalStream(h.stream, device=dev)
for mode in (0, 1):
    for i in range(3):
        h.indirect_newton_dev(1024, 201, mode, phi.data_ptr(), d.data_ptr(), upd.data_ptr())
    h.sync()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        a.record()
        for i in range(10):
            h.indirect_newton_dev(1024, 201, mode, phi.data_ptr(), d.data_ptr(), upd.data_ptr())
        b.record()
    h.sync()
    print("newton kernel mode", mode, "%.3f ms for 1024 x 201 nodes" % (a.elapsed_time(b) / 10))
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        a.record()
        for i in range(10):
            h.indirect_newton_resolve_dev(1024, 201, mode, d.data_ptr(), upd.data_ptr())
        b.record()
    h.sync()
    print("newton resolve mode", mode, "%.3f ms" % (a.elapsed_time(b) / 10))
