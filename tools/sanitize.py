"""Small invocations of every kernel added in this round, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from lowthrustopt_b200 import capi, synthetic as S
h = capi.Handle(0)
p1 = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
p2 = capi.indirect_params(p=2.0, thrustLimit=10.0)
for nd in (12, 14):
    b = S.indirect_batch(300, ndim=nd, seed=1)
    r = h.indirect(b["x0"], b["t0"], b["t1"], params=p1)
    r0 = h.indirect(b["x0"], b["t0"], b["t1"], params=p1, jac=False)
    assert np.all(r["status"] == 0) and np.all(r0["status"] == 0)
d = S.direct_batch(200, nstate=7, seed=2)
rd = h.direct(d["Xa"], d["Xb"], d["ua"], d["ub"], d["ta"], d["tb"])
rda = h.direct(d["Xa"], d["Xb"], d["ua"], d["ub"], d["ta"], d["tb"], params=capi.direct_params(mode=capi.LTO_ADAPTIVE))
c = S.continuation_batch(n_traj=6, n_seg_per_traj=29, ndim=12)
rt = h.indirect_traj(c["XC_all"], c["t_TU"], params=p2)
for adj in (False, True):
    u, st = h.indirect_newton(rt["phi"].reshape(6, 29, 12, 12), rt["defect"].reshape(6, 29, 12), adj)
    assert np.all(st == 0)
XC = c["XC_all"].copy(); XC[:, :, 6:] *= 0.1
s = h.indirect_solve_batch(XC, c["t_TU"], params=p2, max_iter=6)
from lowthrustopt_b200 import solvers as SV
fx = SV.demo_fixtures()
gpu = SV.GpuBackend(handle=h)
XCg, tg, tau1, tau2, s0, sf = SV.trajectory_stack_guess(fx[1], fx[3], backend=gpu)
for n in (6, 7):
    Xs = np.stack([XCg[:6]] * 3)
    if n == 7:
        Xs = np.concatenate([Xs, 1000.0 * np.ones((3, 1, 30))], axis=1)
    Xb, Ub, db, itb = SV.multiShoot_CRTBP_direct_batch(Xs, np.zeros((3, 3, 30)), tau1, tau2, np.stack([tg] * 3), capi.MU, capi.DU, capi.TU, 30, 10, 1e3, 2000.0,
                                                       *fx, backend=gpu)
    assert np.abs(db).max() <= 1e-6, (n, np.abs(db).max())
    st0, stf = SV.interpEndStates(tau1, tau2, *fx, capi.MU)
    rr = h.direct_solve_batch(Xs.transpose(0, 2, 1), np.zeros((3, 30, 3)), np.stack([tg] * 3), np.stack([st0] * 3), np.stack([stf] * 3), max_iter=12)
    assert np.all(rr["er"] <= 1e-6)
print("sanitize run ok: solve flags", s["status_flag"], "iters", s["iters"], "launches", h.launches)
h.close()
