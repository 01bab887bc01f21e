"""Timing of the direct QP kernel (lto_direct_qp_dev) and of the batched direct solve on a B200."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from lowthrustopt_b200 import capi, solvers as SV
h = capi.Handle(0)
gpu = SV.GpuBackend(handle=h)
fx = SV.demo_fixtures()
XCg, tg, tau1, tau2, s0, sf = SV.trajectory_stack_guess(fx[1], fx[3], backend=gpu)
T, N = 1024, 30
rng = np.random.default_rng(3)
for n in (6, 7):
    Xs = np.stack([XCg[:6] + 1e-4 * rng.standard_normal((6, N)) for _ in range(T)])
    if n == 7:
        Xs = np.concatenate([Xs, 1000.0 * np.ones((T, 1, N))], axis=1)
    Us = np.zeros((T, 3, N)); tb = np.stack([tg] * T)
    for rep in range(2):
        t0 = time.perf_counter(); l0 = h.launches
        Xb, Ub, db, itb = SV.multiShoot_CRTBP_direct_batch(Xs, Us, tau1, tau2, tb, capi.MU, capi.DU, capi.TU, N, 10, 1e3, 2000.0, *fx, backend=gpu)
        dt = time.perf_counter() - t0
    print("nstate %d: batched direct solve (Python driver) of %d trajectories x %d nodes: %.1f ms wall, iterations %s, max defect %.1e, launches %d"
          % (n, T, N, dt * 1e3, np.bincount(itb), np.abs(db).max(), h.launches - l0))
    st0, stf = SV.interpEndStates(tau1, tau2, *fx, capi.MU)
    Xr = Xs.transpose(0, 2, 1).copy(); Ur = Us.transpose(0, 2, 1).copy()
    for rep in range(3):
        t0 = time.perf_counter(); l0 = h.launches
        rr = h.direct_solve_batch(Xr, Ur, tb, np.stack([st0] * T), np.stack([stf] * T), mass=1e3, nsteps=10, max_iter=100)
        dt = time.perf_counter() - t0
    print("   lto_direct_solve_batch (resident loop): %.1f ms wall, %.2f ms on the device, iterations %s, max defect %.1e, launches %d, agreement with the Python driver %.1e"
          % (dt * 1e3, h.last_kernel_ms, np.bincount(rr["iters"]), np.abs(rr["defect"]).max(), h.launches - l0, np.abs(rr["X_all"].transpose(0, 2, 1) - Xb).max()))
    # the QP kernel alone
    r = h.direct_traj(Xs.transpose(0, 2, 1).copy(), Us.transpose(0, 2, 1).copy(), tb, nsteps=10, jac=True)
    dev = torch.device("cuda", 0)
    jac = torch.from_numpy(r["jac"]).to(dev); dfc = torch.from_numpy(r["defect"]).to(dev)
    U = torch.zeros((T, N, 3), dtype=torch.float64, device=dev); tt = torch.from_numpy(tb).to(dev)
    b0 = torch.zeros((T, 6 + (n == 7)), dtype=torch.float64, device=dev); bf = torch.zeros((T, 6), dtype=torch.float64, device=dev)
    xu = torch.empty((T, N, n), dtype=torch.float64, device=dev); uu = torch.empty((T, N, 3), dtype=torch.float64, device=dev)
    import ctypes as C
    L = capi.lib()
    def qp():
        rc = L.lto_direct_qp_dev(h._h, T, N, n, jac.data_ptr(), dfc.data_ptr(), U.data_ptr(), tt.data_ptr(), b0.data_ptr(), bf.data_ptr(), xu.data_ptr(), uu.data_ptr(), None)
        assert rc == 0
    st = torch.cuda.ExternalStream(h.stream, device=dev)
    for _ in range(3):
        qp()
    h.sync()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        a.record()
        for _ in range(10):
            qp()
        b.record()
    h.sync()
    print("   QP kernel: %.3f ms for %d KKT systems of %d unknowns" % (a.elapsed_time(b) / 10, T, (6 + (n == 7)) + (N - 1) * (2 * n + 3) + n + 9))
