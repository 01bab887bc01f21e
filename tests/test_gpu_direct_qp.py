"""GPU: the direct solver's linear subproblem on the device, lto_direct_qp, against the host mirror of
optimizeTraj (solvers._qp_direct: the dense KKT system of the same equality-constrained QP, multiShoot_CRTBP_direct.jl:248-403)."""
import numpy as np
import pytest

from lowthrustopt_b200 import capi, solvers as S

pytestmark = pytest.mark.gpu
MU, DU, TU = capi.MU, capi.DU, capi.TU


@pytest.fixture(scope="module")
def demo(lto):
    gpu = S.GpuBackend(handle=lto)
    fx = S.demo_fixtures()
    return gpu, fx, S.trajectory_stack_guess(fx[1], fx[3], backend=gpu)


@pytest.mark.parametrize("n", [6, 7])
def test_direct_qp_vs_host_kkt(demo, lto, n):
    gpu, fx, (XC, t_TU, tau1, tau2, s0, sf) = demo
    N, T = 30, 5
    rng = np.random.default_rng(8)
    tau = (t_TU - t_TU[0]) / (t_TU[-1] - t_TU[0]) * 2 - 1
    state_0, state_f = S.interpEndStates(tau1, tau2, *fx, MU)
    Xs, Us, refs = [], [], []
    for j in range(T):
        X = XC[:6] + 1e-3 * rng.standard_normal((6, N))
        if n == 7:
            X = np.vstack([X, 1000.0 - 0.01 * np.arange(N)[None]])
        U = 0.02 * rng.standard_normal((3, N))
        Xs.append(X); Us.append(U)
    Xb = np.stack([x.T for x in Xs]); Ub = np.stack([u.T for u in Us]); tb = np.broadcast_to(t_TU, (T, N)).copy()
    r = lto.direct_traj(Xb, Ub, tb, nsteps=10, params=capi.direct_params(Isp=2000.0), jac=True)
    jac = r["jac"].reshape(T, N - 1, 2 * (n + 3), n); dfc = r["defect"].reshape(T, N - 1, n)
    b0 = np.stack([np.concatenate([state_0 - Xs[j][:6, 0], [1e3 - Xs[j][6, 0]] if n == 7 else []]) for j in range(T)])
    bf = np.stack([state_f - Xs[j][:6, -1] for j in range(T)])
    xu, uu, st = lto.direct_qp(jac, dfc, Ub, tb, b0, bf)
    assert np.all(st == 0)
    for j in range(T):
        Jf = S._band_direct(jac[j].transpose(0, 2, 1), n, N)
        Jf = np.hstack([Jf, np.zeros((Jf.shape[0], 1))])
        xr, ur, *_ = S._qp_direct(Xs[j], Us[j], np.zeros(3), np.zeros(3), dfc[j].T, Jf, n, N, state_0, state_f, 1e3, tau, t_TU[0], t_TU[-1], DU, TU, False)
        assert np.abs(xu[j].T - xr).max() < 1e-10 * max(1.0, np.abs(xr).max()), (j, np.abs(xu[j].T - xr).max())
        assert np.abs(uu[j].T - ur).max() < 1e-10 * max(1.0, np.abs(ur).max()), (j, np.abs(uu[j].T - ur).max())


def test_direct_solver_with_device_qp_matches_host_qp(demo):
    gpu, fx, (XC, t_TU, tau1, tau2, s0, sf) = demo
    for n in (6, 7):
        X_all = XC[:6].copy()
        if n == 7:
            X_all = np.vstack([X_all, 1000.0 * np.ones((1, 30))])
        outs, logs = [], []
        for dev in (False, True):
            log = []
            out = S.multiShoot_CRTBP_direct(X_all, np.zeros((3, 30)), tau1, tau2, t_TU, np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3, 2000.0,
                                            *fx, backend=gpu, log=log, device_qp=dev)
            outs.append(out); logs.append(log)
        assert len(logs[0]) == len(logs[1]) and logs[1][-1]["er"] < 1e-6
        assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-8 and np.abs(outs[0][1] - outs[1][1]).max() < 1e-8
        assert abs(logs[0][-1]["cost"] - logs[1][-1]["cost"]) < 1e-10


def test_direct_qp_flags_singular_input(lto):
    T, N, n = 2, 6, 6
    jac = np.zeros((T, N - 1, 18, 6)); dfc = np.zeros((T, N - 1, 6)); U = np.zeros((T, N, 3)); t = np.tile(np.linspace(0, 1, N), (T, 1))
    xu, uu, st = lto.direct_qp(jac, dfc, U, t, np.zeros((T, 6)), np.zeros((T, 6)))
    assert np.all(st == 1)


def test_direct_batch_matches_per_trajectory_solves(demo):
    """multiShoot_CRTBP_direct_batch (every heavy step batched on the device) against the one-trajectory mirror with the host QP."""
    gpu, fx, (XC, t_TU, tau1, tau2, s0, sf) = demo
    rng = np.random.default_rng(12)
    T = 4
    for n in (6, 7):
        Xs = np.stack([XC[:6] + (2e-4 * j) * rng.standard_normal((6, 30)) for j in range(T)])
        if n == 7:
            Xs = np.concatenate([Xs, 1000.0 * np.ones((T, 1, 30))], axis=1)
        Us = np.zeros((T, 3, 30)); tb = np.broadcast_to(t_TU, (T, 30)).copy()
        Xb, Ub, db, itb = S.multiShoot_CRTBP_direct_batch(Xs, Us, tau1, tau2, tb, MU, DU, TU, 30, 10, 1e3, 2000.0, *fx, backend=gpu)
        for j in range(T):
            log = []
            out = S.multiShoot_CRTBP_direct(Xs[j], Us[j], tau1, tau2, t_TU, np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3, 2000.0, *fx, backend=gpu, log=log)
            assert itb[j] == len(log), (n, j, itb[j], len(log))
            assert np.abs(Xb[j] - out[0]).max() < 1e-8 and np.abs(Ub[j] - out[1]).max() < 1e-8, (n, j)
            assert np.abs(db[j]).max() <= 1e-6


def test_direct_solve_batch_resident_loop(demo, lto):
    """lto_direct_solve_batch (the SQP loop resident on the device) against the one-trajectory mirror with the host QP."""
    gpu, fx, (XC, t_TU, tau1, tau2, s0, sf) = demo
    rng = np.random.default_rng(21)
    T = 6
    state_0, state_f = S.interpEndStates(tau1, tau2, *fx, MU)
    for n in (6, 7):
        Xs = np.stack([XC[:6] + (2e-4 * j) * rng.standard_normal((6, 30)) for j in range(T)])
        if n == 7:
            Xs = np.concatenate([Xs, 1000.0 * np.ones((T, 1, 30))], axis=1)
        Us = np.zeros((T, 3, 30))
        for max_iter in (100, 2):
            r = lto.direct_solve_batch(Xs.transpose(0, 2, 1), Us.transpose(0, 2, 1), np.stack([t_TU] * T), np.stack([state_0] * T), np.stack([state_f] * T),
                                       mass=1e3, nsteps=10, max_iter=max_iter, params=capi.direct_params(Isp=2000.0))
            for j in range(T):
                log = []
                out = S.multiShoot_CRTBP_direct(Xs[j], Us[j], tau1, tau2, t_TU, np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3, 2000.0, *fx, False, False,
                                                0.0, False, max_iter, backend=gpu, log=log)
                assert r["iters"][j] == len(log), (n, max_iter, j, r["iters"][j], len(log))
                assert np.abs(r["X_all"][j].T - out[0]).max() < 1e-8 and np.abs(r["u_all"][j].T - out[1]).max() < 1e-8, (n, max_iter, j)
                assert np.abs(r["defect"][j].T - out[7]).max() < 1e-9
            if max_iter == 100:
                assert np.all(r["er"] <= 1e-6)
    r = lto.direct_solve_batch(np.zeros((0, 5, 6)), np.zeros((0, 5, 3)), np.zeros((0, 5)), np.zeros((0, 6)), np.zeros((0, 6)))
    assert r["X_all"].shape == (0, 5, 6)
