"""GPU: the device-side Newton update and the batched indirect solver against the
reference's formulation of the same step.

  * lto_indirect_newton  vs  `-sparse(Jac_full) \\ defect_vec` (multiShoot_CRTBP_indirect.jl:127-142, :169-183) restated
    with numpy: band assembly, pinned / removed columns, dense least squares over the structurally non-empty columns.
  * lto_indirect_solve_batch  vs  the host loop of solvers.multiShoot_CRTBP_indirect run once per trajectory
    (same iteration counts, same status flags, converged XC_all within 1e-8 -- BASELINE.json north_star).
"""
import numpy as np
import pytest

from lowthrustopt_b200 import capi, solvers as S, synthetic

pytestmark = pytest.mark.gpu
MU, DU, TU = capi.MU, capi.DU, capi.TU
TOL_TRAJ = 1e-8


def _lstsq_update(phi_rc, d, adjoints_only):
    """The reference's step: phi_rc (N-1, 12, 12) [row, col], d (N-1, 12) -> xc_update (N, 12)."""
    N = phi_rc.shape[0] + 1
    J = S._band_indirect(phi_rc, 6, N)                                    # :127-142
    keep = np.ones(N * 12, dtype=bool)
    if adjoints_only:                                                     # :169-178
        for i in range(N - 1):
            keep[i * 12:i * 12 + 6] = False
    Jk = J[:, keep]
    live = np.any(Jk != 0.0, axis=0)
    sol = np.zeros(Jk.shape[1])
    sol[live] = -np.linalg.lstsq(Jk[:, live], d.ravel(), rcond=None)[0]   # :181-182
    full = np.zeros(N * 12)
    full[keep] = sol
    return full.reshape(N, 12)


@pytest.mark.parametrize("n_traj,n_seg", [(9, 29), (5, 200), (3, 1), (2, 2)])
@pytest.mark.parametrize("adjoints_only", [False, True])
def test_newton_update_vs_reference_least_squares(lto, n_traj, n_seg, adjoints_only):
    c = synthetic.continuation_batch(n_traj=n_traj, n_seg_per_traj=n_seg, ndim=12, seed=11 + n_seg)
    p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
    r = lto.indirect_traj(c["XC_all"], c["t_TU"], params=p, thrustLimit=c["thrustLimit"])
    phi = r["phi"].reshape(n_traj, n_seg, 12, 12)                         # [.., col, row]
    d = r["defect"].reshape(n_traj, n_seg, 12)
    upd, status = lto.indirect_newton(phi, d, adjoints_only)
    assert upd.shape == (n_traj, n_seg + 1, 12) and np.all(status == 0)
    for j in range(n_traj):
        ref = _lstsq_update(phi[j].transpose(0, 2, 1), d[j], adjoints_only)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(upd[j] - ref).max() < 1e-9 * scale, (j, np.abs(upd[j] - ref).max(), scale)
    # structurally empty columns get exact zeros (SuiteSparseQR's basic solution)
    if adjoints_only:
        assert np.all(upd[:, :, :6] == 0.0)
    else:
        assert np.all(upd[:, 0, :6] == 0.0) and np.all(upd[:, -1, :6] == 0.0)


@pytest.mark.parametrize("adjoints_only", [False, True])
def test_newton_resolve_reuses_the_factorisation(lto, adjoints_only):
    """lto_indirect_newton_resolve_dev (the second-order correction's solve, :207): same matrix, new defects, no new factorisation."""
    import torch
    n_traj, n_seg = 7, 200
    c = synthetic.continuation_batch(n_traj=n_traj, n_seg_per_traj=n_seg, ndim=12, seed=5)
    r = lto.indirect_traj(c["XC_all"], c["t_TU"], params=capi.indirect_params(p=2.0, thrustLimit=10.0))
    dev = torch.device("cuda", 0)
    phi = torch.from_numpy(r["phi"]).to(dev); d1 = torch.from_numpy(r["defect"]).to(dev)
    d2 = torch.from_numpy(np.random.default_rng(3).standard_normal(r["defect"].shape) * 1e-3).to(dev)
    u1 = torch.empty((n_traj, n_seg + 1, 12), dtype=torch.float64, device=dev); u2 = torch.empty_like(u1); u2b = torch.empty_like(u1)
    st = torch.zeros(n_traj, dtype=torch.int32, device=dev)
    with pytest.raises(capi.LtoError):
        lto.indirect_newton_resolve_dev(n_traj + 1, n_seg + 1, adjoints_only, d2.data_ptr(), u2.data_ptr())      # no such factorisation held
    lto.indirect_newton_dev(n_traj, n_seg + 1, adjoints_only, phi.data_ptr(), d1.data_ptr(), u1.data_ptr())
    lto.indirect_newton_resolve_dev(n_traj, n_seg + 1, adjoints_only, d2.data_ptr(), u2.data_ptr(), st.data_ptr())
    lto.indirect_newton_dev(n_traj, n_seg + 1, adjoints_only, phi.data_ptr(), d2.data_ptr(), u2b.data_ptr())     # the same solve, factorised afresh
    lto.sync()
    assert int(st.abs().max()) == 0
    a, b = u2.cpu().numpy(), u2b.cpu().numpy()
    assert np.abs(a - b).max() < 1e-11 * max(1.0, np.abs(b).max())
    ref = _lstsq_update(r["phi"].reshape(n_traj, n_seg, 12, 12)[0].transpose(0, 2, 1), d2.cpu().numpy().reshape(n_traj, n_seg, 12)[0], adjoints_only)
    assert np.abs(a[0] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_newton_update_flags_singular_systems(lto):
    phi = np.zeros((2, 4, 12, 12)); d = np.ones((2, 4, 12))
    phi[1] = np.eye(12)                                                   # trajectory 0: Phi = 0 (singular)
    for r in range(6):
        phi[1, :, 6 + r, r] = 1.0                                         # trajectory 1: Phi = [[I, I], [0, I]] ([col, row] storage): well posed
    upd, status = lto.indirect_newton(phi, d, False)
    assert status[0] == 1 and status[1] == 0
    assert np.all(np.isfinite(upd[1]))


@pytest.fixture(scope="module")
def demo_p2(lto):
    """The demo pipeline up to the p = 2 indirect solution (CRTBP_Multishoot_indirect_demo.jl:166-192), for 4 costate seeds."""
    gpu = S.GpuBackend(handle=lto)
    fx = S.demo_fixtures()
    guess = S.trajectory_stack_guess(fx[1], fx[3], backend=gpu)
    XCg, t_TU, tau1, tau2, s0, sf = guess
    out = S.multiShoot_CRTBP_direct(XCg[:6].copy(), np.zeros((3, 30)), tau1, tau2, t_TU, np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3,
                                    2000.0, *fx, backend=gpu)
    X_all = out[0]
    starts, p2 = [], []
    for seed in (42, 43, 44, 45):
        rng = np.random.default_rng(seed)
        XC0 = np.vstack([X_all, 0.1 * rng.standard_normal((6, 30))])
        XC0[:6, 0] = s0[:6]; XC0[:6, -1] = sf[:6]
        XC0[:, 1:-1] += 1e-10 * rng.standard_normal((12, 28))
        starts.append(XC0)
        XC, d, st = S.multiShoot_CRTBP_indirect(XC0, t_TU, MU, DU, TU, 30, 1e3, 10.0, False, False, 50, 2.0, 1.0, backend=gpu)
        assert st == 0
        p2.append(XC)
    return gpu, t_TU, np.stack(starts), np.stack(p2)


def test_host_loop_with_device_newton_matches_dense_least_squares(demo_p2):
    gpu, t_TU, starts, p2 = demo_p2
    la, lb = [], []
    Xa, da, sa = S.multiShoot_CRTBP_indirect(p2[0], t_TU, MU, DU, TU, 30, 1e3, 0.05, False, False, 30, 1.0, 1.0, backend=gpu, log=la)
    Xb, db, sb = S.multiShoot_CRTBP_indirect(p2[0], t_TU, MU, DU, TU, 30, 1e3, 0.05, False, False, 30, 1.0, 1.0, backend=gpu, log=lb,
                                             device_newton=True)
    assert sa == sb == 0 and len(la) == len(lb)
    assert np.abs(Xa - Xb).max() < TOL_TRAJ


def test_solve_batch_matches_per_trajectory_host_loops(demo_p2):
    gpu, t_TU, starts, p2 = demo_p2
    T = p2.shape[0]
    rho = np.array([1.0, 0.7, 1.0, 0.5]); tl = np.array([0.05, 0.05, 0.06, 0.05])
    tt = np.broadcast_to(t_TU, (T, 30)).copy()
    for max_iter in (30, 2):                                             # converging run (line search from iteration 4 on) / cut short
        Xb, db, fb, ib = S.multiShoot_CRTBP_indirect_batch(p2, tt, MU, DU, TU, 30, 1e3, tl, False, max_iter, 1.0, rho, backend=gpu)
        for j in range(T):
            log = []
            Xh, dh, fh = S.multiShoot_CRTBP_indirect(p2[j], t_TU, MU, DU, TU, 30, 1e3, float(tl[j]), False, False, max_iter, 1.0,
                                                     float(rho[j]), backend=gpu, log=log)
            assert fb[j] == fh, (max_iter, j, fb[j], fh)
            assert ib[j] == len(log), (max_iter, j, ib[j], len(log))
            assert np.abs(Xb[j] - Xh).max() < TOL_TRAJ, (max_iter, j, np.abs(Xb[j] - Xh).max())
            if fh == 0:
                assert np.abs(db[j]).max() <= 1e-10                       # the reference's convergence threshold (:280)
    # p = 2 from the random-costate starts: adjoints only (stalls: status 1, :182-186 of the demo), then the full solve
    Xa, da, fa, ia = S.multiShoot_CRTBP_indirect_batch(starts, tt, MU, DU, TU, 30, 1e3, 10.0, True, 10, 2.0, 1.0, backend=gpu)
    for j in range(T):
        Xh, dh, fh = S.multiShoot_CRTBP_indirect(starts[j], t_TU, MU, DU, TU, 30, 1e3, 10.0, False, True, 10, 2.0, 1.0, backend=gpu)
        assert fa[j] == fh == 1
        assert np.abs(Xa[j] - Xh).max() < 1e-7, (j, np.abs(Xa[j] - Xh).max())
    Xf, df, ff, itf = S.multiShoot_CRTBP_indirect_batch(Xa, tt, MU, DU, TU, 30, 1e3, 10.0, False, 50, 2.0, 1.0, backend=gpu)
    assert np.all(ff == 0)
    assert np.abs(Xf - p2).max() < 1e-7                                   # same p = 2 solutions as the host pipeline (different path)


def test_solve_batch_continuation_slab(lto):
    """A slab of BASELINE configs[4] (continuation batch, 201 nodes): every trajectory either converges to 1e-10 or is flagged."""
    gpu = S.GpuBackend(handle=lto)
    c = synthetic.continuation_batch(n_traj=16, n_seg_per_traj=200, ndim=12)
    XC = c["XC_all"].copy(); XC[:, :, 6:] *= 0.01                         # small costates: near-ballistic start
    r = lto.indirect_solve_batch(XC, c["t_TU"], params=capi.indirect_params(p=2.0, thrustLimit=10.0), max_iter=12)
    ok = r["status_flag"] == 0
    assert ok.sum() >= 8
    assert np.all(r["er"][ok] <= 1e-10)
    assert np.all(r["XC_all"][:, 0, :6] == XC[:, 0, :6]) and np.all(r["XC_all"][:, -1, :6] == XC[:, -1, :6])   # pinned end states (:324-325)


def test_continuation_batch_matches_per_trajectory_ladders(demo_p2):
    """reduceFuel_indirect (HelperFunctions.jl:105-193): the batched driver (one lto_indirect_solve_batch call per round) against the
    one-trajectory mirror run once per trajectory (host loop, dense least squares)."""
    gpu, t_TU, starts, p2 = demo_p2
    T = p2.shape[0]
    p1 = np.stack([S.multiShoot_CRTBP_indirect(p2[j], t_TU, MU, DU, TU, 30, 1e3, 0.05, False, False, 30, 1.0, 1.0, backend=gpu)[0] for j in range(T)])
    tt = np.broadcast_to(t_TU, (T, 30)).copy()
    target = np.array([1e-2, 0.125, 0.05, 0.5])
    Xb, db, sb, rounds = S.reduceFuel_indirect_batch(p1, tt, MU, DU, TU, 30, 1e3, 0.05, 1.0, target, backend=gpu)
    assert np.all(sb == 0) and rounds == 8                                  # 1 -> 1e-2 by halving: 1 + 7 solver calls for the longest ladder
    for j in range(T):
        Xh, dh, sh = S.reduceFuel_indirect(p1[j], t_TU, MU, DU, TU, 30, 1e3, 0.05, 1.0, float(target[j]), backend=gpu)
        assert sh == 0
        assert np.abs(Xb[j] - Xh).max() < TOL_TRAJ, (j, np.abs(Xb[j] - Xh).max())
        assert np.abs(db[j]).max() <= 1e-10


def test_solve_batch_edge_cases(lto):
    p = capi.indirect_params(p=2.0, thrustLimit=10.0)
    c = synthetic.continuation_batch(n_traj=3, n_seg_per_traj=200, ndim=12)
    c = dict(XC_all=c["XC_all"][:, :2].copy(), t_TU=c["t_TU"][:, :2].copy())         # two nodes: one (short) segment per trajectory
    r = lto.indirect_solve_batch(c["XC_all"], c["t_TU"], params=p, max_iter=0)      # maxIter = 0: "Reached max iteration count" at once (:282-286)
    assert np.all(r["status_flag"] == 1) and np.all(r["iters"] == 0) and np.array_equal(r["XC_all"], c["XC_all"])
    r = lto.indirect_solve_batch(c["XC_all"], c["t_TU"], params=p, max_iter=20)
    assert np.all(r["status_flag"] == 0) and np.all(r["er"] <= 1e-10)
    assert np.array_equal(r["XC_all"][:, :, :6], c["XC_all"][:, :, :6])             # both nodes' states are pinned: only costates move
    r = lto.indirect_solve_batch(np.zeros((0, 5, 12)), np.zeros((0, 5)), params=p)  # empty batch
    assert r["XC_all"].shape == (0, 5, 12) and r["status_flag"].shape == (0,)
    with pytest.raises(capi.LtoError):
        lto.indirect_solve_batch(np.zeros((1, 1, 12)), np.zeros((1, 1)), params=p)  # n_nodes < 2
    bad = c["XC_all"].copy(); bad[1, 0, 0] = np.nan
    r = lto.indirect_solve_batch(bad, c["t_TU"], params=p, max_iter=5)
    assert r["status_flag"][1] == 2 and r["status_flag"][0] == 0 and r["status_flag"][2] == 0   # isnan(XC_all[1]) -> 2 (:339-341)


def test_nan_inside_a_batch_is_never_reported_as_converged(lto):
    """the reference tests XC_all[1,1] only (:339), a pinned entry that the update never touches, so a trajectory whose
    interior went NaN came back with status_flag 0 and the continuation driver walked on from it.  Here: a NaN in an interior
    (non-pinned) costate of one trajectory of a batch -> that trajectory reports 2, its neighbours converge, and the batched
    continuation ladder does not accept it."""
    p = capi.indirect_params(p=2.0, thrustLimit=10.0)
    c = synthetic.continuation_batch(n_traj=4, n_seg_per_traj=29, ndim=12)
    XC = c["XC_all"].copy(); XC[:, :, 6:] *= 0.1
    XC[2, 7, 9] = np.nan                                                         # interior node, costate component
    r = lto.indirect_solve_batch(XC, c["t_TU"], params=p, max_iter=8)
    assert r["status_flag"][2] == 2 and np.all(r["status_flag"][[0, 1, 3]] == 0)
    assert not np.isnan(XC[2, 0, 0])                                             # the entry the reference looks at is fine
    # host mirror, same input
    out, d, st = S.multiShoot_CRTBP_indirect(XC[2].T.copy(), c["t_TU"][2], MU, DU, TU, 30, 1e3, 10.0, False, False, 8, 2.0, 1.0,
                                             backend=S.GpuBackend(handle=lto))
    assert st == 2
