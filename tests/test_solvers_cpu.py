"""CPU: the host-side mirror of the reference's outer solvers (lowthrustopt_b200/solvers.py), driven through the
ORACLE backend (tests/oracle_backend.py) because there is no GPU here.  Checks the host logic -- band assembly,
QP / least-squares update, SOC, line search, end-state pinning -- against the behavioural anchors of
SURVEY.md Appendix C (BASELINE configs 1-2, the two demo scripts)."""
import numpy as np
import pytest

from lowthrustopt_b200 import capi, solvers as S
from oracle_backend import OracleBackend

MU, DU, TU = capi.MU, capi.DU, capi.TU


@pytest.fixture(scope="module")
def demo():
    be = OracleBackend()
    X0t, X0, Xft, Xf = S.demo_fixtures()
    XC, t_TU, tau1, tau2, s0, sf = S.trajectory_stack_guess(X0, Xf, backend=be)
    return dict(be=be, fx=(X0t, X0, Xft, Xf), XC=XC, t=t_TU, tau1=tau1, tau2=tau2, s0=s0, sf=sf)


def test_trajectory_stack_guess(demo):
    # CRTBP_Multishoot_direct_demo.jl:117-157; anchors SURVEY App. C
    assert demo["XC"].shape == (12, 30) and abs(demo["tau2"] - 0.274) < 1e-12
    d, _ = demo["be"].direct_defect(demo["XC"][:6].T[None], np.zeros((1, 30, 3)), demo["t"][None], 10, 2000.0, MU, DU, TU)
    assert abs(np.abs(d).max() - 2.6265e-2) < 2e-6                            # initial max defect of the direct demo


def run_direct(demo, be, nstate=6):
    X0t, X0, Xft, Xf = demo["fx"]
    X_all = demo["XC"][:6].copy()
    if nstate == 7:
        X_all = np.vstack([X_all, 1000.0 * np.ones((1, 30))])
    log = []
    out = S.multiShoot_CRTBP_direct(X_all, np.zeros((3, 30)), demo["tau1"], demo["tau2"], demo["t"], np.zeros(3), np.zeros(3), MU, DU, TU,
                                    30, 10, 1e3, 2000.0, X0t, X0, Xft, Xf, False, False, 0.0, False, 100, backend=be, log=log)
    return out, log


def test_direct_demo_converges_like_the_anchors(demo):
    out, log = run_direct(demo, demo["be"])
    ers = [l["er"] for l in log]
    assert len(ers) == 4 and ers[-1] < 1e-6                                   # multiShoot_CRTBP_direct.jl:491
    for got, want in zip(ers, (2.67e-3, 1.40e-5, 1.71e-6, 2.65e-10)):
        assert abs(got - want) / want < 0.01
    assert abs(log[-1]["cost"] - 0.00521) < 1e-5
    X_all, u_all = out[0], out[1]
    assert abs(np.linalg.norm(u_all, axis=0).max() - 0.0526) < 1e-4
    assert np.allclose(X_all[:, 0], demo["s0"][:6]) and np.allclose(X_all[:, -1], demo["sf"][:6])   # hard-fixed endpoints (:374-375)


def test_direct_variational_jacobian_is_a_drop_in_for_the_reference_fd(demo):
    """Same solver, Jacobian from the reference's own forward differences (pert 1e-8): same iteration history."""
    _, log_var = run_direct(demo, demo["be"])
    out_fd, log_fd = run_direct(demo, OracleBackend(jac="fd"))
    assert len(log_fd) == len(log_var)
    for a, b in zip(log_var[:3], log_fd[:3]):
        assert abs(a["er"] - b["er"]) / a["er"] < 1e-3


def test_direct_nstate7_mass_row(demo):
    out, log = run_direct(demo, demo["be"], nstate=7)
    assert log[-1]["er"] < 1e-6 and len(log) <= 6
    m = out[0][6]
    assert m[0] == 1000.0 and np.all(np.diff(m) < 0) and 995 < m[-1] < 1000


def test_direct_batch_driver_matches_the_one_trajectory_mirror(demo):
    """multiShoot_CRTBP_direct_batch (batched calls, per-trajectory stopping) against multiShoot_CRTBP_direct, both oracle-backed."""
    rng = np.random.default_rng(12)
    Xs = np.stack([demo["XC"][:6] + (2e-4 * j) * rng.standard_normal((6, 30)) for j in range(2)])
    Us = np.zeros((2, 3, 30)); tb = np.stack([demo["t"]] * 2)
    Xb, Ub, db, itb = S.multiShoot_CRTBP_direct_batch(Xs, Us, demo["tau1"], demo["tau2"], tb, MU, DU, TU, 30, 10, 1e3, 2000.0, *demo["fx"], backend=demo["be"])
    for j in range(2):
        log = []
        out = S.multiShoot_CRTBP_direct(Xs[j], Us[j], demo["tau1"], demo["tau2"], demo["t"], np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3, 2000.0,
                                        *demo["fx"], backend=demo["be"], log=log)
        assert itb[j] == len(log) and np.abs(Xb[j] - out[0]).max() < 1e-12 and np.abs(Ub[j] - out[1]).max() < 1e-12
        assert np.abs(db[j]).max() <= 1e-6


def test_direct_flagEnd_is_refused(demo):
    X0t, X0, Xft, Xf = demo["fx"]
    with pytest.raises(NotImplementedError):
        S.multiShoot_CRTBP_direct(demo["XC"][:6], np.zeros((3, 30)), 0.75, 0.274, demo["t"], np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10,
                                  1e3, 2000.0, X0t, X0, Xft, Xf, False, True, backend=demo["be"])


def test_indirect_demo_sequence(demo):
    be = demo["be"]
    (X_all, *_), _ = run_direct(demo, be)
    rng = np.random.default_rng(42)
    XC = np.vstack([X_all, 0.1 * rng.standard_normal((6, 30))])                # CRTBP_Multishoot_indirect_demo.jl:166-176
    XC[:6, 0] = demo["s0"][:6]; XC[:6, -1] = demo["sf"][:6]
    XC[:, 1:-1] += 1e-10 * rng.standard_normal((12, 28))
    log = []
    XC, d, st = S.multiShoot_CRTBP_indirect(XC, demo["t"], MU, DU, TU, 30, 1e3, 10.0, False, True, 10, 2.0, 1.0, backend=be, log=log)
    assert st == 1 and abs(log[-1]["er"] - 1.14e-4) < 1e-6                     # adjoints-only stalls: states cannot move
    log = []
    XC, d, st = S.multiShoot_CRTBP_indirect(XC, demo["t"], MU, DU, TU, 30, 1e3, 10.0, False, False, 50, 2.0, 1.0, backend=be, log=log)
    assert st == 0 and len(log) == 1 and log[0]["er"] < 1e-10
    peak = max(np.linalg.norm(S.controlLaw_cart(XC[9:12, i], 10.0, 2.0, 1.0, 1e3)) for i in range(30))
    assert abs(peak - 0.0547) < 2e-4
    log = []
    XC, d, st = S.multiShoot_CRTBP_indirect(XC, demo["t"], MU, DU, TU, 30, 1e3, 0.05, False, False, 30, 1.0, 1.0, backend=be, log=log)
    assert st == 0 and len(log) == 4 and abs(log[0]["er"] - 6.93e-2) < 1e-4 and log[-1]["er"] < 1e-10
    assert np.array_equal(XC[:6, 0], demo["s0"][:6]) and np.array_equal(XC[:6, -1], demo["sf"][:6])   # :324-325
    # rho-continuation (HelperFunctions.jl:105-193), three halvings
    clog = []
    XC2, d2, st2 = S.reduceFuel_indirect(XC, demo["t"], MU, DU, TU, 30, 1e3, 0.05, 1.0, 0.125, backend=be, log=clog)
    assert st2 == 0 and [c["rho"] for c in clog] == [1.0, 0.5, 0.25, 0.125] and all(c["status"] == 0 for c in clog)
    # the batched driver walks every trajectory's own ladder (different targets) and ends where the one-trajectory mirror ends
    XC3, d3, st3 = S.reduceFuel_indirect(XC, demo["t"], MU, DU, TU, 30, 1e3, 0.05, 1.0, 0.5, backend=be)
    blog = []
    Xb, db, sb, rounds = S.reduceFuel_indirect_batch(np.stack([XC, XC]), np.stack([demo["t"]] * 2), MU, DU, TU, 30, 1e3, 0.05, 1.0,
                                                     np.array([0.125, 0.5]), backend=be, log=blog)
    assert rounds == 4 and [b["n"] for b in blog] == [2, 2, 1, 1] and np.all(sb == 0)
    assert np.abs(Xb[0] - XC2).max() < 1e-12 and np.abs(Xb[1] - XC3).max() < 1e-12


def test_continuation_ladder_backs_off_like_the_reference():
    """The ladder logic alone (HelperFunctions.jl:155-187) with a scripted solver: success halves rho, failure multiplies it by
    3 (1 + rand) and restarts from the last converged trajectory."""
    class R:
        def random(self):
            return 0.5
    g = S._reduceFuel_steps(np.zeros((12, 3)), 1.0, 0.1, R())
    X, rho = next(g); seen = [rho]
    script = [0, 0, 1, 0, 0, 0, 0, 0]                                          # the third call (rho 0.25) fails once
    out = None
    for k, st in enumerate(script):
        try:
            X, rho = g.send((np.full((12, 3), float(k + 1)), None, st))
        except StopIteration as fin:
            out = fin.value
            break
        seen.append(rho)
        if st != 0:
            assert np.all(X == float(k))                                       # restart from the last converged XC_all, not the failed one
    assert seen[:6] == [1.0, 0.5, 0.25, 0.25 * 4.5, 0.5625, 0.28125] and seen[-1] == 0.1 and out[2] == 0


def test_invalid_p_and_nan_status():
    be = OracleBackend()
    XC = np.zeros((12, 3)); XC[0] = 1.0; XC[9] = 0.1
    with pytest.raises(ValueError, match="Invalid value of p"):
        S.GpuBackend._ip(None, (MU, DU, TU, 1.0, 1e3, 1.0, 0.5, 1.0))
    XC[:, 0] = np.nan
    out, d, st = S.multiShoot_CRTBP_indirect(XC, np.array([0.0, 0.1, 0.2]), MU, DU, TU, 3, 1e3, 0.05, False, False, 3, 1.0, 1.0, backend=be)
    assert st == 2                                                              # isnan(XC_all[1]) -> status_flag 2 (:339-341)
    # a NaN in an interior, non-pinned entry: XC_all[1,1] stays finite, which the reference would report as "converged" 
    XC = np.zeros((12, 4)); XC[0] = 1.0; XC[9] = 0.1; XC[10, 2] = np.nan
    out, d, st = S.multiShoot_CRTBP_indirect(XC, np.array([0.0, 0.1, 0.2, 0.3]), MU, DU, TU, 4, 1e3, 0.05, False, False, 3, 1.0, 1.0, backend=be)
    assert st == 2 and np.isfinite(out[0, 0])


def test_densify_and_jacobi_constant(demo):
    """densify (HelperFunctions.jl:51-101) on a ballistic trajectory whose nodes lie on ONE solution: every dense column equals
    the propagation from the first node (1e-10), there are exactly n_desired columns on LinRange(t_1, t_N, n_desired), node times
    return the node itself, and the Jacobi constant (HelperFunctions.jl:10-15) is conserved along it."""
    be = demo["be"]
    X0 = S.demo_fixtures()[1]
    params = (MU, DU, TU, 0.0, 1e3, 1.0, 2.0, 1.0)                         # thrustLimit 0: ballistic, costates ride along
    N = 12
    t_TU = np.linspace(0.0, 1.5, N)
    x0 = np.concatenate([X0[:, 3], 0.1 * np.ones(6)])
    XC = np.empty((12, N)); XC[:, 0] = x0
    XC[:, 1:] = be.propagate(np.tile(x0, (N - 1, 1)), np.zeros(N - 1), t_TU[1:], params).T
    XC_dense, t_dense = S.densify(XC, t_TU, params, 56, backend=be)         # 56 = 5 * 11 + 1: the nodes are dense times too
    assert XC_dense.shape == (12, 56) and np.array_equal(t_dense, np.linspace(0.0, 1.5, 56))
    want = be.propagate(np.tile(x0, (56, 1)), np.zeros(56), t_dense, params).T
    assert np.abs(XC_dense - want).max() < 1e-10
    assert np.array_equal(XC_dense[:, 0], XC[:, 0]) and np.abs(XC_dense[:, 5] - XC[:, 1]).max() < 1e-13
    C = S.jacobiConstant(XC_dense[:6], MU, DU)
    assert C.shape == (56,) and np.abs(C - C[0]).max() < 1e-11 and abs(C[0] - 3.0327) < 1e-3   # C = 3.0327000 on L2_Anderson_1
    # nodes that do NOT lie on one solution: a dense point belongs to the segment that starts at or before it
    XC2 = XC.copy(); XC2[0, 4] += 1e-3
    D2, _ = S.densify(XC2, t_TU, params, 56, backend=be)
    assert np.array_equal(D2[:, :20], XC_dense[:, :20]) and np.array_equal(D2[:, 20], XC2[:, 4]) and np.abs(D2[0, 21] - XC_dense[0, 21]) > 1e-4
    assert np.array_equal(D2[:, 25:], XC_dense[:, 25:])                     # the end column comes from the last segment's propagation


def test_bangbang_fixture_is_a_converged_rho_1e4_solution(oracle):
    """tests/golden/bangbang_v1.json (made by make_bangbang.py) holds the demo's continuation down to the reference's target
    rho = 1e-4 (CRTBP_Multishoot_indirect_demo.jl:276-281): each stored trajectory satisfies its own defect constraints to the
    reference's threshold (multiShoot_CRTBP_indirect.jl:280) under the oracle, and the last one is bang-bang."""
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "bangbang_v1.json")) as f:
        g = json.load(f)
    t = np.array(g["t_TU"])
    be = OracleBackend()
    for key, rho in (("0.01", 1e-2), ("0.001", 1e-3), ("0.0001", 1e-4)):
        XC = np.array(g["XC_nodes"][key])
        d = be.indirect_defect(XC[None], t[None], (MU, DU, TU, g["thrustLimit"], g["mass"], 1.0, g["p"], rho))
        assert np.abs(d).max() < 1e-10, (key, np.abs(d).max())
    lv = np.linalg.norm(XC[:, 9:12], axis=1)
    assert (lv > 1.0).any() and (lv < 1.0).any()
    # the long-double truth sees the same defects to 1e-8: what the 1e-13 state-only controller (robust estimate + sharp-law
    # safeguard) leaves at rho = 1e-4
    xt, st = oracle.indirect_prop_ld(XC[:-1], t[:-1], t[1:], oracle.iparams(g["thrustLimit"], p=1.0, rho=1e-4), atol=1e-18, rtol=1e-18,
                                     nthreads=oracle.num_threads())
    assert np.abs(xt - XC[1:]).max() < 5e-8


def _guess14(oracle, rho=1e-2, seed=3):
    """A 14-dim start for the indirect loop ([r v m | lr lv lm], BASELINE configs[1]): the converged 12-dim trajectory of the
    bang-bang fixture with a mass history propagated segment by segment (so masses and lm are consistent with 12-dim costates that
    are NOT a 14-dim solution: the thrust acceleration now follows the falling mass), interior costates perturbed."""
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "bangbang_v1.json")) as f:
        g = json.load(f)
    XC12 = np.array(g["XC_nodes"]["%g" % rho]).T; t = np.array(g["t_TU"]); N = XC12.shape[1]
    XC = np.zeros((14, N)); XC[:6] = XC12[:6]; XC[7:13] = XC12[6:12]; XC[6] = 1000.0
    ip = oracle.iparams(g["thrustLimit"], p=1.0, rho=rho, Isp=2000.0)
    for i in range(N - 1):
        xe = oracle.indirect_prop(XC[:, i:i + 1].T.copy(), t[i:i + 1], t[i + 1:i + 2], ip)[0]
        XC[6, i + 1] = xe[0, 6]; XC[13, i + 1] = xe[0, 13]
    XC[13] -= XC[13, -1]                                                       # lm(t_f) = 0 is the value held at the last node
    XC[7:13, 1:-1] += 1e-3 * np.random.default_rng(seed).standard_normal((6, N - 2))
    return XC, t, g["thrustLimit"]


def test_indirect_loop_on_the_14_dim_system(oracle):
    """BASELINE configs[1] as north_star words it -- state + costate + mass (14-dim), 14 x 14 STM, single trajectory, Newton to
    convergence -- through the host loop with the oracle backend: converges to the reference's threshold (:280); r, v, m of the first
    node and r, v, lm of the last stay where they were put (:324-325 with 7-element pins); the final mass is an OUTPUT."""
    XC0, t, tl = _guess14(oracle)
    log = []
    XC, d, st = S.multiShoot_CRTBP_indirect(XC0.copy(), t, MU, DU, TU, 30, 1000.0, tl, False, False, 30, 1.0, 1e-2, backend=OracleBackend(), log=log)
    assert st == 0 and log[-1]["er"] <= 1e-10 and 2 <= len(log) <= 12
    assert np.array_equal(XC[:7, 0], XC0[:7, 0]) and np.array_equal(XC[:6, -1], XC0[:6, -1]) and XC[13, -1] == XC0[13, -1]
    assert 990.0 < XC[6, -1] < 1000.0 and np.all(np.diff(XC[6]) <= 0.0)       # propellant is spent, never gained
    assert abs(XC[6, -1] - XC0[6, -1]) > 1e-6                                 # ... and the final mass moved: it is solved for, not held
    with pytest.raises(ValueError):
        S.multiShoot_CRTBP_indirect(np.zeros((10, 5)), np.linspace(0, 1, 5), MU, DU, TU, 5, 1e3, 0.05, backend=OracleBackend())
