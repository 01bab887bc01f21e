"""Host-side mirror of the direct solver's inner closures.

Same names, argument order and return values as the reference's
`defectCalc` / `jacobianCalc` (src/multiShoot_CRTBP_direct.jl:66-166); arrays use the
reference's shapes (nstate x n_nodes, 3 x n_nodes).  Every propagation runs in
liblto_b200.so on the GPU -- there is no CPU implementation behind these.
"""
import numpy as np

from . import capi

_handle = None


def handle(device=0):
    global _handle
    if _handle is None:
        _handle = capi.Handle(device)
    return _handle


def set_handle(h):
    global _handle
    _handle = h


def _params(Isp, MU, DU, TU, mode=capi.LTO_FIXED, tol=1e-13):
    return capi.direct_params(Isp=Isp, mode=mode, tol=tol, MU_=MU, DU_=DU, TU_=TU)


def defectCalc(X_all, u_all, t_TU, nstate, n_nodes, nsteps, Isp, odefun=None, MU=capi.MU, DU=capi.DU, TU=capi.TU):
    """(defect1, errors) = defectCalc(...)   multiShoot_CRTBP_direct.jl:66-109.
    `odefun` is accepted for signature parity; the RHS is CRTBP_prop_EP_deriv (:472), compiled in."""
    X = np.ascontiguousarray(np.asarray(X_all, dtype=np.float64).T)      # (n_nodes, nstate): Julia's memory order
    U = np.ascontiguousarray(np.asarray(u_all, dtype=np.float64).T)
    assert X.shape == (n_nodes, nstate) and U.shape == (n_nodes, 3)
    r = handle().direct_traj(X, U, np.asarray(t_TU, dtype=np.float64), nsteps=nsteps, params=_params(Isp, MU, DU, TU), jac=False)
    return r["defect"].T.copy(), r["errors"]


def jacobianBlocks(X_all, u_all, t_TU, nstate, n_nodes, nsteps, Isp, MU=capi.MU, DU=capi.DU, TU=capi.TU):
    """defect, errors and the dense per-segment blocks of Jac_temp (:122,:139-140):
    Jac_temp[(i-1)n+1 : i n, :] = blocks[i-1]  (nstate x 2(nstate+3))."""
    X = np.ascontiguousarray(np.asarray(X_all, dtype=np.float64).T)
    U = np.ascontiguousarray(np.asarray(u_all, dtype=np.float64).T)
    r = handle().direct_traj(X, U, np.asarray(t_TU, dtype=np.float64), nsteps=nsteps, params=_params(Isp, MU, DU, TU), jac=True)
    return r["defect"].T.copy(), r["errors"], r["jac"].transpose(0, 2, 1)


def jacobianCalc(X_all, u_all, t_TU, defect, nstate, n_nodes, nsteps, Isp, odefun=None, pert=1e-8,
                 MU=capi.MU, DU=capi.DU, TU=capi.TU):
    """Jac_full = jacobianCalc(...)   multiShoot_CRTBP_direct.jl:111-166.
    `defect` and `pert` are unused: the blocks come from the variational equations, the exact
    derivative of the discrete map the reference differences with pert = 1e-8."""
    _, _, blocks = jacobianBlocks(X_all, u_all, t_TU, nstate, n_nodes, nsteps, Isp, MU, DU, TU)
    n = nstate
    Jac_full = np.zeros((n * (n_nodes - 1), n_nodes * (n + 3)))           # :146
    for i in range(n_nodes - 1):                                          # :147-162
        rows = slice(i * n, (i + 1) * n)
        Jac_full[rows, i * n:(i + 2) * n] = blocks[i][:, :2 * n]          # :159
        c0 = n * n_nodes + 3 * i
        Jac_full[rows, c0:c0 + 6] = blocks[i][:, 2 * n:]                  # :161
    return Jac_full
