"""Host-side mirror of the indirect solver's inner closures
(src/multiShoot_CRTBP_indirect.jl:63-146): same names, arguments and return values.
`params` is the reference's tuple (MU, DU, TU, thrustLimit, mass, time_direction, p, rho)
(:260).  All propagation runs in liblto_b200.so on the GPU.
"""
import numpy as np

from . import capi
from .direct import handle


def _params(params, **kw):
    MU, DU, TU, thrustLimit, mass, td, p, rho = params
    if not (p == 0 or p >= 1):
        raise ValueError("Invalid value of p!")                           # CRTBP_stateCostate_deriv.jl:52
    return capi.indirect_params(thrustLimit=thrustLimit, mass=mass, time_direction=td, p=p, rho=rho,
                                MU_=MU, DU_=DU, TU_=TU, **kw)


def defectCalc(XC_all, t_TU, nstate, n_nodes, odefun=None, params=None):
    """(defect1, errors) = defectCalc(...)   multiShoot_CRTBP_indirect.jl:63-90.  errors is all zero (:85)."""
    XC = np.ascontiguousarray(np.asarray(XC_all, dtype=np.float64).T)
    assert XC.shape == (n_nodes, 2 * nstate)
    r = handle().indirect_traj(XC, np.asarray(t_TU, dtype=np.float64), params=_params(params), jac=False)
    return r["defect"].T.copy(), np.zeros(n_nodes - 1)


def jacobianBlocks(XC_all, t_TU, nstate, n_nodes, params):
    """defect and the Phi_i blocks (ForwardDiff.jacobian(f, x0), :121): (n_nodes-1, 2n, 2n)."""
    XC = np.ascontiguousarray(np.asarray(XC_all, dtype=np.float64).T)
    r = handle().indirect_traj(XC, np.asarray(t_TU, dtype=np.float64), params=_params(params), jac=True)
    return r["defect"].T.copy(), r["phi"].transpose(0, 2, 1), r["status"], r["nsteps"]


def jacobianCalc(XC_all, t_TU, nstate, n_nodes, odefun=None, params=None):
    """Jac_full = jacobianCalc(...)   multiShoot_CRTBP_indirect.jl:93-146."""
    _, phi, _, _ = jacobianBlocks(XC_all, t_TU, nstate, n_nodes, params)
    m = 2 * nstate
    Jac_full = np.zeros((m * (n_nodes - 1), n_nodes * m))                 # :127
    eye = np.eye(m)
    for i in range(n_nodes - 1):                                          # :128-138  [Phi_i | -I]
        Jac_full[i * m:(i + 1) * m, i * m:(i + 1) * m] = phi[i]
        Jac_full[i * m:(i + 1) * m, (i + 1) * m:(i + 2) * m] = -eye
    Jac_full[:, :nstate] = 0.0                                            # :141
    Jac_full[:, -2 * nstate:-nstate] = 0.0                                # :142
    return Jac_full
