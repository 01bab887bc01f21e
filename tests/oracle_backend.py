"""TEST INFRASTRUCTURE: a `backend` for lowthrustopt_b200.solvers that routes every propagation
through the CPU oracle instead of the GPU, so the host loops can be (a) exercised on a CPU-only
machine and (b) run twice -- oracle-backed and GPU-backed -- to compare converged trajectories."""
import numpy as np

from oracle import oracle as O


class OracleBackend:
    def __init__(self, jac="var"):
        O.build()
        self.nt = O.num_threads()
        self.jac = jac          # "var": variational / dual numbers; "fd": the reference's forward differences (direct only)
        self.calls = 0

    @staticmethod
    def _pairs(X, U, t):
        B, N, n = X.shape
        return (X[:, :-1].reshape(-1, n), X[:, 1:].reshape(-1, n), U[:, :-1].reshape(-1, 3), U[:, 1:].reshape(-1, 3),
                t[:, :-1].ravel(), t[:, 1:].ravel())

    def direct_defect(self, X, U, t, nsteps, Isp, MU, DU, TU):
        self.calls += 1
        B, N, n = X.shape
        d, e, st, _ = O.direct_defect(*self._pairs(X, U, t), nsteps=nsteps, dp=O.dparams(MU, DU, TU, Isp), nthreads=self.nt)
        return d.reshape(B, N - 1, n), e.reshape(B, N - 1)

    def direct_blocks(self, X, U, t, nsteps, Isp, MU, DU, TU):
        self.calls += 1
        B, N, n = X.shape
        pr = self._pairs(X, U, t)
        dp = O.dparams(MU, DU, TU, Isp)
        if self.jac == "fd":
            d, e, st, _ = O.direct_defect(*pr, nsteps=nsteps, dp=dp, nthreads=self.nt)
            J = O.direct_jac_fd(*pr, d, nsteps=nsteps, dp=dp, nthreads=self.nt)
        else:
            d, e, J, st = O.direct_jac_var(*pr, nsteps=nsteps, dp=dp, nthreads=self.nt)
        return d.reshape(B, N - 1, n), e.reshape(B, N - 1), J.reshape(B, N - 1, n, 2 * (n + 3))

    @staticmethod
    def _ip(params):
        MU, DU, TU, thrustLimit, mass, td, p, rho = params[:8]
        kw = {"Isp": params[8]} if len(params) > 8 else {}
        return O.iparams(thrustLimit, mass=mass, td=td, p=p, rho=rho, MU_=MU, DU_=DU, TU_=TU, **kw)

    def indirect_defect(self, XC, t, params):
        self.calls += 1
        B, N, m = XC.shape
        xe, st, _, _ = O.indirect_prop(XC[:, :-1].reshape(-1, m), t[:, :-1].ravel(), t[:, 1:].ravel(), self._ip(params), nthreads=self.nt)
        return (xe - XC[:, 1:].reshape(-1, m)).reshape(B, N - 1, m)

    def indirect_blocks(self, XC, t, params):
        self.calls += 1
        B, N, m = XC.shape
        xe, phi, st, _, _ = O.indirect_prop_jac(XC[:, :-1].reshape(-1, m), t[:, :-1].ravel(), t[:, 1:].ravel(), self._ip(params),
                                                nthreads=self.nt)
        return (xe - XC[:, 1:].reshape(-1, m)).reshape(B, N - 1, m), phi.reshape(B, N - 1, m, m)

    def propagate(self, x0, t0, t1, params):
        self.calls += 1
        return O.indirect_prop(x0, t0, t1, self._ip(params), nthreads=self.nt)[0]

    def indirect_solve_batch(self, XC, t, params, thrustLimit=None, rho=None, max_iter=50, flag_adjointsOnly=False):
        """Stand-in for lto_indirect_solve_batch on a machine without a GPU: the host loop, one trajectory after the other."""
        from lowthrustopt_b200 import solvers as S
        MU, DU, TU, tl0, mass, td, p, rho0 = params
        T, N, m = XC.shape
        Xo = np.empty_like(XC); Do = np.empty((T, N - 1, m)); fl = np.empty(T, dtype=np.int32); it = np.empty(T, dtype=np.int32)
        for j in range(T):
            log = []
            X, d, st = S.multiShoot_CRTBP_indirect(XC[j].T, t[j], MU, DU, TU, N, mass, tl0 if thrustLimit is None else float(thrustLimit[j]), False,
                                                   flag_adjointsOnly, max_iter, p, rho0 if rho is None else float(rho[j]), backend=self, log=log)
            Xo[j] = X.T; Do[j] = d.T; fl[j] = st; it[j] = len(log)
        return dict(XC_all=Xo, defect=Do, status_flag=fl, iters=it)

    def direct_qp(self, blocks_cm, defect, u_all, t, b0, bf):
        """Stand-in for lto_direct_qp without a GPU: the host KKT mirror (solvers._qp_direct), one trajectory after the other."""
        from lowthrustopt_b200 import solvers as S
        T, Nm1, nv, n = blocks_cm.shape
        N = Nm1 + 1
        xu = np.zeros((T, N, n)); uu = np.zeros((T, N, 3))
        for j in range(T):
            Jf = S._band_direct(blocks_cm[j].transpose(0, 2, 1), n, N)
            Jf = np.hstack([Jf, np.zeros((Jf.shape[0], 1))])
            tau = (t[j] - t[j, 0]) / (t[j, -1] - t[j, 0]) * 2 - 1
            # b0 / bf are the right-hand sides of the end constraints: feed them through X_all = 0
            mass = b0[j][6] if n == 7 else 1e3
            xr, ur, *_ = S._qp_direct(np.zeros((n, N)), u_all[j].T, np.zeros(3), np.zeros(3), defect[j].T, Jf, n, N, b0[j][:6].copy(), bf[j].copy(), mass,
                                      tau, t[j, 0], t[j, -1], capi_DU, capi_TU, False)
            xu[j] = xr.T; uu[j] = ur.T
        return xu, uu, np.zeros(T, dtype=np.int32)


from lowthrustopt_b200.capi import DU as capi_DU, TU as capi_TU  # noqa: E402
