"""ctypes wrapper of the CPU ORACLE (oracle/liblto_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under lowthrustopt_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MU = 0.012150585609624037   # src/LowThrustOpt.jl:24
DU = 384747.96285603708     # src/LowThrustOpt.jl:25
TU = 375699.81732246041     # src/LowThrustOpt.jl:26


def build(force=False):
    so = os.path.join(_HERE, "liblto_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("lto_oracle_capi.cpp", "lto_oracle.hpp", "Makefile")]
    src += [os.path.join(_HERE, "..", "lowthrustopt_b200", "csrc", f) for f in ("lto_prop_generic.cuh", "lto_math.cuh", "lto_tableau.h")]   # equal-algorithm baseline
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return int(lib().oracle_num_threads())


def dparams(MU_=MU, DU_=DU, TU_=TU, Isp=2000.0):
    return np.array([MU_, DU_, TU_, Isp], dtype=np.float64)


def iparams(thrustLimit, mass=1000.0, td=1.0, p=1.0, rho=1.0, Isp=2000.0, MU_=MU, DU_=DU, TU_=TU):
    return np.array([MU_, DU_, TU_, thrustLimit, mass, td, p, rho, Isp], dtype=np.float64)


def ep_rhs(state, control, td=1.0, dp=None):
    state = _f64(state); control = _f64(control); dp = dparams() if dp is None else dp
    out = np.zeros_like(state)
    lib().oracle_ep_rhs(C.c_int(state.size), _p(state), _p(dp), _p(control), C.c_double(td), _p(out))
    return out


def sc_rhs(s, ip):
    s = _f64(s); out = np.zeros_like(s)
    rc = lib().oracle_sc_rhs(C.c_int(s.size), _p(s), _p(ip), _p(out))
    if rc:
        raise ValueError("Invalid value of p!")
    return out


def sc_rhs_jac(s, ip):
    s = _f64(s); n = s.size; A = np.zeros((n, n))
    rc = lib().oracle_sc_rhs_jac(C.c_int(n), _p(s), _p(ip), _p(A))
    if rc:
        raise ValueError("Invalid value of p!")
    return A.T.copy()   # C side is column-major


def ode7_8_ep(x0, t0, t1, nsteps, control, td=1.0, dp=None):
    x0 = _f64(x0); control = _f64(control); dp = dparams() if dp is None else dp
    n = x0.size
    Xout = np.zeros((nsteps, n)); me = C.c_double(0)
    lib().oracle_ode7_8_ep(C.c_int(n), C.c_double(t0), C.c_double(t1), C.c_int(nsteps), _p(x0), _p(dp), _p(control),
                           C.c_double(td), _p(Xout), C.byref(me))
    return Xout.T.copy(), me.value


def ode78_ep(x0, t0, t1, tol, control, td=1.0, dp=None):
    x0 = _f64(x0); control = _f64(control); dp = dparams() if dp is None else dp
    xend = np.zeros_like(x0); na = C.c_int(0); nt = C.c_int(0)
    st = lib().oracle_ode78_ep(C.c_int(x0.size), C.c_double(t0), C.c_double(t1), C.c_double(tol), _p(x0), _p(dp),
                               _p(control), C.c_double(td), _p(xend), C.byref(na), C.byref(nt))
    return xend, st, na.value, nt.value


def _seg_args(Xa, Xb, ua, ub, ta, tb):
    Xa = _f64(Xa); Xb = _f64(Xb); ua = _f64(ua); ub = _f64(ub); ta = _f64(ta); tb = _f64(tb)
    n_seg, n = Xa.shape
    assert Xb.shape == (n_seg, n) and ua.shape == (n_seg, 3) and ub.shape == (n_seg, 3)
    return Xa, Xb, ua, ub, ta, tb, n_seg, n


def direct_defect(Xa, Xb, ua, ub, ta, tb, nsteps=10, dp=None, mode=0, tol=1e-13, nthreads=1):
    """Rows are segments (AoS).  Returns defect (n_seg,n), errors, status, steps(n_seg,2)."""
    Xa, Xb, ua, ub, ta, tb, n_seg, n = _seg_args(Xa, Xb, ua, ub, ta, tb)
    dp = dparams() if dp is None else dp
    defect = np.zeros((n_seg, n)); errors = np.zeros(n_seg)
    status = np.zeros(n_seg, dtype=np.int32); steps = np.zeros((n_seg, 2), dtype=np.int32)
    lib().oracle_direct_defect(C.c_longlong(n_seg), C.c_int(n), C.c_int(nsteps), _p(Xa), _p(Xb), _p(ua), _p(ub), _p(ta),
                               _p(tb), _p(dp), C.c_int(mode), C.c_double(tol), _p(defect), _p(errors), _p(status),
                               _p(steps), C.c_int(nthreads))
    return defect, errors, status, steps


def direct_jac_fd(Xa, Xb, ua, ub, ta, tb, defect_nom, nsteps=10, dp=None, mode=0, tol=1e-13, pert=1e-8, nthreads=1):
    """Reference-faithful forward-FD Jacobian.  Returns (n_seg, n, 2(n+3)) with [.., row, col]."""
    Xa, Xb, ua, ub, ta, tb, n_seg, n = _seg_args(Xa, Xb, ua, ub, ta, tb)
    dp = dparams() if dp is None else dp
    nvar = 2 * (n + 3)
    jac = np.zeros((n_seg, nvar, n)); defect_nom = _f64(defect_nom)
    lib().oracle_direct_jac_fd(C.c_longlong(n_seg), C.c_int(n), C.c_int(nsteps), _p(Xa), _p(Xb), _p(ua), _p(ub), _p(ta),
                               _p(tb), _p(dp), C.c_int(mode), C.c_double(tol), C.c_double(pert), _p(defect_nom), _p(jac),
                               C.c_int(nthreads))
    return jac.transpose(0, 2, 1).copy()


def direct_jac_var(Xa, Xb, ua, ub, ta, tb, nsteps=10, dp=None, mode=0, tol=1e-13, with_partials=False, nthreads=1):
    """Dual-number (variational) Jacobian of the same discrete map.  Returns defect, errors, jac, status."""
    Xa, Xb, ua, ub, ta, tb, n_seg, n = _seg_args(Xa, Xb, ua, ub, ta, tb)
    dp = dparams() if dp is None else dp
    nvar = 2 * (n + 3)
    defect = np.zeros((n_seg, n)); errors = np.zeros(n_seg); jac = np.zeros((n_seg, nvar, n))
    status = np.zeros(n_seg, dtype=np.int32)
    lib().oracle_direct_jac_var(C.c_longlong(n_seg), C.c_int(n), C.c_int(nsteps), _p(Xa), _p(Xb), _p(ua), _p(ub), _p(ta),
                                _p(tb), _p(dp), C.c_int(mode), C.c_double(tol), C.c_int(int(with_partials)), _p(defect),
                                _p(errors), _p(jac), _p(status), C.c_int(nthreads))
    return defect, errors, jac.transpose(0, 2, 1).copy(), status


def direct_variational_host(Xa, Xb, ua, ub, ta, tb, nsteps=10, dp=None, nthreads=1):
    """BASELINE ONLY: the GPU path's own arithmetic (state + variational equations, one integration per leg) on the host cores.
    Returns defect (n_seg, n), jac (n_seg, n, 2(n+3)) [.., row, col]."""
    Xa, Xb, ua, ub, ta, tb, n_seg, n = _seg_args(Xa, Xb, ua, ub, ta, tb)
    dp = dparams() if dp is None else dp
    nv = 2 * (n + 3)
    defect = np.zeros((n_seg, n)); jac = np.zeros((n_seg, nv, n))
    lib().oracle_direct_variational_host(C.c_longlong(n_seg), C.c_int(n), C.c_int(nsteps), _p(Xa), _p(Xb), _p(ua), _p(ub), _p(ta), _p(tb), _p(dp),
                                         _p(defect), _p(jac), C.c_int(nthreads))
    return defect, jac.transpose(0, 2, 1).copy()


def indirect_prop(x0, t0, t1, ip, thrustLimit=None, rho=None, atol=1e-13, rtol=1e-13, controller=0, nthreads=1, est=-1):
    x0 = _f64(x0); t0 = _f64(t0); t1 = _f64(t1)
    n_seg, ndim = x0.shape
    tl = None if thrustLimit is None else _f64(thrustLimit); rh = None if rho is None else _f64(rho)
    xend = np.zeros_like(x0); status = np.zeros(n_seg, dtype=np.int32)
    na = np.zeros(n_seg, dtype=np.int32); nt = np.zeros(n_seg, dtype=np.int32)
    lib().oracle_indirect_prop(C.c_longlong(n_seg), C.c_int(ndim), _p(x0), _p(t0), _p(t1), _p(ip), _p(tl), _p(rh),
                               C.c_double(atol), C.c_double(rtol), C.c_int(controller), _p(xend), _p(status), _p(na),
                               _p(nt), C.c_int(nthreads), C.c_int(est))
    return xend, status, na, nt


def indirect_prop_ld(x0, t0, t1, ip, thrustLimit=None, rho=None, atol=1e-17, rtol=1e-17, nthreads=1):
    """End states in 80-bit long double at a tolerance below the double runs' (truth for error budgets); rounded to double."""
    x0 = _f64(x0); t0 = _f64(t0); t1 = _f64(t1)
    n_seg, ndim = x0.shape
    tl = None if thrustLimit is None else _f64(thrustLimit); rh = None if rho is None else _f64(rho)
    xend = np.zeros_like(x0); status = np.zeros(n_seg, dtype=np.int32)
    lib().oracle_indirect_prop_ld(C.c_longlong(n_seg), C.c_int(ndim), _p(x0), _p(t0), _p(t1), _p(ip), _p(tl), _p(rh),
                                  C.c_double(atol), C.c_double(rtol), _p(xend), _p(status), C.c_int(nthreads))
    return xend, status


def indirect_prop_jac(x0, t0, t1, ip, thrustLimit=None, rho=None, atol=1e-13, rtol=1e-13, controller=0, nthreads=1):
    """Returns xend (n_seg,ndim), phi (n_seg,ndim,ndim) with phi[s,r,c] = d xend_r / d x0_c, status, nacc, natt."""
    x0 = _f64(x0); t0 = _f64(t0); t1 = _f64(t1)
    n_seg, ndim = x0.shape
    tl = None if thrustLimit is None else _f64(thrustLimit); rh = None if rho is None else _f64(rho)
    xend = np.zeros_like(x0); phi = np.zeros((n_seg, ndim, ndim)); status = np.zeros(n_seg, dtype=np.int32)
    na = np.zeros(n_seg, dtype=np.int32); nt = np.zeros(n_seg, dtype=np.int32)
    lib().oracle_indirect_prop_jac(C.c_longlong(n_seg), C.c_int(ndim), _p(x0), _p(t0), _p(t1), _p(ip), _p(tl), _p(rh),
                                   C.c_double(atol), C.c_double(rtol), C.c_int(controller), _p(xend), _p(phi), _p(status),
                                   _p(na), _p(nt), C.c_int(nthreads))
    return xend, phi.transpose(0, 2, 1).copy(), status, na, nt
