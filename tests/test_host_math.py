"""CPU: the kernels' __host__ __device__ arithmetic (lto_math.cuh, lto_prop_generic.cuh), built
for the host as a TEST-ONLY library, against the oracle's dual numbers."""
import ctypes as C

import numpy as np
import pytest

from conftest import ptr
from lowthrustopt_b200 import synthetic as S

LAWS = [(1.0, 1.0, 0.05), (1.0, 1e-2, 0.05), (2.0, 1.0, 10.0), (2.0, 1.0, 1e-4), (0.0, 1.0, 0.05), (1.5, 1.0, 10.0)]


@pytest.mark.parametrize("nd", [12, 14])
@pytest.mark.parametrize("law", LAWS)
def test_variational_matrix_matches_dual_numbers(nd, law, oracle, hostcheck):
    p, rho, tl = law
    rng = np.random.default_rng(5)
    X = S.load_orbit(1)
    for trial in range(4):
        s = np.concatenate([X[:, 7 * trial + 3], [900.0] if nd == 14 else [], rng.standard_normal(6) * (0.2 + trial), [0.01] if nd == 14 else []])
        ip = oracle.iparams(tl, p=p, rho=rho)
        f = np.zeros(nd); A = np.zeros((nd, nd))
        assert hostcheck.hc_sc_rhs_jac(nd, ptr(s), ptr(ip), ptr(f), ptr(A)) == 0
        fo = oracle.sc_rhs(s, ip); Ao = oracle.sc_rhs_jac(s, ip)
        assert np.abs(f - fo).max() < 1e-13 * max(1.0, np.abs(fo).max())
        assert np.abs(A.T - Ao).max() < 1e-13 * max(1.0, np.abs(Ao).max())


@pytest.mark.parametrize("nd", [12, 14])
def test_zero_costate_guard(nd, oracle, hostcheck):
    """|lv| = 0 -> control = 0 with zero sensitivity (CRTBP_stateCostate_deriv.jl:59-64)."""
    X = S.load_orbit(2)
    s = np.concatenate([X[:, 3], [1000.0] if nd == 14 else [], [0.1, 0.2, 0.3, 0, 0, 0], [0.0] if nd == 14 else []])
    ip = oracle.iparams(0.05, p=1.0, rho=1.0)
    f = np.zeros(nd); A = np.zeros((nd, nd))
    assert hostcheck.hc_sc_rhs_jac(nd, ptr(s), ptr(ip), ptr(f), ptr(A)) == 0
    assert np.abs(f - oracle.sc_rhs(s, ip)).max() < 1e-14 and np.all(np.isfinite(A))


@pytest.mark.parametrize("ns", [6, 7])
@pytest.mark.parametrize("mode", [0, 1])
def test_direct_leg_matches_oracle(ns, mode, oracle, hostcheck):
    b = S.direct_batch(5, nstate=ns, seed=3)
    dp = oracle.dparams()
    do, eo, Jo, st = oracle.direct_jac_var(b["Xa"], b["Xb"], b["ua"], b["ub"], b["ta"], b["tb"], mode=mode, tol=1e-12)
    for s in range(5):
        ta, tb = b["ta"][s], b["tb"][s]; tm = ta + (tb - ta) / 2
        xf = np.zeros(ns); xb = np.zeros(ns); Sf = np.zeros((ns + 3, ns)); Sb = np.zeros((ns + 3, ns)); me = C.c_double(); na = C.c_int()
        errs = []
        for back, X, U, xo, So in ((0, b["Xa"][s], b["ua"][s], xf, Sf), (1, b["Xb"][s], b["ub"][s], xb, Sb)):
            hostcheck.hc_ep_leg(ns, 1, ptr(X), ptr(U), back, C.c_double(ta), C.c_double(tm), mode, 10, C.c_double(1e-12), 0, ptr(dp),
                                ptr(xo), ptr(So), C.byref(me), C.byref(na))
            errs.append(me.value)
        J = np.hstack([Sf.T[:, :ns], -Sb.T[:, :ns], Sf.T[:, ns:], -Sb.T[:, ns:]])
        tol = 1e-14 if mode == 0 else 1e-11
        scale = np.maximum(1.0, np.abs(b["Xa"][s]))
        assert np.all(np.abs((xf - xb) - do[s]) / scale < tol)
        assert np.abs(J - Jo[s]).max() < (1e-13 if mode == 0 else 1e-10)
        if mode == 0:
            # maxErr = 41/840*h*|k1+k11-k12-k13| is itself rounding noise (~1e-18) at the demo step size
            assert abs(max(errs) - eo[s]) < 2e-17 * scale.max()


@pytest.mark.parametrize("nd", [12, 14])
@pytest.mark.parametrize("ctl", [0, 1])
def test_indirect_segment_matches_oracle(nd, ctl, oracle, hostcheck):
    b = S.indirect_batch(4, ndim=nd, seed=9)
    b["x0"][:, (9 if nd == 12 else 10):(12 if nd == 12 else 13)] *= 8.0
    # smooth switch, unclamped p = 2, sharp switch, constant thrust (p = 0), clamped p = 2 (umag' = 0, CRTBP_stateCostate_deriv.jl:48-50)
    for law in ((1.0, 1.0, 0.05), (2.0, 1.0, 10.0), (1.0, 1e-2, 0.05), (0.0, 1.0, 0.05), (2.0, 1.0, 1e-3)):
        ip = oracle.iparams(law[2], p=law[0], rho=law[1])
        xo, Po, so, nao, nto = oracle.indirect_prop_jac(b["x0"], b["t0"], b["t1"], ip, controller=ctl)
        for s in range(4):
            xe = np.zeros(nd); Phi = np.zeros((nd, nd)); na = C.c_int(); nt = C.c_int()
            st = hostcheck.hc_sc_seg(nd, 1, ptr(b["x0"][s]), C.c_double(0.0), C.c_double(b["t1"][s]), C.c_double(1e-13), C.c_double(1e-13),
                                     ctl, 1, ptr(ip), ptr(xe), ptr(Phi), C.byref(na), C.byref(nt))
            assert st == 0 and so[s] == 0
            sc = np.maximum(1.0, np.abs(xo[s]))
            assert np.all(np.abs(xe - xo[s]) / sc < 1e-11)
            assert np.abs(Phi.T - Po[s]).max() < 1e-10 * max(1.0, np.abs(Po[s]).max())
            assert abs(na.value - nao[s]) <= 1


def test_kernel_arithmetic_14_matches_symbolic_golden(oracle, hostcheck, golden14):
    """The kernels' own 14-dim arithmetic (lto_math.cuh sc_stage<14> / sc_col<14> through the generic driver, compiled for the
    host) against the system derived from the Hamiltonian by sympy (tests/golden/make_golden14.py): every law of the golden
    set, time_direction = -1 included."""
    for g in golden14["indirect14"]:
        ip = oracle.iparams(g["thrustLimit"], td=g["td"], p=g["p"], rho=g["rho"], Isp=g["Isp"])
        x0 = np.array(g["x0"]); want = np.array(g["xend"])
        xe = np.zeros(14); Phi = np.zeros((14, 14)); na = C.c_int(); nt = C.c_int()
        st = hostcheck.hc_sc_seg(14, 1, ptr(x0), C.c_double(g["t0"]), C.c_double(g["t1"]), C.c_double(1e-13), C.c_double(1e-13), 0, 1,
                                 ptr(ip), ptr(xe), ptr(Phi), C.byref(na), C.byref(nt))
        assert st == 0 and (np.abs(xe - want) / np.maximum(1.0, np.abs(want))).max() < 1e-11
        if "phi_richardson" in g:
            P = np.array(g["phi_richardson"])
            assert np.abs(Phi.T - P).max() < 5e-8 * max(1.0, np.abs(P).max() / 10)


@pytest.mark.parametrize("law", [dict(p=1.0, rho=1.0, thrustLimit=0.05), dict(p=2.0, rho=1.0, thrustLimit=10.0), dict(p=1.0, rho=1e-2, thrustLimit=0.05),
                                 dict(p=0.0, rho=1.0, thrustLimit=0.05), dict(p=2.0, rho=1.0, thrustLimit=1e-3)])
@pytest.mark.parametrize("td", [1.0, -1.0])
def test_half_column_formulation_matches_dual_numbers(law, td, oracle, hostcheck):
    """lto_hc_math.cuh -- the arithmetic of the half-column kernel K3 (second-order variables (r, v, lv, lv'), every STM column as a
    lane pair in Nystrom form) replayed thread by thread on the host -- against the oracle's dual numbers through RKF7(8):
    end states 1e-11, STM 1e-10 relative, about the same number of accepted steps (the initial step is taken over the state alone)."""
    from lowthrustopt_b200 import synthetic as S
    b = S.indirect_batch(24, ndim=12, seed=31)
    b["x0"][::3, 9:12] *= 8.0
    ip = oracle.iparams(law["thrustLimit"], td=td, p=law["p"], rho=law["rho"])
    xo, Po, so, nao, nto = oracle.indirect_prop_jac(b["x0"], b["t0"], b["t1"], ip)
    xs, ss, nas, nts = oracle.indirect_prop(b["x0"], b["t0"], b["t1"], ip)
    hostcheck.hc_halfcol_seg12.restype = C.c_int
    for s in range(24):
        for joint, xr, nar in ((1, xo[s], nao[s]), (0, xs[s], nas[s])):
            xe = np.zeros(12); Phi = np.zeros((12, 12)); na = C.c_int(); nt = C.c_int()
            st = hostcheck.hc_halfcol_seg12(ptr(b["x0"][s]), C.c_double(b["t0"][s]), C.c_double(b["t1"][s]), C.c_double(1e-13), C.c_double(1e-13),
                                            joint, ptr(ip), ptr(xe), ptr(Phi), C.byref(na), C.byref(nt))
            assert st == 0
            assert (np.abs(xe - xr) / np.maximum(1.0, np.abs(xr))).max() < 1e-11, (s, joint)
            assert abs(na.value - nar) <= max(2, nar // 8)
            assert np.abs(Phi.T - Po[s]).max() < (1e-10 if joint else 1e-8) * max(1.0, np.abs(Po[s]).max()), (s, joint)
