#!/usr/bin/env python
"""Generates tests/golden/golden_v1.json.

The reference is Julia and cannot run in this image, and it ships no expected outputs.
These vectors therefore come from a SECOND, independent restatement written here in
numpy/scipy (not from oracle/): the direct RHS and ode7_8 follow the reference line by
line with numpy's matvec standing in for Julia's `f*beta_[:,j]`; the indirect end states
come from scipy's DOP853 at rtol = atol = 1e-13 on a term-by-term numpy copy of
CRTBP_stateCostate_deriv!.  The C++ oracle and the CUDA kernels are both checked against
this file.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
from scipy.integrate import solve_ivp

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
from lowthrustopt_b200 import synthetic as S  # noqa: E402  (input generation only)

MU = 0.012150585609624037
DU = 384747.96285603708
TU = 375699.81732246041


def ep_rhs(state, Isp, control, td):                      # src/CRTBP_prop_EP_deriv.jl:8-61
    x, y, z, xd, yd, zd = state[:6]
    m = state[6] if len(state) == 7 else 1000.0
    r1 = np.sqrt((x + MU) ** 2 + y ** 2 + z ** 2)
    r2 = np.sqrt((x + MU - 1) ** 2 + y ** 2 + z ** 2)
    r1_3 = r1 ** 3; r2_3 = r2 ** 3
    nu = np.linalg.norm(control)
    T_mag = nu / m / 1e3 * TU ** 2 / DU
    T = control if nu == 0 else control / nu * T_mag
    mdot = -td * nu / (Isp * 9.81) * TU
    om = td
    xdd = -(1 - MU) * (x + MU) / r1_3 - MU * (x - 1 + MU) / r2_3 + 2 * om * yd + x + T[0]
    ydd = -(1 - MU) * y / r1_3 - MU * y / r2_3 - 2 * om * xd + y + T[1]
    zdd = -(1 - MU) * z / r1_3 - MU * z / r2_3 + T[2]
    out = [xd, yd, zd, xdd, ydd, zdd]
    if len(state) == 7:
        out.append(mdot)
    return np.array(out)


ALPHA = np.array([2 / 27, 1 / 9, 1 / 6, 5 / 12, 0.5, 5 / 6, 1 / 6, 2 / 3, 1 / 3, 1, 0, 1])
BETA = np.zeros((13, 12))                                  # GeneralCode/ode.jl:877-889
BETA[:, 0] = [2 / 27, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
BETA[:, 1] = [1 / 36, 1 / 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
BETA[:, 2] = [1 / 24, 0, 1 / 8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
BETA[:, 3] = [5 / 12, 0, -25 / 16, 25 / 16, 0, 0, 0, 0, 0, 0, 0, 0, 0]
BETA[:, 4] = [0.05, 0, 0, 0.25, 0.2, 0, 0, 0, 0, 0, 0, 0, 0]
BETA[:, 5] = [-25 / 108, 0, 0, 125 / 108, -65 / 27, 125 / 54, 0, 0, 0, 0, 0, 0, 0]
BETA[:, 6] = [31 / 300, 0, 0, 0, 61 / 225, -2 / 9, 13 / 900, 0, 0, 0, 0, 0, 0]
BETA[:, 7] = [2, 0, 0, -53 / 6, 704 / 45, -107 / 9, 67 / 90, 3, 0, 0, 0, 0, 0]
BETA[:, 8] = [-91 / 108, 0, 0, 23 / 108, -976 / 135, 311 / 54, -19 / 60, 17 / 6, -1 / 12, 0, 0, 0, 0]
BETA[:, 9] = [2383 / 4100, 0, 0, -341 / 164, 4496 / 1025, -301 / 82, 2133 / 4100, 45 / 82, 45 / 164, 18 / 41, 0, 0, 0]
BETA[:, 10] = [3 / 205, 0, 0, 0, 0, -6 / 41, -3 / 205, -3 / 41, 3 / 41, 6 / 41, 0, 0, 0]
BETA[:, 11] = [-1777 / 4100, 0, 0, -341 / 164, 4496 / 1025, -289 / 82, 2193 / 4100, 51 / 82, 33 / 164, 12 / 41, 0, 1, 0]
CHI = np.array([0, 0, 0, 0, 0, 34 / 105, 9 / 35, 9 / 35, 9 / 280, 9 / 280, 0, 41 / 840, 41 / 840])
PSI = np.array([1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, -1])


def ode7_8(fun, tspan, x0):                                # GeneralCode/ode.jl:773-953
    h = np.diff(tspan)
    neq = len(x0); N = len(tspan)
    Xout = np.zeros((neq, N)); f = np.zeros((neq, 13))
    Xout[:, 0] = x0
    maxErr = 0.0
    for ind in range(1, N):
        hi = h[ind - 1]; xi = Xout[:, ind - 1]
        f[:, 0] = fun(xi)
        for j in range(12):
            f[:, j + 1] = fun(xi + hi * f @ BETA[:, j])
        Xout[:, ind] = xi + hi * f @ CHI
        gamma1 = hi * 41 / 840 * f @ PSI
        delta = np.linalg.norm(gamma1, np.inf)
        if delta > maxErr:
            maxErr = delta
    return Xout, maxErr


def linrange(a, b, n):                                     # Julia LinRange / Base.lerpi
    t = np.arange(n) / (n - 1)
    return (1 - t) * a + t * b


def defect_pair(Xa, Xb, ua, ub, ta, tb, nsteps, Isp):      # src/multiShoot_CRTBP_direct.jl:77-105
    tmid = ta + (tb - ta) / 2
    tspan = linrange(ta, tmid, nsteps)
    sf, ef = ode7_8(lambda s: ep_rhs(s, Isp, ua, 1.0), tspan, Xa.copy())
    x0 = Xb.copy(); x0[3:6] = -x0[3:6]
    sb, eb = ode7_8(lambda s: ep_rhs(s, Isp, ub, -1.0), tspan, x0)
    e = sb[:, -1].copy(); e[3:6] = -e[3:6]
    return sf[:, -1] - e, max(ef, eb)


def jac_fd(Xa, Xb, ua, ub, ta, tb, nsteps, Isp, d0, pert=1e-8):   # :111-143
    n = len(Xa); nvar = 2 * (n + 3)
    XU = np.concatenate([Xa, Xb, ua, ub]); J = np.zeros((n, nvar))
    for j in range(nvar):
        m = XU.copy(); m[j] += pert
        d, _ = defect_pair(m[:n], m[n:2 * n], m[2 * n:2 * n + 3], m[2 * n + 3:], ta, tb, nsteps, Isp)
        J[:, j] = (d - d0) / pert
    return J


def jac_richardson(Xa, Xb, ua, ub, ta, tb, nsteps, Isp):
    """Central differences at two step sizes, Richardson-extrapolated: the FD-noise-free check."""
    n = len(Xa); nvar = 2 * (n + 3)
    XU = np.concatenate([Xa, Xb, ua, ub]); J = np.zeros((n, nvar))

    def D(j, h):
        a = XU.copy(); b = XU.copy(); a[j] += h; b[j] -= h
        da, _ = defect_pair(a[:n], a[n:2 * n], a[2 * n:2 * n + 3], a[2 * n + 3:], ta, tb, nsteps, Isp)
        db, _ = defect_pair(b[:n], b[n:2 * n], b[2 * n:2 * n + 3], b[2 * n + 3:], ta, tb, nsteps, Isp)
        return (da - db) / (2 * h)
    for j in range(nvar):
        h = 1e-4 * max(1.0, abs(XU[j]))
        J[:, j] = (4 * D(j, h / 2) - D(j, h)) / 3
    return J


def sc_rhs12(s, thrustLimit, mass, td, p, rho):            # src/CRTBP_stateCostate_deriv.jl:9-90
    X1, X2, X3, X4, X5, X6, L1, L2, L3, L4, L5, L6 = s
    lv = s[9:12]
    aL = thrustLimit / mass / 1e3 * TU ** 2 / DU
    if p == 0:
        umag = aL
    elif p == 1:
        g = np.linalg.norm(lv) - 1
        umag = 1 / 2 * (1 + np.tanh(g / (2 * rho))) * aL
    elif p > 1:
        umag = (1 / p * np.linalg.norm(lv)) ** (1 / (p - 1))
        umag = min(umag, aL)
    else:
        raise ValueError("Invalid value of p!")
    with np.errstate(invalid="ignore", divide="ignore"):
        ca = -umag * lv / np.linalg.norm(lv)
    if np.isnan(ca[0]):
        ca = np.zeros(3)
    r1_3 = ((X1 + MU) ** 2 + X2 ** 2 + X3 ** 2) ** 1.5
    r2_3 = ((X1 + MU - 1) ** 2 + X2 ** 2 + X3 ** 2) ** 1.5
    t1 = ((MU + X1 - 1) ** 2 + X2 ** 2 + X3 ** 2)
    t2 = ((MU + X1) ** 2 + X2 ** 2 + X3 ** 2)
    t3 = (2 * MU + 2 * X1 - 2)
    d = np.zeros(12)
    d[0:3] = s[3:6]
    d[3] = -(1 - MU) * (X1 + MU) / r1_3 - MU * (X1 - 1 + MU) / r2_3 + 2 * td * X5 + X1 + ca[0]
    d[4] = -(1 - MU) * X2 / r1_3 - MU * X2 / r2_3 - 2 * td * X4 + X2 + ca[1]
    d[5] = -(1 - MU) * X3 / r1_3 - MU * X3 / r2_3 + ca[2]
    d[6] = (- L5 * ((3 * MU * X2 * t3) / (2 * t1 ** 2.5) - (3 * X2 * (MU - 1) * (2 * MU + 2 * X1)) / (2 * t2 ** 2.5))
            - L6 * ((3 * MU * X3 * t3) / (2 * t1 ** 2.5) - (3 * X3 * (MU - 1) * (2 * MU + 2 * X1)) / (2 * t2 ** 2.5))
            - L4 * ((MU - 1) / t2 ** 1.5 - MU / t1 ** 1.5 + (3 * MU * (MU + X1 - 1) * t3) / (2 * t1 ** 2.5)
                    - (3 * (MU + X1) * (MU - 1) * (2 * MU + 2 * X1)) / (2 * t2 ** 2.5) + 1))
    d[7] = (L6 * ((3 * X2 * X3 * (MU - 1)) / t2 ** 2.5 - (3 * MU * X2 * X3) / t1 ** 2.5)
            - L5 * ((MU - 1) / t2 ** 1.5 - MU / t1 ** 1.5 - (3 * X2 ** 2 * (MU - 1)) / t2 ** 2.5 + (3 * MU * X2 ** 2) / t1 ** 2.5 + 1)
            - L4 * ((3 * MU * X2 * (MU + X1 - 1)) / t1 ** 2.5 - (3 * X2 * (MU + X1) * (MU - 1)) / t2 ** 2.5))
    d[8] = (L6 * (MU / t1 ** 1.5 - (MU - 1) / t2 ** 1.5 + (3 * X3 ** 2 * (MU - 1)) / t2 ** 2.5 - (3 * MU * X3 ** 2) / t1 ** 2.5)
            + L5 * ((3 * X2 * X3 * (MU - 1)) / t2 ** 2.5 - (3 * MU * X2 * X3) / t1 ** 2.5)
            - L4 * ((3 * MU * X3 * (MU + X1 - 1)) / t1 ** 2.5 - (3 * X3 * (MU + X1) * (MU - 1)) / t2 ** 2.5))
    d[9] = 2 * L5 * td - L1
    d[10] = -L2 - 2 * L4 * td
    d[11] = -L3
    return d


def prop12(s0, t0, t1, law, tol=1e-13):
    sol = solve_ivp(lambda t, y: sc_rhs12(y, *law), (t0, t1), s0, method="DOP853", rtol=tol, atol=tol)
    return sol.y[:, -1]


def phi_richardson(s0, t0, t1, law):
    n = len(s0); P = np.zeros((n, n))

    def D(j, h):
        a = s0.copy(); b = s0.copy(); a[j] += h; b[j] -= h
        return (prop12(a, t0, t1, law, 1e-14) - prop12(b, t0, t1, law, 1e-14)) / (2 * h)
    for j in range(n):
        h = 2e-4
        P[:, j] = (4 * D(j, h / 2) - D(j, h)) / 3
    return P


def main():
    out = {"about": "independent numpy/scipy restatement; see make_golden.py", "MU": MU, "DU": DU, "TU": TU}
    Isp = 2000.0; nsteps = 10
    direct = []
    for nstate, seed in ((6, 11), (7, 12)):
        b = S.direct_batch(6, nstate=nstate, seed=seed)
        if nstate == 6:   # the demo's first iteration has u = 0 (CRTBP_Multishoot_direct_demo.jl:179)
            b["ua"][0] = 0.0; b["ub"][0] = 0.0
        for s in range(6):
            a = {k: b[k][s] for k in b}
            d, e = defect_pair(a["Xa"], a["Xb"], a["ua"], a["ub"], a["ta"], a["tb"], nsteps, Isp)
            Jfd = jac_fd(a["Xa"], a["Xb"], a["ua"], a["ub"], a["ta"], a["tb"], nsteps, Isp, d)
            Jr = jac_richardson(a["Xa"], a["Xb"], a["ua"], a["ub"], a["ta"], a["tb"], nsteps, Isp)
            direct.append(dict(nstate=nstate, nsteps=nsteps, Isp=Isp, Xa=a["Xa"].tolist(), Xb=a["Xb"].tolist(),
                               ua=a["ua"].tolist(), ub=a["ub"].tolist(), ta=float(a["ta"]), tb=float(a["tb"]),
                               defect=d.tolist(), err=float(e), jac_fd=Jfd.tolist(), jac_richardson=Jr.tolist()))
    out["direct"] = direct
    indirect = []
    laws = [(0.05, 1000.0, 1.0, 1.0, 1.0), (10.0, 1000.0, 1.0, 2.0, 1.0), (0.05, 1000.0, 1.0, 1.0, 1e-2),
            (0.05, 1000.0, 1.0, 0.0, 1.0), (0.0, 1000.0, 1.0, 1.0, 1.0)]
    b = S.indirect_batch(len(laws) * 2, ndim=12, seed=21)
    b["x0"][:, 9:12] *= 8.0       # |lv| near 1 so the p = 1 switch is exercised
    for i in range(len(laws) * 2):
        law = laws[i % len(laws)]
        s0 = b["x0"][i]
        xe = prop12(s0, 0.0, float(b["t1"][i]), law)
        rec = dict(ndim=12, x0=s0.tolist(), t0=0.0, t1=float(b["t1"][i]), thrustLimit=law[0], mass=law[1], td=law[2], p=law[3],
                   rho=law[4], xend=xe.tolist())
        if i < len(laws):
            rec["phi_richardson"] = phi_richardson(s0, 0.0, float(b["t1"][i]), law).tolist()
        indirect.append(rec)
    out["indirect"] = indirect
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
