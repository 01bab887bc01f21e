#!/usr/bin/env python
"""Generates tests/golden/golden14_v1.json -- an independent pin for the 14-dim system
[r v m | lr lv lm] (state + costate + mass, 14 x 14 STM) that BASELINE's north_star asks for.

The reference has no CRTBP right-hand side with mass and mass costate (its only 14-dim RHS
is the two-body one, GeneralCode/twoBody_stateCostate_mass_deriv.jl:11-78), so there is nothing of the
reference's to restate.  Instead of a second hand-written copy of our own equations, this script DERIVES
them: it writes down the Hamiltonian of the controlled CRTBP with a thrust FORCE tau (N) held fixed,

    H = lr . v + lv . (grad Omega(r) + 2 td (vy, -vx, 0) + k tau / m) + lm . (-c |tau|),
    Omega = (x^2 + y^2)/2 + (1 - MU)/r1 + MU/r2,   k = TU^2/(DU 1e3)  (CRTBP_stateCostate_deriv.jl:33),
    c = TU/(Isp g0), g0 = 9.81                      (CRTBP_prop_EP_deriv.jl:41-42),

lets sympy form  x' = dH/dl  and  l' = -dH/dx  (Pontryagin), substitutes the reference's control law
(CRTBP_stateCostate_deriv.jl:36-64: direction -lv/|lv|, magnitude from p / rho with the acceleration limit
thrustLimit k / m(t)), integrates with scipy's DOP853 at rtol = atol = 3e-14, and differentiates the flow
by Richardson-extrapolated central differences.  oracle/ (dual numbers through RKF7(8)) and the CUDA kernel
K3-14 are both checked against this file.  Run from the repo root:  python tests/golden/make_golden14.py
"""
import json
import os
import sys

import numpy as np
import sympy as sp
from scipy.integrate import solve_ivp

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
from lowthrustopt_b200 import synthetic as S  # noqa: E402  (input generation only)

MU = 0.012150585609624037
DU = 384747.96285603708
TU = 375699.81732246041
G0 = 9.81
ISP = 2000.0


def derive():
    """Returns f(s14, tau3, td) -> d s14 / dt, lambdified from the symbolic Euler-Lagrange equations."""
    x, y, z, vx, vy, vz, m = sp.symbols("x y z vx vy vz m", real=True)
    l1, l2, l3, l4, l5, l6, lm = sp.symbols("l1 l2 l3 l4 l5 l6 lm", real=True)
    t1, t2, t3, td = sp.symbols("tau1 tau2 tau3 td", real=True)
    k = sp.Float(TU) ** 2 / (sp.Float(DU) * 1000)
    c = sp.Float(TU) / (sp.Float(ISP) * sp.Float(G0))
    mu = sp.Float(MU)
    r1 = sp.sqrt((x + mu) ** 2 + y ** 2 + z ** 2)
    r2 = sp.sqrt((x + mu - 1) ** 2 + y ** 2 + z ** 2)
    Om = (x ** 2 + y ** 2) / 2 + (1 - mu) / r1 + mu / r2
    tnorm = sp.sqrt(t1 ** 2 + t2 ** 2 + t3 ** 2)
    acc = [sp.diff(Om, x) + 2 * td * vy + k * t1 / m,
           sp.diff(Om, y) - 2 * td * vx + k * t2 / m,
           sp.diff(Om, z) + k * t3 / m]
    mdot = -c * tnorm
    state = [x, y, z, vx, vy, vz, m]
    cost = [l1, l2, l3, l4, l5, l6, lm]
    H = l1 * vx + l2 * vy + l3 * vz + l4 * acc[0] + l5 * acc[1] + l6 * acc[2] + lm * mdot
    rhs = [sp.diff(H, q) for q in cost] + [-sp.diff(H, q) for q in state]     # tau held fixed (Pontryagin)
    return sp.lambdify([state + cost, [t1, t2, t3], td], rhs, modules="numpy", cse=True)


F = derive()


def thrust(s, law):
    """The control law of CRTBP_stateCostate_deriv.jl:36-64 as a thrust FORCE in N, with the acceleration limit of :33 taken at m(t)."""
    thrustLimit, td, p, rho = law
    m = s[6]; lv = s[10:13]
    k = TU ** 2 / (DU * 1e3)
    aL = thrustLimit * k / m
    n = np.linalg.norm(lv)
    if p == 0:
        umag = aL
    elif p == 1:
        umag = 0.5 * (1 + np.tanh((n - 1) / (2 * rho))) * aL
    elif p > 1:
        umag = min((n / p) ** (1 / (p - 1)), aL)
    else:
        raise ValueError("Invalid value of p!")
    if n == 0:
        return np.zeros(3)
    return -(umag * m / k) * lv / n


def rhs14(s, law):
    tau = thrust(s, law)
    if not np.any(tau):
        tau = np.array([1e-300, 0.0, 0.0])          # |tau| -> 0 limit without 0/0 in d|tau| terms (none survive: tau is held fixed)
    return np.array(F(list(s), list(tau), law[1]), dtype=np.float64)


def prop14(s0, t0, t1, law, tol=3e-14):
    sol = solve_ivp(lambda t, yy: rhs14(yy, law), (t0, t1), s0, method="DOP853", rtol=tol, atol=tol)
    return sol.y[:, -1]


def phi_richardson(s0, t0, t1, law):
    n = len(s0); P = np.zeros((n, n))

    def D(j, h):
        a = s0.copy(); b = s0.copy(); a[j] += h; b[j] -= h
        return (prop14(a, t0, t1, law) - prop14(b, t0, t1, law)) / (2 * h)
    for j in range(n):
        h = 2e-4 * max(1.0, abs(s0[j]))              # the mass component is ~1e3
        P[:, j] = (4 * D(j, h / 2) - D(j, h)) / 3
    return P


def main():
    laws = [(0.05, 1.0, 1.0, 1.0), (10.0, 1.0, 2.0, 1.0), (0.05, 1.0, 1.0, 1e-2), (0.05, 1.0, 0.0, 1.0), (10.0, -1.0, 2.0, 1.0)]
    b = S.indirect_batch(2 * len(laws), ndim=14, seed=31)
    b["x0"][:, 10:13] *= 8.0                        # |lv| near 1 so the p = 1 switch is exercised
    cases = []
    for i in range(2 * len(laws)):
        law = laws[i % len(laws)]
        s0 = b["x0"][i]; t1 = float(b["t1"][i])
        rec = dict(ndim=14, x0=s0.tolist(), t0=0.0, t1=t1, thrustLimit=law[0], td=law[1], p=law[2], rho=law[3], Isp=ISP,
                   xend=prop14(s0, 0.0, t1, law).tolist())
        if i < len(laws):
            rec["phi_richardson"] = phi_richardson(s0, 0.0, t1, law).tolist()
        cases.append(rec)
    out = {"about": "14-dim system derived symbolically from its Hamiltonian (sympy) + scipy DOP853; see make_golden14.py",
           "MU": MU, "DU": DU, "TU": TU, "indirect14": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden14_v1.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
