"""Seeded synthetic segment batches for the benchmark configurations.

Inputs only -- nothing here propagates anything.  The halo-orbit lookups restate
interpInitialStates (src/HelperFunctions.jl:18-35): an interpolating natural cubic
spline per row on the uniform grid LinRange(0, 1, 100), tau wrapped to [0, 1].
"""
import os

import numpy as np
from scipy.interpolate import CubicSpline

from .capi import MU, DU, TU, DAY

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
PERIODS = (2.9077, 3.1290)          # TU, L2_Anderson_1 / _2
DT_DEMO = 20.0 * DAY / TU / 29.0    # 0.158601 TU (CRTBP_Multishoot_direct_demo.jl:117-121)


def load_orbit(which):
    """L2_Anderson_{1,2}.txt as a (6, 100) array (CRTBP_Multishoot_direct_demo.jl:68-71)."""
    return np.loadtxt(os.path.join(_DATA, "L2_Anderson_%d.txt" % which))


_splines = {}


def orbit_spline(which):
    if which not in _splines:
        X = load_orbit(which)
        _splines[which] = CubicSpline(np.linspace(0.0, 1.0, X.shape[1]), X.T, bc_type="natural")
    return _splines[which]


def interp_initial_states(tau, which):
    """interpInitialStates (HelperFunctions.jl:18-35), vectorised over tau.  Returns (len(tau), 6)."""
    tau = np.mod(np.asarray(tau, dtype=np.float64), 1.0)
    return orbit_spline(which)(tau)


def direct_batch(n_seg, nstate=7, seed=20180001, zero_control=False):
    """Config 3: independent direct-method segments (pairs form)."""
    rng = np.random.default_rng(seed)
    which = np.where(np.arange(n_seg) % 2 == 0, 1, 2)      # odd (1-based) index -> orbit 1
    tau = rng.uniform(0.0, 1.0, n_seg)
    dt = rng.uniform(0.10, 0.20, n_seg)
    Xa = np.empty((n_seg, nstate)); Xb = np.empty((n_seg, nstate))
    for w in (1, 2):
        m = which == w
        Xa[m, :6] = interp_initial_states(tau[m], w)
        Xb[m, :6] = interp_initial_states(tau[m] + dt[m] / PERIODS[w - 1], w)
    Xa[:, :6] += 1e-3 * rng.standard_normal((n_seg, 6))
    Xb[:, :6] += 1e-3 * rng.standard_normal((n_seg, 6))
    if nstate == 7:
        Xa[:, 6] = rng.uniform(800.0, 1000.0, n_seg)
        Xb[:, 6] = Xa[:, 6] - rng.uniform(0.0, 0.1, n_seg)
    ua = 0.05 * rng.standard_normal((n_seg, 3)); ub = 0.05 * rng.standard_normal((n_seg, 3))
    if zero_control:
        ua[:] = 0.0; ub[:] = 0.0
    ta = rng.uniform(0.0, 4.0, n_seg)
    tb = ta + dt
    return dict(Xa=Xa, Xb=Xb, ua=ua, ub=ub, ta=ta, tb=tb)


def indirect_batch(n_seg, ndim=12, seed=20180002, dt=DT_DEMO):
    """Config 4: perturbed indirect-shooting initial guesses (pairs form)."""
    rng = np.random.default_rng(seed)
    which = np.where(np.arange(n_seg) % 2 == 0, 1, 2)
    tau = rng.uniform(0.0, 1.0, n_seg)
    x0 = np.empty((n_seg, ndim))
    for w in (1, 2):
        m = which == w
        x0[m, :6] = interp_initial_states(tau[m], w)
    lam = 0.1 * rng.standard_normal((n_seg, 6))            # CRTBP_Multishoot_indirect_demo.jl:166
    if ndim == 12:
        x0[:, 6:12] = lam
    else:
        x0[:, 6] = 1000.0
        x0[:, 7:13] = lam
        x0[:, 13] = 0.1 * rng.standard_normal(n_seg)
    pert = 1e-3 * rng.standard_normal((n_seg, ndim))
    if ndim == 14:
        pert[:, 6] = 0.0
    x0 += pert
    t0 = np.zeros(n_seg); t1 = np.full(n_seg, dt)
    return dict(x0=x0, t0=t0, t1=t1)


def continuation_batch(n_traj=1024, n_seg_per_traj=200, ndim=12, seed=20180003, tl_hi=10.0, tl_lo=0.05):
    """Config 5: ballistic stack on L2_Anderson_2 sampled at 201 nodes over 20 days,
    per-trajectory thrustLimit ladder, random costates."""
    rng = np.random.default_rng(seed)
    n_nodes = n_seg_per_traj + 1
    tof = 20.0 * DAY / TU
    t = np.linspace(0.0, tof, n_nodes)
    tau0 = rng.uniform(0.0, 1.0, n_traj)
    XC = np.empty((n_traj, n_nodes, ndim))
    tau = tau0[:, None] + t[None, :] / PERIODS[1]
    XC[:, :, :6] = interp_initial_states(tau.ravel(), 2).reshape(n_traj, n_nodes, 6)
    lam = 0.1 * rng.standard_normal((n_traj, n_nodes, 6))
    if ndim == 12:
        XC[:, :, 6:12] = lam
    else:
        XC[:, :, 6] = 1000.0
        XC[:, :, 7:13] = lam
        XC[:, :, 13] = 0.1 * rng.standard_normal((n_traj, n_nodes))
    t_TU = np.broadcast_to(t, (n_traj, n_nodes)).copy()
    thrustLimit = np.geomspace(tl_hi, tl_lo, n_traj)
    return dict(XC_all=XC, t_TU=t_TU, thrustLimit=thrustLimit)
