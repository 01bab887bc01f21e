"""Generates tests/golden/bangbang_v1.json: the indirect demo (CRTBP_Multishoot_indirect_demo.jl:166-281) solved with the
ORACLE-backed host loops down to rho = 1e-4 (the demo's rho_target, :276-281) -- a converged bang-bang trajectory whose
segments cross the thrust switches.  Used as inputs of the rho = 1e-3 / 1e-4 parity cases (tests/test_gpu_parity_scale.py).

    python tests/golden/make_bangbang.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from lowthrustopt_b200 import capi, solvers as S      # noqa: E402  (host loops only; every propagation goes through the oracle)
from oracle_backend import OracleBackend               # noqa: E402

MU, DU, TU = capi.MU, capi.DU, capi.TU
cpu = OracleBackend()
X0t, X0, Xft, Xf = S.demo_fixtures()
XCg, t_TU, tau1, tau2, s0, sf = S.trajectory_stack_guess(X0, Xf, backend=cpu)
out = S.multiShoot_CRTBP_direct(XCg[:6].copy(), np.zeros((3, 30)), tau1, tau2, t_TU, np.zeros(3), np.zeros(3), MU, DU, TU, 30, 10, 1e3,
                                2000.0, X0t, X0, Xft, Xf, backend=cpu)
rng = np.random.default_rng(42)
XC = np.vstack([out[0], 0.1 * rng.standard_normal((6, 30))])
XC[:6, 0] = s0[:6]; XC[:6, -1] = sf[:6]
XC[:, 1:-1] += 1e-10 * rng.standard_normal((12, 28))
XC, d, st = S.multiShoot_CRTBP_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 10.0, False, True, 10, 2.0, 1.0, backend=cpu)
XC, d, st = S.multiShoot_CRTBP_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 10.0, False, False, 50, 2.0, 1.0, backend=cpu)
XC, d, st = S.multiShoot_CRTBP_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 0.05, False, False, 30, 1.0, 1.0, backend=cpu)
assert st == 0
res = {}
for rho in (1e-2, 1e-3, 1e-4):
    XC, d, st = S.reduceFuel_indirect(XC, t_TU, MU, DU, TU, 30, 1e3, 0.05, 1.0 if rho == 1e-2 else rho * 10, rho, backend=cpu)
    assert st == 0 and np.abs(d).max() < 1e-10
    res["%g" % rho] = XC.T.tolist()          # 30 nodes x 12
with open(os.path.join(HERE, "bangbang_v1.json"), "w") as f:
    json.dump({"t_TU": t_TU.tolist(), "thrustLimit": 0.05, "mass": 1e3, "p": 1.0, "XC_nodes": res,
               "how": "tests/golden/make_bangbang.py (oracle-backed solvers.reduceFuel_indirect, defects < 1e-10)"}, f)
lvn = np.linalg.norm(XC[9:12], axis=0)
print("rho 1e-4 converged; |lv| along the nodes:", np.round(lvn, 3))
