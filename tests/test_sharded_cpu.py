"""CPU: the N > 1 path (lowthrustopt_b200/sharded.py) on world_size-2 gloo, plus the shard plan's arithmetic."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lowthrustopt_b200 import sharded, synthetic as S

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


@pytest.mark.parametrize("n,world,chunks", [(0, 1, 2), (1, 8, 2), (7, 2, 2), (1024, 8, 4), (65536, 3, 2), (5, 4, 16)])
def test_shard_plan_covers_every_unit_once(n, world, chunks):
    p = sharded.ShardPlan(n, world, chunks)
    seen = np.zeros(p.padded, dtype=int)
    for c in range(p.n_chunks):
        for r in range(world):
            u0, cnt = p.local(r, c)
            assert 0 <= cnt <= p.cs
            seen[u0:u0 + cnt] += 1
            if cnt:
                assert p.owner(u0) == (c, r) and p.owner(u0 + cnt - 1) == (c, r)
    assert np.all(seen[:n] == 1) and np.all(seen[n:] == 0) and p.padded >= n and (p.padded - n < world * p.cs or n == 0)


def test_world2_gloo_allgather_matches_single_process(tmp_path, oracle):
    out = str(tmp_path / "gathered.npz")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    g = np.load(out)
    assert list(g["plan"]) == [2, 2, 8, 2, 3, 12]
    c = S.continuation_batch(n_traj=7, n_seg_per_traj=5, ndim=12)
    X = c["XC_all"]
    x0 = X[:, :-1].reshape(-1, 12); t0 = c["t_TU"][:, :-1].ravel(); t1 = c["t_TU"][:, 1:].ravel()
    tl = np.repeat(c["thrustLimit"], 5)
    xe, phi, st, na, nt = oracle.indirect_prop_jac(x0, t0, t1, oracle.iparams(0.05, p=1.0), thrustLimit=tl, rho=np.ones(35))
    assert g["defect"].shape == (7, 5, 12) and g["phi"].shape == (7, 5, 12, 12)
    assert np.array_equal(g["defect"].reshape(-1, 12), xe - X[:, 1:].reshape(-1, 12))        # same code, same order: bitwise
    assert np.array_equal(g["phi"].reshape(-1, 12, 12), phi.transpose(0, 2, 1))
    assert np.array_equal(g["nsteps"].reshape(-1, 2)[:, 0], na) and np.all(g["status"] == 0)
    xe0, st0, _, _ = oracle.indirect_prop(x0, t0, t1, oracle.iparams(0.05, p=1.0), thrustLimit=tl, rho=np.ones(35))
    assert np.array_equal(g["d_defect"].reshape(-1, 12), xe0 - X[:, 1:].reshape(-1, 12))
    b = S.direct_batch(11, nstate=7, seed=3)
    d, e, J, st = oracle.direct_jac_var(b["Xa"], b["Xb"], b["ua"], b["ub"], b["ta"], b["tb"])
    assert np.array_equal(g["dir_defect"], d) and np.array_equal(g["dir_jac"], J.transpose(0, 2, 1)) and np.array_equal(g["dir_errors"], e)
