"""TEST INFRASTRUCTURE: one rank of a world_size-N gloo run of lowthrustopt_b200.sharded on CPU.
The propagation callable is the CPU oracle (there is no GPU here); what is under test is the shard plan,
the padding, the chunked all-gather and the global ordering.  Launched by tests/test_sharded_cpu.py via
torch.distributed.run; rank 0 writes the gathered arrays to argv[1]."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from lowthrustopt_b200 import sharded, synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def indirect_compute(sh, params, jac, u0, cnt, outs):
    nn, nd = sh.n_nodes, sh.nd
    XC = sh.XC[u0:u0 + cnt].numpy(); t = sh.t[u0:u0 + cnt].numpy()
    tl = np.repeat(sh.tl[u0:u0 + cnt].numpy(), nn - 1); rho = np.repeat(sh.rho[u0:u0 + cnt].numpy(), nn - 1)
    x0 = XC[:, :-1].reshape(-1, nd); t0 = t[:, :-1].ravel(); t1 = t[:, 1:].ravel()
    ip = O.iparams(0.05, p=params["p"], rho=1.0)
    if jac:
        xe, phi, st, na, nt = O.indirect_prop_jac(x0, t0, t1, ip, thrustLimit=tl, rho=rho)
        outs["phi"][:cnt] = torch.from_numpy(phi.transpose(0, 2, 1).reshape(cnt, nn - 1, nd, nd).copy())
    else:
        xe, st, na, nt = O.indirect_prop(x0, t0, t1, ip, thrustLimit=tl, rho=rho)
    outs["defect"][:cnt] = torch.from_numpy((xe - XC[:, 1:].reshape(-1, nd)).reshape(cnt, nn - 1, nd))
    outs["status"][:cnt] = torch.from_numpy(st.reshape(cnt, nn - 1))
    outs["nsteps"][:cnt] = torch.from_numpy(np.stack([na, nt], axis=1).reshape(cnt, nn - 1, 2))


def direct_compute(sh, params, jac, u0, cnt, outs):
    i = {k: v[u0:u0 + cnt].numpy() for k, v in sh.inp.items()}
    if jac:
        d, e, J, st = O.direct_jac_var(i["Xa"], i["Xb"], i["ua"], i["ub"], i["ta"], i["tb"], nsteps=sh.nsteps)
        outs["jac"][:cnt] = torch.from_numpy(J.transpose(0, 2, 1).copy())
    else:
        d, e, st, _ = O.direct_defect(i["Xa"], i["Xb"], i["ua"], i["ub"], i["ta"], i["tb"], nsteps=sh.nsteps)
    outs["defect"][:cnt] = torch.from_numpy(d); outs["errors"][:cnt] = torch.from_numpy(e); outs["status"][:cnt] = torch.from_numpy(st)


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    n_traj, spt, nd = 7, 5, 12                     # 7 trajectories over 2 ranks x 2 chunks: ragged, padded tail
    c = S.continuation_batch(n_traj=n_traj, n_seg_per_traj=spt, ndim=nd)
    sh = sharded.ShardedIndirect(None, n_traj, spt + 1, nd, "cpu", n_chunks=2, compute=indirect_compute)
    # only the solver rank (0) holds the inputs; the others pass nothing
    if rank == 0:
        sh.load(c["XC_all"], c["t_TU"], c["thrustLimit"], 1.0)
    else:
        sh.load()
    rj, plan = sh.run({"p": 1.0}, jac=True)
    rd, _ = sh.run({"p": 1.0}, jac=False)
    b = S.direct_batch(11, nstate=7, seed=3)
    sd = sharded.ShardedDirect(None, 11, 7, "cpu", n_chunks=3, compute=direct_compute)
    sd.load(b if rank == 0 else None)
    rdir, pdir = sd.run(None, jac=True)
    res = {k: v.numpy() for k, v in rj.items()}
    res.update({"d_" + k: v.numpy() for k, v in rd.items()})
    res.update({"dir_" + k: v.numpy() for k, v in rdir.items()})
    res["plan"] = np.array([plan.cs, plan.n_chunks, plan.padded, pdir.cs, pdir.n_chunks, pdir.padded])
    # every rank must hold the same full arrays
    chk = torch.tensor([float(np.sum(res["defect"])), float(np.sum(res["phi"])), float(np.sum(res["dir_jac"]))], dtype=torch.float64)
    allc = [torch.zeros_like(chk) for _ in range(dist.get_world_size())]
    dist.all_gather(allc, chk)
    assert all(torch.equal(a, allc[0]) for a in allc)
    if rank == 0:
        np.savez(out_path, **res)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
