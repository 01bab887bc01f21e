#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 40 python - > $O/first.log 2>&1 <<'PY'
import numpy as np
from lowthrustopt_b200 import capi, synthetic as S
from oracle import oracle as O
h = capi.Handle(0)
for n in (40, 1000, 20000):
    b = S.indirect_batch(n, ndim=12, seed=202)
    for norm in (capi.LTO_NORM_STATE_SENS, capi.LTO_NORM_STATE):
        p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05, err_norm=norm)
        r = h.indirect(b["x0"], b["t0"], b["t1"], params=p)
        xo, Po, so, nao, nto = O.indirect_prop_jac(b["x0"], b["t0"], b["t1"], O.iparams(0.05, p=1.0, rho=1.0), nthreads=O.num_threads())
        ex = np.abs(r["defect"] - xo).max(); ep = (np.abs(r["phi"].transpose(0, 2, 1) - Po).reshape(n, -1).max(axis=1) / np.maximum(1, np.abs(Po).max(axis=(1, 2)))).max()
        print(n, "norm", norm, "status", r["status"].max(), "ex %.2e ep %.2e" % (ex, ep), "steps", r["nsteps"][:, 0].mean(), flush=True)
h.close()
PY
rc=$?; echo "first rc=$rc" >> $O/first.log; tail -7 $O/first.log
if [ $rc -ne 0 ]; then echo "first run failed: stopping"; exit 0; fi
LTO_ICW_PROF=1 timeout 60 python tools/ihc_prof.py > $O/prof.log 2>&1; head -6 $O/prof.log
