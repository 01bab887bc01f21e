#!/bin/bash
# round 2, visit B: first run of the half-column K3 (lto_indirect_hc.cu): parity tests of the indirect 12-dim path, bench vs the old kernel
O=gpurun_out/r2b; mkdir -p $O
timeout 40 python - > $O/first.log 2>&1 <<'PY'
import numpy as np, time
from lowthrustopt_b200 import capi, synthetic as S
from oracle import oracle as O
h = capi.Handle(0)
for n in (40, 96, 1000, 20000, 131072):
    b = S.indirect_batch(n, ndim=12, seed=202)
    p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
    t = time.time(); r = h.indirect(b["x0"], b["t0"], b["t1"], params=p); dt = time.time() - t
    xo, Po, so, nao, nto = O.indirect_prop_jac(b["x0"], b["t0"], b["t1"], O.iparams(0.05, p=1.0, rho=1.0), nthreads=O.num_threads())
    ex = np.abs(r["defect"] - xo).max(); ep = np.abs(r["phi"].transpose(0, 2, 1) - Po).max()
    print(n, "status", r["status"].max(), "ex %.2e ep %.2e" % (ex, ep), "steps", r["nsteps"][:, 0].mean(), nao.mean(), "t %.3f" % dt, flush=True)
h.close()
PY
rc=$?; echo "first rc=$rc" >> $O/first.log; tail -7 $O/first.log
if [ $rc -ne 0 ]; then echo "first run failed: stopping"; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_scale.py tests/test_gpu_newton.py tests/test_gpu_solvers.py -m gpu -x -q -k "indirect or newton or solve or continuation or densify" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for v in hc cw; do
  LTO_K3=$v timeout 120 python bench.py --workload indirect12 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_indirect12_$v.json 2> $O/bench_indirect12_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_indirect12_$v.json").read().strip().splitlines()[-1])
    print("$v", d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"))
except Exception as e: print("$v failed", e)
PY
done
LTO_ICW_PROF=1 timeout 120 python tools/ihc_prof.py > $O/prof.log 2>&1; tail -20 $O/prof.log
