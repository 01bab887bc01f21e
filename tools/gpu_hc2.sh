#!/bin/bash
# K3-hc2 (LTO_K3=hc2 in the variant library): smoke-sized parity first, then the experimental-layout test, the bench line and the cycle counters
O=gpurun_out/hc2; mkdir -p $O
export LTO_B200_LIB=$PWD/tools/experiments/lib/liblto_k3x.so
export LTO_K3=hc2
timeout 90 python - > $O/first.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from lowthrustopt_b200 import capi, synthetic as S
from oracle import oracle as O
O.build()
h = capi.Handle(0)
for n in (5, 64, 1000, 20000):
    b = S.indirect_batch(n, ndim=12, seed=11)
    p = capi.indirect_params(p=1.0, rho=1.0, thrustLimit=0.05)
    r = h.indirect(b["x0"], b["t0"], b["t1"], params=p)
    ip = O.iparams(0.05, p=1.0, rho=1.0)
    xo, Po, so, nao, nto = O.indirect_prop_jac(b["x0"], b["t0"], b["t1"], ip, nthreads=8)
    print(n, "status", np.unique(r["status"]), "x", np.abs(r["defect"] - xo).max(), "phi", np.abs(r["phi"].transpose(0, 2, 1) - Po).max(),
          "na", np.abs(r["nsteps"][:, 0] - nao).max(), flush=True)
h.close()
PY
echo "first rc=$?"; tail -5 $O/first.log
grep -q "^20000 status \[0\]" $O/first.log || { echo "FIRST FAILED"; exit 0; }
timeout 300 python bench.py --workload indirect12 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_indirect12.json 2> $O/bench_indirect12.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_indirect12.json").read().splitlines() if l.startswith("{")][-1])
    print("indirect12 hc2", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
except Exception as e: print("bench failed", e); print(open("$O/bench_indirect12.err").read()[-800:])
PY
IHC_NCW=9 timeout 60 python tools/ihc_prof.py > $O/prof.log 2>&1; head -6 $O/prof.log
unset LTO_B200_LIB LTO_K3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "experimental" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
