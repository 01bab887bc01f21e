import os, sys, time
os.environ["LTO_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, ".")
import numpy as np
from lowthrustopt_b200 import capi, synthetic as S
h = capi.Handle(0)
n = 65536
b = S.direct_batch(n, nstate=7)
pin = {k: capi.PinnedBuffer(v.shape) for k, v in b.items() if k in ("Xa", "Xb", "ua", "ub", "ta", "tb")}
for k in pin: pin[k].array[...] = b[k]
out = None
for i in range(4):
    t0 = time.perf_counter()
    r = h.direct(*[pin[k].array for k in ("Xa", "Xb", "ua", "ub", "ta", "tb")], nsteps=10, out=out)
    dt = time.perf_counter() - t0
    out = r
    print("call %d: %.3f ms wall" % (i, dt * 1e3), file=sys.stderr)
