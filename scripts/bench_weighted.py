"""Timing of the weighted gradient pass (nmf.py:28-41 with an M x N weight matrix) at the config-2 shape."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proxmin_b200 import _ffi, workloads
from proxmin_b200 import nmf as pnmf

M, N, K = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 64
Y, A, S = workloads.cfg2(M, N, K, seed=1234)
rng = np.random.default_rng(5)
W = (0.25 + 0.75 * rng.random((M, N), dtype=np.float32))
ctx = _ffi.context()
for tag, w in (("unweighted", None), ("weighted", W)):
    prob = pnmf.Problem(Y, A, S, W=w)
    for _ in range(3):
        prob.gradient()
    ctx.sync()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        prob.gradient()
    ctx.sync()
    ms = (time.perf_counter() - t0) / n * 1e3
    bytes_ = 4.0 * M * N * (2 if w is not None else 1)
    print("%-10s gradient pass %.3f ms  (%.2f TB/s of Y%s)" % (tag, ms, bytes_ / ms / 1e9, " + W" if w is not None else ""))
    prob.close()
