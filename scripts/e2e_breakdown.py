"""Where the end-to-end time of one nmf.nmf() solve goes (config 2, pinned host arrays)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from proxmin_b200 import _ffi, workloads  # noqa: E402
from proxmin_b200 import nmf as pnmf  # noqa: E402

M, N, K = 8192, 65536, 64
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
Y0, A0, S0 = workloads.cfg2(M, N, K, seed=1234)
Y, _ = bench.pinned_array((M, N))
Y[...] = Y0
ctx = _ffi.context()
plus = [(_ffi.OP_PLUS, 0, 0, 0.0)]
unity_plus = [(_ffi.OP_PLUS, 0, 0, 0.0), (_ffi.OP_UNITY, 0, 0, 0.0)]
for rep in range(3):
    ctx.sync()
    t0 = time.perf_counter()
    prob = pnmf.Problem(Y, A0, S0)
    ctx.sync()
    t1 = time.perf_counter()
    prob.pgm_begin(plus, unity_plus, False, (0.0, 0.0))
    prob.pgm_run(steps)
    ctx.sync()
    t2 = time.perf_counter()
    A = prob.get(_ffi.A); S = prob.get(_ffi.S); GA = prob.get(_ffi.GA); GS = prob.get(_ffi.GS)
    t3 = time.perf_counter()
    prob.close()
    ctx.sync()
    t4 = time.perf_counter()
    print("create+upload %.1f ms (%.1f GB/s)  loop %.1f ms (%.3f ms/it)  download %.1f ms  destroy %.1f ms  total %.1f ms -> %.1f it/s"
          % (1e3 * (t1 - t0), Y.nbytes / (t1 - t0) / 1e9, 1e3 * (t2 - t1), 1e3 * (t2 - t1) / steps, 1e3 * (t3 - t2),
             1e3 * (t4 - t3), 1e3 * (t4 - t0), steps / (t4 - t0)), flush=True)
