"""Finer timing of the first iterations of a PGM solve and of create/destroy (config 2)."""
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
Y0, A0, S0 = workloads.cfg2(M, N, K, seed=1234)
Y, _ = bench.pinned_array((M, N))
Y[...] = Y0
ctx = _ffi.context()
plus = [(_ffi.OP_PLUS, 0, 0, 0.0)]
unity_plus = [(_ffi.OP_PLUS, 0, 0, 0.0), (_ffi.OP_UNITY, 0, 0, 0.0)]


def T(label, fn):
    ctx.sync()
    t0 = time.perf_counter()
    r = fn()
    ctx.sync()
    print("  %-28s %8.2f ms" % (label, 1e3 * (time.perf_counter() - t0)), flush=True)
    return r


for rep in range(3):
    print("rep", rep)
    prob = T("create+upload", lambda: pnmf.Problem(Y, A0, S0))
    T("pgm_begin", lambda: prob.pgm_begin(plus, unity_plus, False, (0.0, 0.0)))
    for i in range(5):
        T("pgm_run(1) #%d" % i, lambda: prob.pgm_run(1))
    T("pgm_run(8)", lambda: prob.pgm_run(8))
    T("pgm_run(16)", lambda: prob.pgm_run(16))
    T("pgm_run(171)", lambda: prob.pgm_run(171))
    T("download", lambda: (prob.get(_ffi.A), prob.get(_ffi.S), prob.get(_ffi.GA), prob.get(_ffi.GS)))
    T("destroy", lambda: prob.close())
