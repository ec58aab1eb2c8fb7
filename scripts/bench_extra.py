"""Secondary measurements (not the driver's bench line): BASELINE configs 3, 4, 5 on one B200.

    python scripts/bench_extra.py [--quick]

Prints one JSON line per config with iterations/s and, for the HBM-bound ADMM kernel, the roofline fraction.
"""
import ctypes as C
import json
import os
import sys
import time
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import proxmin_b200 as pmx  # noqa: E402
from proxmin_b200 import _ffi, workloads  # noqa: E402
from proxmin_b200 import nmf as pnmf  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def cfg3_adaprox(M, N, K, iters):
    Y, A, S = workloads.cfg2(M, N, K)
    ctx = _ffi.context()
    prob = pnmf.Problem(Y, A, S)
    plus = [(_ffi.OP_PLUS, 0, 0, 0.0)]
    prob.adaprox_begin(plus, plus, "amsgrad", 0.999, 1e-8, 0.25, (1e-3, 1e-3), False, 1000)
    b1 = np.full(iters + 3, 0.9)
    prob.adaprox_run(3, b1[:3], np.roll(b1, 1)[:3])
    ctx.sync()
    t0 = time.perf_counter()
    done, _, sub = prob.adaprox_run(iters, b1[3:], np.roll(b1, 1)[3:])
    ctx.sync()
    dt = time.perf_counter() - t0
    prob.close()
    return {"config": "cfg3 nmf adaprox/amsgrad plus/plus Y=%dx%d K=%d" % (M, N, K), "it_per_s": done / dt,
            "ms_per_it": 1e3 * dt / done, "sub_iterations": [int(sub[0]), int(sub[1])], "iterations": done}


def cfg4_admm(n, iters):
    b, X = workloads.cfg4(n)
    ctx = _ffi.context()
    L = _ffi.lib()
    o = _ffi.AdmmOpts()
    o.n_g = 1
    o.proxs_g[0] = _ffi.make_prox([(_ffi.OP_SOFT, 1, 0, 0.5)])
    o.e_rel, o.e_abs, o.dual_uses_step_g = 0.0, 0.0, 0   # e_rel = 0: never converges, fixed iteration count
    h = C.c_void_p()
    _ffi.check(L.pmx_admm_create(ctx.handle, n, C.byref(o), C.byref(h)))
    vp = C.c_void_p
    _ffi.check(L.pmx_admm_set(h, X.ctypes.data_as(vp), b.ctypes.data_as(vp)))
    it, conv = C.c_int(0), C.c_int(0)
    err = (C.c_double * 16)()
    _ffi.check(L.pmx_admm_run(h, 0.5, 20, C.byref(it), C.byref(conv), err))  # warm-up
    _ffi.check(L.pmx_admm_set(h, X.ctypes.data_as(vp), b.ctypes.data_as(vp)))
    ctx.sync()
    t0 = time.perf_counter()
    _ffi.check(L.pmx_admm_run(h, 0.5, iters, C.byref(it), C.byref(conv), err))
    ctx.sync()
    dt = time.perf_counter() - t0
    _ffi.check(L.pmx_admm_destroy(h))
    done = it.value - 1
    alg = 7 * 4 * n  # read X,Z,U,b; write X,Z,U
    gbs = alg * done / dt / 1e9
    return {"config": "cfg4 admm + prox_soft LASSO n=%d" % n, "it_per_s": done / dt, "ms_per_it": 1e3 * dt / done,
            "iterations": done, "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks(), "frac": gbs / peaks(),
                                             "unit": "GB/s", "note": "whole loop incl. finalize kernel and host polling"}}


def cfg5_bsdmm(M, N, K, iters):
    Y, A, S = workloads.cfg5(M, N, K)
    ctx = _ffi.context()
    prob = pnmf.Problem(Y, A, S)
    gA = [[(_ffi.OP_PLUS, 0, 0, 0.0)], [(_ffi.OP_UNITY, 0, 0, 0.0)]]
    gS = [[(_ffi.OP_PLUS, 0, 0, 0.0)], [(_ffi.OP_SOFT, 1, 0, 0.01)]]
    prob.bsdmm_begin([], [], gA, gS, (0.0, 0.0), (0.0, 0.0))
    prob.bsdmm_run(2)
    ctx.sync()
    t0 = time.perf_counter()
    done, _ = prob.bsdmm_run(iters)
    ctx.sync()
    dt = time.perf_counter() - t0
    prob.close()
    return {"config": "cfg5 nmf bsdmm CMF Y=%dx%d K=%d" % (M, N, K), "it_per_s": done / dt, "ms_per_it": 1e3 * dt / done,
            "iterations": done, "gradient_kernel": "tcgen05" if K <= 64 else "simt (K > 64)"}


def main():
    quick = "--quick" in sys.argv
    out = []
    out.append(cfg4_admm(1_000_000 if quick else 10_000_000, 200))
    out.append(cfg3_adaprox(2048 if quick else 8192, 8192 if quick else 65536, 64, 10 if quick else 30))
    out.append(cfg5_bsdmm(1024 if quick else 4096, 8192 if quick else 131072, 64, 5 if quick else 10))
    if not quick:
        out.append(cfg5_bsdmm(4096, 131072, 128, 3))
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
