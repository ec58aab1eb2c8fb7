"""NMF / constrained matrix factorisation with the API of proxmin/nmf.py, computed on a B200.

``nmf(Y, A, S, ...)`` keeps the reference's signature and in-place semantics (nmf.py:96-203).  Y, A,
S are uploaded once, the whole iteration loop runs on the device through the C ABI
(``pmx_nmf_*``) and A, S are written back into the caller's arrays at the end.  With a user
``callback`` the factors are copied back every iteration (documented slow path).

Multi-GPU (one process per GPU, communicator initialised with ``proxmin_b200.init_distributed``):
each rank passes its own column stripe of Y and S; A is replicated.
"""
import ctypes as C
import logging
from functools import partial

import numpy as np

from . import _ffi
from . import algorithms
from . import operators
from . import utils

logger = logging.getLogger("proxmin")


# ------------------------------------------------------------------------------------------
# device-side problem object
# ------------------------------------------------------------------------------------------
class Problem(object):
    """Device-resident (Y, A, S) plus solver state; thin wrapper over ``pmx_nmf``."""

    def __init__(self, Y, A, S, ctx=None, W=None):
        Y = np.asarray(Y)
        M, N = Y.shape
        K = A.shape[1]
        assert A.shape == (M, K) and S.shape == (K, N), "shapes of Y, A, S do not match"
        self.ctx = ctx or _ffi.context()
        self.M, self.N, self.K = M, N, K
        self.handle = C.c_void_p()
        L = _ffi.lib()
        _ffi.check(L.pmx_nmf_create(self.ctx.handle, M, N, K, C.byref(self.handle)))
        self._upload_matrix(Y, L.pmx_nmf_set_Y)
        if W is not None:   # weighted likelihood (nmf.py:25, 40): an M x N matrix next to Y
            W = np.asarray(W)
            assert W.shape == (M, N), "W must have the shape of Y"
            self._upload_matrix(W, L.pmx_nmf_set_W)
        self.set(_ffi.A, A)
        self.set(_ffi.S, S)

    def _upload_matrix(self, Y, setter):
        """M x N host matrix -> tiled device copy, in column blocks (fp32 staging of at most ~256 MB for fp64 /
        non-contiguous input)."""
        M, N = Y.shape
        if Y.dtype == np.float32 and Y.flags.c_contiguous:
            _ffi.check(setter(self.handle, Y.ctypes.data_as(C.c_void_p), N, 0, N))
            return
        blk = max(1, (1 << 26) // max(M, 1))
        for c0 in range(0, N, blk):
            c1 = min(N, c0 + blk)
            part = np.ascontiguousarray(Y[:, c0:c1], dtype=np.float32)
            _ffi.check(setter(self.handle, part.ctypes.data_as(C.c_void_p), c1 - c0, c0, c1 - c0))

    def gradient(self, want_loss=False):
        """Gradients of the likelihood at the current factors into the GA / GS buffers (nmf.py:28-41); returns the
        log-likelihood (nmf.py:13-25) when asked for."""
        loss = C.c_double(0)
        _ffi.check(_ffi.lib().pmx_nmf_gradient(self.handle, C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def close(self):
        if self.handle:
            _ffi.check(_ffi.lib().pmx_nmf_destroy(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, which):
        return (self.M, self.K) if which % 2 == 0 else (self.K, self.N)

    def set(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        assert a.shape == self._shape(which)
        _ffi.check(_ffi.lib().pmx_nmf_set(self.handle, which, a.ctypes.data_as(C.c_void_p)))

    def get(self, which, out=None, dtype=np.float32):
        shape = self._shape(which)
        L = _ffi.lib()
        if out is not None and isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous \
                and out.shape == shape:
            _ffi.check(L.pmx_nmf_get(self.handle, which, out.ctypes.data_as(C.c_void_p)))   # straight into the caller's array
            return out
        buf = np.empty(shape, np.float32)
        _ffi.check(L.pmx_nmf_get(self.handle, which, buf.ctypes.data_as(C.c_void_p)))
        if out is not None:
            out[...] = buf
            return out
        return buf.astype(dtype, copy=False)

    def loss(self):
        v = C.c_double(0)
        _ffi.check(_ffi.lib().pmx_nmf_loss(self.handle, C.byref(v)))
        return v.value

    # -- PGM ----------------------------------------------------------------------------
    def pgm_begin(self, prox_A, prox_S, accelerated=False, e_rel=(1e-3, 1e-3), kernel=0, check_every=8):
        o = _ffi.PgmOpts()
        o.prox_A, o.prox_S = _ffi.make_prox(prox_A), _ffi.make_prox(prox_S)
        o.accelerated = int(bool(accelerated))
        o.e_rel_A, o.e_rel_S = float(e_rel[0]), float(e_rel[1])
        o.kernel, o.check_every = kernel, check_every
        _ffi.check(_ffi.lib().pmx_nmf_pgm_begin(self.handle, C.byref(o)))

    def pgm_run(self, n_iter):
        it, cA, cS = C.c_int(0), C.c_int(0), C.c_int(0)
        sA, sS = C.c_float(0), C.c_float(0)
        _ffi.check(_ffi.lib().pmx_nmf_pgm_run(self.handle, n_iter, C.byref(it), C.byref(cA), C.byref(cS),
                                              C.byref(sA), C.byref(sS)))
        return it.value, (bool(cA.value), bool(cS.value)), (sA.value, sS.value)

    # -- adaprox ------------------------------------------------------------------------
    def adaprox_begin(self, prox_A, prox_S, scheme, b2, eps, p, e_rel, check_convergence, prox_max_iter,
                      has_vhat=False, alpha=None, kernel=0):
        o = _ffi.AdaproxOpts()
        o.has_prox_A, o.has_prox_S = int(prox_A is not None), int(prox_S is not None)
        o.prox_A, o.prox_S = _ffi.make_prox(prox_A or []), _ffi.make_prox(prox_S or [])
        o.scheme = _ffi.SCHEMES[scheme]
        o.b2, o.eps, o.p = float(b2), float(eps), float(p)
        o.e_rel_A, o.e_rel_S = float(e_rel[0]), float(e_rel[1])
        o.check_convergence, o.prox_max_iter = int(bool(check_convergence)), int(prox_max_iter)
        o.has_vhat, o.kernel = int(bool(has_vhat)), kernel
        if alpha is None:
            o.step_mode = 0
        else:
            o.step_mode, o.alpha_A, o.alpha_S = 1, float(alpha[0]), float(alpha[1])
        _ffi.check(_ffi.lib().pmx_nmf_adaprox_begin(self.handle, C.byref(o)))

    def adaprox_run(self, n_iter, b1, b1_prev):
        b1 = np.ascontiguousarray(b1, dtype=np.float64)
        b1p = np.ascontiguousarray(b1_prev, dtype=np.float64)
        it, cA, cS = C.c_int(0), C.c_int(0), C.c_int(0)
        subA, subS = C.c_longlong(0), C.c_longlong(0)
        pd = C.POINTER(C.c_double)
        _ffi.check(_ffi.lib().pmx_nmf_adaprox_run(self.handle, n_iter, b1.ctypes.data_as(pd), b1p.ctypes.data_as(pd),
                                                  C.byref(it), C.byref(cA), C.byref(cS), C.byref(subA), C.byref(subS)))
        return it.value, (bool(cA.value), bool(cS.value)), (subA.value, subS.value)

    # -- bsdmm --------------------------------------------------------------------------
    def bsdmm_begin(self, prox_A, prox_S, proxs_g_A, proxs_g_S, e_rel, e_abs, kernel=0):
        o = _ffi.BsdmmOpts()
        o.prox_A, o.prox_S = _ffi.make_prox(prox_A), _ffi.make_prox(prox_S)
        o.n_g_A, o.n_g_S = len(proxs_g_A), len(proxs_g_S)
        for i, g in enumerate(proxs_g_A):
            o.proxs_g_A[i] = _ffi.make_prox(g)
        for i, g in enumerate(proxs_g_S):
            o.proxs_g_S[i] = _ffi.make_prox(g)
        o.e_rel_A, o.e_rel_S = float(e_rel[0]), float(e_rel[1])
        o.e_abs_A, o.e_abs_S = float(e_abs[0]), float(e_abs[1])
        o.kernel = kernel
        _ffi.check(_ffi.lib().pmx_nmf_bsdmm_begin(self.handle, C.byref(o)))

    def bsdmm_run(self, n_iter):
        it, cA, cS = C.c_int(0), C.c_int(0), C.c_int(0)
        _ffi.check(_ffi.lib().pmx_nmf_bsdmm_run(self.handle, n_iter, C.byref(it), C.byref(cA), C.byref(cS)))
        return it.value, [bool(cA.value), bool(cS.value)]


def _check_W(W):
    """Scalar W == 1 -> None; an M x N weight matrix is returned as it is.  Other scalars are not part of the
    reference's contract (its docstring asks for an M x N matrix)."""
    if np.ndim(W) == 0:
        if W == 1:
            return None
        raise NotImplementedError("scalar weights other than 1: pass an M x N weight matrix (nmf.py:19)")
    return np.asarray(W)


def _device_triplet(A, S, Y):
    ctx = _ffi.context()
    A32 = np.ascontiguousarray(A, dtype=np.float32)
    S32 = np.ascontiguousarray(S, dtype=np.float32)
    Y32 = np.ascontiguousarray(Y, dtype=np.float32)
    M, K = A32.shape
    K2, N = S32.shape
    assert K == K2 and Y32.shape == (M, N)
    return ctx, A32, S32, Y32, M, N, K


def log_likelihood(*X, Y=0, W=1):
    """sum(W (Y - A S)^2) / 2 (nmf.py:13-25), one fused residual pass on the device."""
    Wm = _check_W(W)
    A, S = X
    prob = Problem(Y, A, S, W=Wm)
    try:
        v = prob.loss()
    finally:
        prob.close()
    return np.result_type(A.dtype, S.dtype).type(v)


def grad_likelihood(*X, Y=0, W=1):
    """(D S^T, A^T D) with D = W (A S - Y) (nmf.py:28-41), one pass over Y (and W), no M x N temporary."""
    Wm = _check_W(W)
    A, S = X
    dt = np.result_type(A.dtype, S.dtype)
    if Wm is not None:
        prob = Problem(Y, A, S, W=Wm)
        try:
            prob.gradient()
            return prob.get(_ffi.GA, dtype=dt), prob.get(_ffi.GS, dtype=dt)
        finally:
            prob.close()
    ctx, A32, S32, Y32, M, N, K = _device_triplet(A, S, Y)
    dY, dA, dS = ctx.upload(Y32), ctx.upload(A32), ctx.upload(S32)
    dGA, dGS = ctx.malloc(4 * M * K), ctx.malloc(4 * K * N)
    try:
        _ffi.check(_ffi.lib().pmx_nmf_grad(ctx.handle, dY, dA, dS, M, N, K, dGA, dGS, None, 0))
        GA, GS = np.empty((M, K), np.float32), np.empty((K, N), np.float32)
        ctx.d2h(GA, dGA)
        ctx.d2h(GS, dGS)
    finally:
        for p in (dY, dA, dS, dGA, dGS):
            ctx.free(p)
    return GA.astype(dt, copy=False), GS.astype(dt, copy=False)


def _lipschitz(A, S):
    ctx = _ffi.context()
    A32 = np.ascontiguousarray(A, dtype=np.float32)
    S32 = np.ascontiguousarray(S, dtype=np.float32)
    M, K = A32.shape
    N = S32.shape[1]
    dA, dS = ctx.upload(A32), ctx.upload(S32)
    try:
        la, ls = C.c_float(0), C.c_float(0)
        _ffi.check(_ffi.lib().pmx_nmf_lipschitz(ctx.handle, dA, dS, M, N, K, C.byref(la), C.byref(ls)))
    finally:
        ctx.free(dA)
        ctx.free(dS)
    dt = np.result_type(A.dtype, S.dtype).type
    return dt(la.value), dt(ls.value)


def step_A(A, S):
    """1 / lambda_max(S S^T) (nmf.py:44-45)."""
    return 1 / _lipschitz(A, S)[0]


def step_S(A, S):
    """1 / lambda_max(A^T A) (nmf.py:48-49)."""
    return 1 / _lipschitz(A, S)[1]


def step_pgm(*X, it=None, W=1):
    """Lipschitz step sizes for both factors (nmf.py:52-65, W == 1 branch).  With a weight matrix the reference
    evaluates ``if W == 1`` on an array and raises ValueError (nmf.py:63); the same statement raises it here.  (Its
    weighted branch builds M*N x M*K sparse operators and is not reachable; weighted PGM needs a user ``step``.)"""
    if W == 1:
        pass
    else:
        raise NotImplementedError("weighted step_pgm (sparse eigs branch, nmf.py:66-88) is outside the B200 hot path")
    A, S = X
    la, ls = _lipschitz(A, S)
    return 1 / la, 1 / ls


def step_adaprox(*X, it=None):
    """Per-component steps mean/10 (nmf.py:91-93): column means of A (shape K) and row means of S (shape K x 1),
    reduced on the device (the fused adaprox loop computes the same means inside the solver object)."""
    A, S = X
    ctx = _ffi.context()
    out = []
    for X_, axis in ((A, 0), (S, 1)):
        x32 = np.ascontiguousarray(X_, dtype=np.float32)
        rows, cols = x32.shape
        n = cols if axis == 0 else rows
        d = ctx.upload(x32)
        try:
            sums = (C.c_double * n)()
            _ffi.check(_ffi.lib().pmx_axis_sum(ctx.handle, d, rows, cols, axis, sums))
        finally:
            ctx.free(d)
        mean = (np.array(sums[:]) / (rows if axis == 0 else cols)).astype(X_.dtype if X_.dtype.kind == "f" else np.float64)
        out.append(mean)
    return (out[0] / 10, out[1][:, None] / 10)


def nmf(
    Y,
    A,
    S,
    W=1,
    prox_A=operators.prox_plus,
    prox_S=operators.prox_plus,
    algorithm=algorithms.pgm,
    step=None,
    max_iter=1000,
    e_rel=1e-3,
    callback=None,
    **algorithm_args
):
    """Non-negative / constrained matrix factorisation, minimise ||Y - A S||^2 (nmf.py:96-203).

    Same arguments and return values as the reference; A and S are updated in place."""
    assert algorithm in [algorithms.pgm, algorithms.adaprox, algorithms.bsdmm]

    grad = partial(grad_likelihood, Y=Y, W=W)
    X = [A, S]
    prox = [prox_A, prox_S]

    if algorithm is algorithms.pgm:
        if step is None:
            step = partial(step_pgm, W=W)
        return algorithm(X, grad, step, prox=prox, max_iter=max_iter, e_rel=e_rel, callback=callback,
                         **algorithm_args)

    if algorithm is algorithms.adaprox:
        if step is None:
            step = step_adaprox
        return algorithm(X, grad, step, prox=prox, max_iter=max_iter, e_rel=e_rel, callback=callback,
                         **algorithm_args)

    if algorithm is algorithms.bsdmm:
        if step is not None:
            # the reference raises UnboundLocalError here (nmf.py:187-198, SURVEY 8 a-Q)
            raise UnboundLocalError("cannot access local variable 'step_f' where it is not associated with a value")
        return algorithms._bsdmm_nmf(Y, A, S, W, prox_A, prox_S, max_iter=max_iter, e_rel=e_rel,
                                     callback=callback, **algorithm_args)
