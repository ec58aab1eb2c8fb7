"""Helpers with the names of proxmin/utils.py that belong to the hot-path contract.

Only the callback protocol, the Nesterov sequence and the tuple helpers live on the host;
norms, Lipschitz constants and the ADMM variable updates are device kernels (see algorithms.py).
"""
import ctypes as C

import numpy as np

from . import _ffi


def _copy_tuple(X):
    return tuple(item.copy() for item in X)


def _as_tuple(X):
    if type(X) in [list, tuple]:
        return X
    return (X,)


class Traceback(object):
    """Callback that stores a copy of every iterate (utils.py:104-116)."""

    def __init__(self):
        self._trace = []

    def __call__(self, *X, it=None):
        self._trace.append(tuple(x.copy() for x in X))

    @property
    def trace(self):
        return self._trace

    def clear(self):
        self._trace = []


class NullCallback(object):
    def __call__(self, *X, it):
        pass


class NesterovAccelerator(object):
    """t-sequence of FISTA (utils.py:193-206): ``omega`` advances on every read."""

    def __init__(self, accelerated=False):
        self.t = 1.0
        self.accelerated = accelerated

    @property
    def omega(self):
        if self.accelerated:
            t_ = 0.5 * (1 + np.sqrt(4 * self.t * self.t + 1))
            om = (self.t - 1) / t_
            self.t = t_
            return om
        return 0


def get_spectral_norm(L):
    """Squared spectral norm of a dense matrix = lambda_max(L^T L) (utils.py:14-35, dense branch),
    computed on the device (Gram kernel + one-CTA eigen-solver)."""
    if L is None:
        return 1
    if hasattr(L, "spectral_norm"):
        return L.spectral_norm
    L = np.asarray(L)
    if L.ndim != 2:
        raise ValueError("get_spectral_norm expects a matrix")
    import scipy.sparse  # noqa: F401  (kept lazy: only to reject sparse input explicitly)
    if scipy.sparse.issparse(L):
        raise NotImplementedError("sparse linear operators are outside the B200 hot path")
    ctx = _ffi.context()
    M, K = L.shape
    if K > 128:
        raise NotImplementedError("get_spectral_norm: more than 128 columns")
    dA = ctx.upload(L)
    dS = ctx.upload(np.zeros((K, 4), np.float32))
    try:
        lipA, lipS = C.c_float(0), C.c_float(0)
        _ffi.check(_ffi.lib().pmx_nmf_lipschitz(ctx.handle, dA, dS, M, 4, K, C.byref(lipA), C.byref(lipS)))
    finally:
        ctx.free(dA)
        ctx.free(dS)
    return L.dtype.type(lipS.value) if L.dtype.kind == "f" else lipS.value


def l2sq(x):
    """Sum of squares (utils.py:257-260); host helper for user callbacks, not used on the hot path."""
    return (x ** 2).sum()


def l2(x):
    return np.sqrt((x ** 2).sum())


class ConstantStep(object):
    """``step_f(X, it=None) -> value``.  A plain callable for any solver; the device ADMM loop recognises
    it and keeps the whole iteration on the GPU (an arbitrary Python step function forces one host
    round trip per iteration)."""

    def __init__(self, value):
        self.value = value

    def __call__(self, *X, it=None):
        return self.value


class LeastSquaresProx(object):
    """``prox_f(X, step) = X - step * (X - b)``: the gradient-step prox of f = 0.5 |X - b|^2
    (README.md:82-84 pattern).  Recognised by the device ADMM/SDMM loop."""

    def __init__(self, b):
        self.b = b

    def __call__(self, X, step):
        return X - step * (X - self.b)
