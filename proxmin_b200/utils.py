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


def get_step_g(step_f, norm_L2, N=1, M=1):
    """step_g compatible with step_f and ||L||^2 for ADMM / SDMM / bSDMM (utils.py:269-279): scalar host logic."""
    return step_f * norm_L2 * N * M


def get_step_f(step_f, lR2, lS2):
    """Residual balancing of Boyd (2011) section 3.4.1 (utils.py:282-292; no caller in the reference)."""
    mu, tau = 10, 2
    if lR2 > mu * lS2:
        return step_f * tau
    if lS2 > mu * lR2:
        return step_f / tau
    return step_f


def hasNotNone(l):
    """utils.py:409-418 (no caller in the reference): number of entries from the first one that holds a non-None
    item onwards."""
    for i, ll in enumerate(l):
        if ll is not None and hasattr(ll, "__iter__"):
            if any(lll is not None for lll in ll):
                return len(l) - i
    return 0


class ApproximateCache(object):
    """Strided memoisation of a slow step function (utils.py:124-190): pure host control logic around a user
    callable.  ``slack`` = relative change that triggers a re-evaluation, ``max_stride`` = longest skip."""

    def __init__(self, func, slack=0.1, max_stride=100):
        assert slack >= 0 and slack < 1
        self.func, self.slack, self.max_stride = func, slack, max_stride
        self.it, self.stride, self.last, self.stored = 0, 1, -1, None

    def __len__(self):
        return len(self.stride)   # (the reference's __len__ fails the same way: utils.py:163)

    def __call__(self, *args, **kwargs):
        if self.slack == 0:
            self.it += 1
            return self.func(*args, **kwargs)
        if self.it >= self.last + self.stride:
            self.last = self.it
            val = self.func(*args, **kwargs)
            if self.it > 1 and self.slack > 0:
                rel_error = np.abs(self.stored - val) / self.stored
                budget = self.slack / 2
                if rel_error < budget and rel_error > 0:
                    self.stride += max(1, int(budget / rel_error * self.stride))
                    self.stride = min(self.max_stride, self.stride)
            self.stored = val
        else:
            self.it += 1
        return self.stored


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
