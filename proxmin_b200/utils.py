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
    """Squared spectral norm of a dense matrix = lambda_max(L^T L) (utils.py:14-35, dense branch), on the device:
    Gram kernel + one-CTA eigen-solver on the smaller of L^T L / L L^T when that is at most 128 x 128, power iteration
    with device matrix-vector products otherwise."""
    if L is None:
        return 1
    if isinstance(L, MatrixAdapter):
        if L.L is None:
            return 1
        if L._spec_norm is not None:
            return L._spec_norm
        adapter, L = L, L.L
    else:
        adapter = None
    import scipy.sparse
    if scipy.sparse.issparse(L):
        raise NotImplementedError("sparse linear operators are outside the B200 hot path")
    L = np.asarray(L)
    if L.ndim != 2:
        raise ValueError("get_spectral_norm expects a matrix")
    dt = L.dtype.type if L.dtype.kind == "f" else np.float64
    Ls = L if L.shape[1] <= L.shape[0] else L.T       # the nonzero spectra of L^T L and L L^T agree
    M, K = Ls.shape
    if K <= 128:
        ctx = _ffi.context()
        dA = ctx.upload(np.ascontiguousarray(Ls, dtype=np.float32))
        dS = ctx.upload(np.zeros((K, 4), np.float32))
        try:
            lipA, lipS = C.c_float(0), C.c_float(0)
            _ffi.check(_ffi.lib().pmx_nmf_lipschitz(ctx.handle, dA, dS, M, 4, K, C.byref(lipA), C.byref(lipS)))
        finally:
            ctx.free(dA)
            ctx.free(dS)
        return dt(lipS.value)
    from . import _dev

    dm = adapter._device() if adapter is not None else _dev.DeviceMatrix(L)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(dm.shape[1]).astype(np.float32)
    v /= np.sqrt(_dev.sumsq(v))
    lam = 0.0
    for it in range(2000):
        w = dm.Tdot(dm.dot(v))
        nw = np.sqrt(_dev.sumsq(w))
        if not np.isfinite(nw):
            raise np.linalg.LinAlgError("Array must not contain infs or NaNs")
        if nw == 0:
            lam = 0.0
            break
        new = nw                       # |L^T L v| -> lambda_max for a unit vector v
        v = (w / np.float32(nw)).astype(np.float32)
        if it > 4 and abs(new - lam) <= 1e-7 * new:
            lam = new
            break
        lam = new
    if adapter is None:
        dm.close()
    return dt(lam)


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


class BarzilaiBorweinStepper:
    """Barzilai-Borwein step sizes, stabilised after Burdakov et al. (arXiv:1907.06409, Algorithm 2.1)
    (utils.py:209-241).  ``stepper.step`` is a ``step(*X, it=None, grads=None)`` callable for ``pgm``; every
    reduction (sum S^2, sum S Y, sum Y^2, sum G^2, max|X|, max|G|) is one device kernel per block
    (``pmx_ew`` opcodes BB / MAXABS), the remaining arithmetic is a handful of host scalars."""

    def __init__(self, type=1, init_r=0.1):
        assert type in [1, 2]
        self.r = init_r
        self.type = type

    def step(self, *X, it=None, grads=None):
        from . import _dev

        N = len(X)
        if it == 0:
            self.Delta = np.array([np.inf, ] * N)
            self.X_ = _copy_tuple(X)
            self.G_ = grads  # no copy needed, created fresh every single iteration
            return tuple(self.r * _dev.maxabs(X[j]) / _dev.maxabs(grads[j]) for j in range(N))

        G = grads
        red = [_dev.bb_sums(X[j], self.X_[j], G[j], self.G_[j]) for j in range(N)]   # (S.S, S.Y, Y.Y, G.G)
        dt = [np.result_type(X[j].dtype, np.float32).type for j in range(N)]
        self.X_ = _copy_tuple(X)
        self.G_ = grads

        with np.errstate(divide="ignore", invalid="ignore"):
            if self.type == 1:
                A = tuple(dt[j](red[j][0]) / dt[j](red[j][1]) for j in range(N))
            else:
                A = tuple(dt[j](red[j][1]) / dt[j](red[j][2]) for j in range(N))
            if it <= 3:
                self.Delta = np.minimum(self.Delta, tuple(np.sqrt(dt[j](red[j][0])) for j in range(N)))
            Astab = tuple(self.Delta[j] / np.sqrt(dt[j](red[j][3])) for j in range(N))
        return np.minimum(np.abs(A), Astab)


class MatrixAdapter(object):
    """Linear operator wrapper that tolerates ``None`` (= identity) and caches the spectral norm (utils.py:38-101).
    A dense matrix is uploaded once and stays resident on the device; ``dot`` is one device GEMM per call
    (``pmx_matmul``), ``.T`` shares the device copy.  With ``L is None`` the argument of ``dot`` is returned uncopied,
    as in the reference (utils.py:70-74).  scipy.sparse operators are outside the B200 hot path."""

    def __init__(self, L, axis=None, _dm=None, _trans=False):
        spec_norm = None
        while isinstance(L, MatrixAdapter):   # prevent cascade
            spec_norm = L._spec_norm
            axis = L.axis
            _dm, _trans = L._dm, L._trans
            L = L.L
        if L is not None:
            import scipy.sparse
            if scipy.sparse.issparse(L):
                raise NotImplementedError("sparse linear operators are outside the B200 hot path")
            L = np.asarray(L)
        self.L = L
        self.axis = axis
        self._spec_norm = spec_norm
        self._dm, self._trans = _dm, _trans

    def _device(self):
        from . import _dev

        if self._dm is None:
            self._dm = _dev.DeviceMatrix(self.L)
            self._trans = False
        return self._dm

    @property
    def spectral_norm(self):
        if self._spec_norm is None:
            if self.L is not None:
                self._spec_norm = get_spectral_norm(self)
            else:
                self._spec_norm = 1
        return self._spec_norm

    @property
    def T(self):
        if self.L is None:
            return self  # NOT: self.L !!!
        # because we need to preserve axis for dot(), create a new adapter (it shares the device copy)
        self._device()
        out = MatrixAdapter(self.L.T, axis=self.axis, _dm=self._dm, _trans=not self._trans)
        out._spec_norm = self._spec_norm   # lambda_max(L^T L) = lambda_max(L L^T)
        return out

    def _mul(self, X):
        dm = self._device()
        return dm.Tdot(X) if self._trans else dm.dot(X)

    def dot(self, X):
        if self.L is None:
            # CAVEAT (reference): not a copy
            return X
        if self.axis is None:
            return self._mul(X)
        if self.axis == 1:
            return self._mul(X.reshape(-1)).reshape(X.shape[0], -1)
        raise NotImplementedError(
            "MatrixAdapter.dot() is not useful with axis=0.\n"
            "Use regular matrix dot product instead!"
        )

    def __len__(self):
        return len(self.L)

    @property
    def shape(self):
        return self.L.shape

    @property
    def size(self):
        return self.L.size

    @property
    def ndim(self):
        return self.L.ndim


def initZU(X, L):
    """Z = L X, U = 0 per constraint (utils.py:244-254)."""
    if not isinstance(L, list):
        Z = L.dot(X).copy()
        U = np.zeros(Z.shape, dtype=Z.dtype)
    else:
        Z, U = [], []
        for i in range(len(L)):
            Z.append(L[i].dot(X).copy())
            U.append(np.zeros(Z[i].shape, dtype=Z[i].dtype))
    return Z, U


def do_the_mm(X, step_f, Z, U, prox_g, step_g, L):
    """utils.py:295-304 on the device (Z, U updated in place); returns (LX, R, S).  ``L``: a MatrixAdapter."""
    from . import algorithms as _alg

    LX, R, S, _ = _alg._mm(X, Z, U, prox_g, _alg._device_chain(prox_g), step_g, True, L=_alg._linop(L))
    return LX, R, S


def update_variables(X, Z, U, prox_f, step_f, prox_g, step_g, L):
    """utils.py:307-346: one ADMM / SDMM variable update (X, Z, U in place); returns (LX, R, S)."""
    from . import algorithms as _alg

    if hasattr(prox_g, "__iter__"):
        chains = [_alg._device_chain(pg) for pg in prox_g]
        Ls = [_alg._linop(l) for l in L] if isinstance(L, list) else _alg._linop(L)
    else:
        chains = _alg._device_chain(prox_g) if prox_g is not None else None
        Ls = _alg._linop(L)
    LX, R, S, _ = _alg._update_variables(X, Z, U, prox_f, step_f, prox_g, chains, step_g, True, L=Ls)
    return LX, R, S


def get_variable_errors(X, L, LX, Z, U, step_g, e_rel, e_abs=0):
    """utils.py:349-363: tolerances (e_pri, e_dual) of one constraint, norms on the device."""
    from . import _dev
    from . import algorithms as _alg

    Lad = _alg._linop(L)
    spec = _alg._spec(Lad)
    e_pri2 = np.sqrt(Z.size) * e_abs / spec + e_rel * np.max([np.sqrt(np.float32(_dev.sumsq(LX))),
                                                             np.sqrt(np.float32(_dev.sumsq(Z)))])
    LTU = U if Lad is None else Lad.T.dot(U)
    lU = np.sqrt(np.float32(_dev.sumsq(LTU)))
    if step_g is not None:
        lU = lU / np.float32(step_g)
    e_dual2 = np.sqrt(X.size) * e_abs / spec + e_rel * lU
    return e_pri2, e_dual2


def check_constraint_convergence(X, L, LX, Z, U, R, S, step_f, step_g, e_rel, e_abs):
    """utils.py:366-391: Boyd (2011) section 3.3.1 stopping rule, recursive over lists of constraints."""
    from . import _dev

    if isinstance(L, list):
        convergence, errors = True, []
        for i in range(len(L)):
            c, e = check_constraint_convergence(X, L[i], LX[i], Z[i], U[i], R[i], S[i], step_f, step_g[i], e_rel, e_abs)
            convergence &= c
            errors.append(e)
        return convergence, errors
    e_pri, e_dual = get_variable_errors(X, L, LX, Z, U, step_g, e_rel, e_abs)
    lR = np.sqrt(np.float32(_dev.sumsq(R)))
    lS = np.sqrt(np.float32(_dev.sumsq(S)))
    return (lR <= e_pri) and (lS <= e_dual), (e_pri, e_dual, lR, lS)


def check_convergence(newX, oldX, e_rel):
    """utils.py:394-406 (Langville 2014, section 5): sum(new * old) >= (1 - e_rel^2) sum(old^2), sums on the device."""
    from . import _dev

    dot, _ = _dev.dot_diff(newX, np.zeros_like(newX), oldX)      # sum((new - 0) * old)
    norms = [newX.dtype.type(dot) if newX.dtype.kind == "f" else dot, oldX.dtype.type(_dev.sumsq(oldX))
             if oldX.dtype.kind == "f" else _dev.sumsq(oldX)]
    convergent = norms[0] >= (1 - e_rel ** 2) * norms[1]
    return convergent, norms


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
