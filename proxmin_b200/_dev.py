"""Host-array front end of the elementwise device primitives (``pmx_ew`` & friends).

Used by the *callback loops* of algorithms.py, where iterates must live in host arrays because the
user's grad/step/prox callables are Python functions.  Each helper uploads its operands, runs one
CUDA kernel through the C ABI and downloads the result: simple, and every arithmetic expression of
the library stays on the GPU (there is no NumPy implementation behind these).
"""
import ctypes as C

import numpy as np

from . import _ffi


def _f32(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float32)
    if shape is not None and a.shape != shape:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


class _Scope:
    """Uploads on demand, frees everything on exit."""

    def __init__(self):
        self.ctx = _ffi.context()
        self.ptrs = []

    def up(self, arr):
        p = self.ctx.upload(arr)
        self.ptrs.append(p)
        return p

    def alloc(self, nbytes):
        p = self.ctx.malloc(nbytes)
        self.ptrs.append(p)
        return p

    def down(self, p, shape):
        out = np.empty(shape, np.float32)
        self.ctx.d2h(out, p)
        return out

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        for p in self.ptrs:
            self.ctx.free(p)
        self.ptrs = []
        return False


def ew(op, a, b=None, c=None, d=None, s0=0.0, s1=0.0, n_out=1, reduce=False):
    """Run one ``pmx_ew`` opcode on host arrays; returns (list of outputs, reductions or None)."""
    a32 = _f32(a)
    shape, n = a32.shape, a32.size
    with _Scope() as sc:
        pa = sc.up(a32)
        pb = sc.up(_f32(b, shape)) if b is not None else None
        pc = sc.up(_f32(c, shape)) if c is not None else None
        pd = sc.up(_f32(d, shape)) if d is not None else None
        outs = [sc.alloc(4 * max(n, 1)) for _ in range(n_out)]
        o = outs + [None] * (3 - len(outs))
        red = (C.c_double * 5)() if reduce else None
        _ffi.check(_ffi.lib().pmx_ew(sc.ctx.handle, op, n, pa, pb, pc, pd, float(s0), float(s1), o[0], o[1], o[2], red))
        res = [sc.down(p, shape) for p in outs]
    return res, (list(red) if reduce else None)


def extrapolate(X, Xold, omega):
    """X + omega (X - Xold)   (algorithms.py:95)"""
    return ew(_ffi.EW_EXTRAP, X, Xold, s0=omega)[0][0].astype(X.dtype, copy=False)


def sumsq(x):
    return ew(_ffi.EW_SUMSQ, x, n_out=0, reduce=True)[1][0]


def dot_diff(X, Xold, G):
    """(sum((X - Xold) * G), sum((X - Xold)**2))   (algorithms.py:118)"""
    r = ew(_ffi.EW_DOT_DIFF, X, Xold, G, n_out=0, reduce=True)[1]
    return r[0], r[1]


def maxabs(x, scale=1.0):
    """max(abs(scale * x))   (algorithms.py:121)"""
    return ew(_ffi.EW_MAXABS, x, s0=scale, n_out=0, reduce=True)[1][0]


def bb_sums(X, Xold, G, Gold):
    """(sum(S*S), sum(S*Y), sum(Y*Y), sum(G*G)) with S = X - Xold, Y = G - Gold   (utils.py:225-239)"""
    r = ew(_ffi.EW_BB, X, Xold, G, Gold, n_out=0, reduce=True)[1]
    return r[0], r[1], r[2], r[3]


class DeviceMatrix(object):
    """A dense linear operator kept resident on the device: ``dot(X)`` = L X and ``Tdot(X)`` = L^T X upload only the
    argument (utils.py:76-77; the operators of admm / sdmm / bsdmm, utils.py:299-303, :316, :333)."""

    def __init__(self, L):
        self.L32 = _f32(L)
        assert self.L32.ndim == 2, "linear operator must be a matrix"
        self.shape = self.L32.shape
        self.dtype = np.asarray(L).dtype
        self.ctx = _ffi.context()
        self.ptr = self.ctx.upload(self.L32)

    def close(self):
        if self.ptr is not None:
            self.ctx.free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass

    def _mul(self, X, trans):
        X32 = _f32(X)
        vec = X32.ndim == 1
        Xm = np.ascontiguousarray(X32.reshape(-1, 1) if vec else X32)
        rows, cols = self.shape
        p, n = (cols, rows) if trans else (rows, cols)
        if Xm.shape[0] != n:
            raise ValueError("shapes %s and %s not aligned" % ((p, n), X32.shape))
        m = Xm.shape[1]
        with _Scope() as sc:
            pX = sc.up(Xm)
            pO = sc.alloc(4 * max(p * m, 1))
            _ffi.check(_ffi.lib().pmx_matmul(sc.ctx.handle, self.ptr, pX, pO, p, n, m, 1 if trans else 0))
            out = sc.down(pO, (p, m))
        dt = np.result_type(self.dtype, np.asarray(X).dtype)
        out = out.astype(dt if dt.kind == "f" else np.float64, copy=False)
        return out.reshape(p) if vec else out

    def dot(self, X):
        return self._mul(X, False)

    def Tdot(self, X):
        return self._mul(X, True)


def matmul(L, X):
    """L.dot(X) for a dense matrix L (p x n) and a vector (n,) or matrix (n x m) X on the device."""
    dm = DeviceMatrix(L)
    try:
        return dm.dot(X)
    finally:
        dm.close()


def axpy(s, a, b=None):
    """s * a + b (b optional), two roundings like NumPy."""
    return ew(_ffi.EW_AXPY, a, b, s0=s)[0][0].astype(np.asarray(a).dtype, copy=False)


def admm_xarg(X, Zs, Us, ratios):
    """X - sum_i ratio_i (X - Z_i + U_i)   (utils.py:316-317, 331-338)"""
    dX = None
    for Z, U, r in zip(Zs, Us, ratios):
        dX = ew(_ffi.EW_DX_ACC, X, Z, U, dX, s0=r)[0][0]
    return ew(_ffi.EW_SUB, X, dX)[0][0].astype(X.dtype, copy=False)


def add(a, b):
    return ew(_ffi.EW_ADD, a, b)[0][0].astype(a.dtype, copy=False)


def admm_zu(X, Znew, Z, U, step_g, dual_uses_step_g=True):
    """do_the_mm after prox_g (utils.py:299-303) + the five norms of utils.py:349-363.

    Updates Z and U in place; returns (R, S, norms) with norms = (|X|, |Z'|, |U'(/step_g)|, |R|, |S|)."""
    cS = np.float32(-1 / step_g)
    (R, S, Un), red = ew(_ffi.EW_ZU, X, Znew, Z, U, s0=cS, s1=(np.float32(step_g) if dual_uses_step_g else 0.0),
                         n_out=3, reduce=True)
    Z[...] = np.asarray(Znew).reshape(Z.shape)
    U[...] = Un.reshape(U.shape)
    norms = tuple(np.sqrt(np.float32(v)) for v in red)
    return R.astype(X.dtype, copy=False), S.astype(X.dtype, copy=False), norms


def _alpha_spec(sc, alpha, shape):
    """(device ptr, mode, value) for a scalar / per-column / per-row step (nmf.py:91-93 broadcasting)."""
    if np.ndim(alpha) == 0:
        return None, 0, float(alpha)
    al = np.asarray(alpha, dtype=np.float32)
    if len(shape) == 2:
        rows, cols = shape
        if al.shape in ((cols,), (1, cols)):
            return sc.up(al.reshape(-1)), 2, 0.0
        if al.shape == (rows, 1):
            return sc.up(al.reshape(-1)), 3, 0.0
    raise NotImplementedError("step of shape %s does not broadcast per row/column of %s" % (al.shape, shape))


def adaprox_moments(scheme, G, M, V, Vhat, X, alpha, b1, b1_prev, b2, eps, p, t):
    """Moment update + X -= Alpha*Phi/Psi (algorithms.py:147-245, :378) in place; returns (Psi, max(Psi))."""
    shape = X.shape
    rows, cols = (1, X.size) if X.ndim != 2 else shape
    with _Scope() as sc:
        pG, pM, pV, pX = sc.up(_f32(G, shape)), sc.up(M), sc.up(V), sc.up(X)
        pVh = sc.up(Vhat) if Vhat is not None else None
        pPsi = sc.alloc(4 * max(X.size, 1))
        pal, mode, val = _alpha_spec(sc, alpha, shape if X.ndim == 2 else (1, X.size))
        pm = C.c_float(0)
        _ffi.check(_ffi.lib().pmx_adaprox_moments(sc.ctx.handle, _ffi.SCHEMES[scheme], pG, pM, pV, pVh, pX, pPsi, rows,
                                                  cols, pal, mode, val, float(b1), float(b1_prev), float(b2), float(eps),
                                                  float(p), int(t), C.byref(pm)))
        M[...] = sc.down(pM, shape)
        V[...] = sc.down(pV, shape)
        if Vhat is not None:
            Vhat[...] = sc.down(pVh, shape)
        X[...] = sc.down(pX, shape)
        Psi = sc.down(pPsi, shape)
    return Psi, np.float32(pm.value)


def adaprox_sub(ops, z, X, Psi, alpha, psimax):
    """z' = prox(z - gamma/Alpha * Psi * (z - X), gamma) with a built-in chain (algorithms.py:387);
    returns (z', |z'-z|^2, |z|^2)."""
    shape = z.shape
    rows, cols = (1, z.size) if z.ndim != 2 else shape
    if z.ndim != 2:
        ops = [(o, r, 1 if o == _ffi.OP_UNITY else a, t) for (o, r, a, t) in ops]
    with _Scope() as sc:
        pz, pX, pPsi = sc.up(z), sc.up(X), sc.up(Psi)
        pout = sc.alloc(4 * max(z.size, 1))
        pal, mode, val = _alpha_spec(sc, alpha, shape if z.ndim == 2 else (1, z.size))
        norms = (C.c_double * 3)()
        prox = _ffi.make_prox(ops)
        _ffi.check(_ffi.lib().pmx_adaprox_sub(sc.ctx.handle, C.byref(prox), pz, pX, pPsi, pout, rows, cols, pal, mode,
                                              val, float(psimax), norms))
        out = sc.down(pout, shape)
    return out.astype(z.dtype, copy=False), norms[0], norms[2]
