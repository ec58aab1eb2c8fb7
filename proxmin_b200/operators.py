"""Proximal operators with the call contract of proxmin/operators.py: ``prox(X, step, **kw) -> X``
(in place, returns its argument).  Every operator runs as a CUDA kernel through the C ABI
(``pmx_prox_apply``); there is no NumPy implementation behind them.

The solvers never call these wrappers on the hot path: they translate a callable into a
primitive-op chain with :func:`describe` (by identity, through ``functools.partial`` and inside
``AlternatingProjections``, the idiom of operators.py:213-224) and hand the chain to the fused
device kernels.  Calling an operator directly round-trips the array through the GPU.
"""
import ctypes as C
import functools

import numpy as np

from . import _ffi


def _apply(X, step, ops):
    """Run a primitive-op chain on a host array in place."""
    if not isinstance(X, np.ndarray):
        raise TypeError("proximal operators expect a numpy.ndarray")
    if X.size == 0:
        return X
    if X.ndim <= 1:
        rows, cols = 1, X.size
        ops = [(o, r, 1 if o == _ffi.OP_UNITY else a, t) for (o, r, a, t) in ops]  # 1-D: the only axis
    elif X.ndim == 2:
        rows, cols = X.shape
    else:
        if any(o == _ffi.OP_UNITY for (o, _, _, _) in ops):
            raise NotImplementedError("prox_unity on arrays with more than 2 dimensions")
        rows, cols = 1, X.size
    # thresholds of type="relative" follow NumPy's scalar arithmetic of the reference (operators.py:4-14)
    res = []
    for (o, rel, a, t) in ops:
        if rel and o in (_ffi.OP_MIN, _ffi.OP_MAX, _ffi.OP_HARD, _ffi.OP_SOFT, _ffi.OP_MAXENT, _ffi.OP_MAXENT64):
            if np.ndim(step) != 0:
                raise NotImplementedError("array-valued step with a relative threshold")
            t = np.float32(t * step)
            rel = 0
        res.append((o, rel, a, float(t)))
    ctx = _ffi.context()
    x32 = np.ascontiguousarray(X, dtype=np.float32)
    d = ctx.upload(x32)
    try:
        prox = _ffi.make_prox(res)
        st = float(step) if np.ndim(step) == 0 else 0.0
        _ffi.check(_ffi.lib().pmx_prox_apply(ctx.handle, C.byref(prox), d, rows, cols, st))
        ctx.d2h(x32, d)
    finally:
        ctx.free(d)
    if x32 is not X:
        X[...] = x32.reshape(X.shape)
    return X


def _step_gamma(step, gamma):
    """gamma * step: the parameter of a continuous penalty in units of the step (operators.py:4-14)."""
    return gamma * step


def _rel(type):
    assert type in ["relative", "absolute"]
    return type == "relative"


def prox_id(X, step):
    """Identity (operators.py:20-23)."""
    return X


def prox_zero(X, step):
    """Projection onto zero (operators.py:26-30)."""
    return _apply(X, step, [(_ffi.OP_ZERO, 0, 0, 0.0)])


def prox_plus(X, step):
    """Projection onto non-negative numbers (operators.py:33-38)."""
    return _apply(X, step, [(_ffi.OP_PLUS, 0, 0, 0.0)])


def prox_unity(X, step, axis=0):
    """Divide by the sum along ``axis`` (operators.py:41-45)."""
    return _apply(X, step, [(_ffi.OP_UNITY, 0, axis, 0.0)])


def prox_unity_plus(X, step, axis=0):
    """prox_plus, then prox_unity (operators.py:48-52)."""
    return _apply(X, step, [(_ffi.OP_PLUS, 0, 0, 0.0), (_ffi.OP_UNITY, 0, axis, 0.0)])


def prox_min(X, step, thresh=0, type="relative"):
    """Projection onto numbers above ``thresh`` (operators.py:55-69)."""
    return _apply(X, step, [(_ffi.OP_MIN, _rel(type), 0, thresh)])


def prox_max(X, step, thresh=0, type="relative"):
    """Projection onto numbers below ``thresh`` (operators.py:72-84)."""
    return _apply(X, step, [(_ffi.OP_MAX, _rel(type), 0, thresh)])


def prox_hard(X, step, thresh=0, type="relative"):
    """Hard thresholding (operators.py:109-125)."""
    return _apply(X, step, [(_ffi.OP_HARD, _rel(type), 0, thresh)])


def prox_hard_plus(X, step, thresh=0, type="relative"):
    """Hard thresholding, then projection onto non-negative numbers (operators.py:128-135)."""
    return _apply(X, step, [(_ffi.OP_HARD, _rel(type), 0, thresh), (_ffi.OP_PLUS, 0, 0, 0.0)])


def prox_soft(X, step, thresh=0, type="relative"):
    """Soft thresholding (operators.py:138-150)."""
    return _apply(X, step, [(_ffi.OP_SOFT, _rel(type), 0, thresh)])


def prox_soft_plus(X, step, thresh=0, type="relative"):
    """Soft thresholding, then projection onto non-negative numbers (operators.py:153-160)."""
    return _apply(X, step, [(_ffi.OP_SOFT, _rel(type), 0, thresh), (_ffi.OP_PLUS, 0, 0, 0.0)])


def prox_max_entropy(X, step, gamma=1, type="relative"):
    """Proximal operator of g(x) = gamma sum_i x_i ln(x_i) (operators.py:163-184):
    X[X > 0] = gamma_ W(exp(X / gamma_ - 1) / gamma_) with the Lambert W function evaluated on the device
    (Halley iterations in fp64, like scipy's lambertw it replaces); elements <= 0 are left untouched."""
    op = _ffi.OP_MAXENT64 if isinstance(X, np.ndarray) and X.dtype == np.float64 else _ffi.OP_MAXENT
    return _apply(X, step, [(op, _rel(type), 0, gamma)])


def prox_components(X, step, prox=None, axis=0):
    """operators.py:87-106.  Dead code in the reference: its body reads the undefined name ``prox_list`` and raises
    NameError on every call; the same exception is raised here."""
    raise NameError("name 'prox_list' is not defined")


class AlternatingProjections(object):
    """Sequential composition of proximal operators (operators.py:187-224): the list is applied in
    reverse order, ``repeat`` times.  When every member is a built-in, the whole composition is one
    fused device chain; otherwise the members are called one after the other."""

    def __init__(self, prox_list=None, repeat=1):
        self.operators = []
        self.repeat = repeat
        if prox_list is not None:
            self.operators += prox_list

    def __call__(self, X, step):
        ops = describe(self)
        if ops is not None and len(ops) <= _ffi.PMX_MAX_OPS:
            return _apply(X, step, ops)
        for r in range(self.repeat):
            for prox in self.operators[::-1]:
                X = prox(X, step)
        return X

    def find(self, cls):
        for i, prox in enumerate(self.operators):
            if isinstance(prox, functools.partial):
                if prox.func is cls:
                    return i
            elif prox is cls:
                return i
        return -1


_THRESHOLDED = {}


def _builtin_table():
    if not _THRESHOLDED:
        _THRESHOLDED.update({
            prox_min: [_ffi.OP_MIN], prox_max: [_ffi.OP_MAX], prox_hard: [_ffi.OP_HARD], prox_soft: [_ffi.OP_SOFT],
            prox_hard_plus: [_ffi.OP_HARD, _ffi.OP_PLUS], prox_soft_plus: [_ffi.OP_SOFT, _ffi.OP_PLUS]})
    return _THRESHOLDED


def describe(prox):
    """Translate a callable into a list of primitive ops ``(op, relative, axis, thresh)`` in application
    order, or ``None`` when it is not (entirely) made of the built-ins above."""
    if prox is None or prox is prox_id:
        return []
    kw = {}
    fn = prox
    if isinstance(prox, functools.partial):
        if prox.args:
            return None
        fn, kw = prox.func, dict(prox.keywords or {})
        if isinstance(fn, functools.partial) or isinstance(fn, AlternatingProjections):
            return None
    if isinstance(fn, AlternatingProjections):
        out = []
        for _ in range(fn.repeat):
            for p in fn.operators[::-1]:
                d = describe(p)
                if d is None:
                    return None
                out += d
        return out
    if fn is prox_id and not kw:
        return []
    if fn is prox_zero and not kw:
        return [(_ffi.OP_ZERO, 0, 0, 0.0)]
    if fn is prox_plus and not kw:
        return [(_ffi.OP_PLUS, 0, 0, 0.0)]
    if fn in (prox_unity, prox_unity_plus):
        axis = kw.pop("axis", 0)
        if kw or axis not in (0, 1):
            return None
        pre = [(_ffi.OP_PLUS, 0, 0, 0.0)] if fn is prox_unity_plus else []
        return pre + [(_ffi.OP_UNITY, 0, axis, 0.0)]
    if fn is prox_max_entropy:
        gamma = kw.pop("gamma", 1)
        type_ = kw.pop("type", "relative")
        if kw or type_ not in ("relative", "absolute") or np.ndim(gamma) != 0:
            return None
        return [(_ffi.OP_MAXENT, type_ == "relative", 0, float(gamma))]
    table = _builtin_table()
    if fn in table:
        thresh = kw.pop("thresh", 0)
        type_ = kw.pop("type", "relative")
        if kw or type_ not in ("relative", "absolute") or np.ndim(thresh) != 0:
            return None
        out = []
        for op in table[fn]:
            if op == _ffi.OP_PLUS:
                out.append((op, 0, 0, 0.0))
            else:
                out.append((op, type_ == "relative", 0, float(thresh)))
        return out
    return None
