"""Proximal solvers with the signatures of proxmin/algorithms.py, executed on a B200.

Two execution modes per solver:

* **fused device loop** -- when ``grad``/``step``/``prox`` are recognised library callables
  (``partial(nmf.grad_likelihood, Y=...)``, ``nmf.step_pgm``/``nmf.step_adaprox``, built-in
  ``prox_*`` operators, ``AlternatingProjections`` of them, ``utils.LeastSquaresProx`` /
  ``utils.ConstantStep``) the whole iteration loop runs on the GPU through ``pmx_nmf_*`` /
  ``pmx_admm_*``; the host only polls a stop flag.
* **callback loop** -- arbitrary user callables are honoured exactly as in the reference: they are
  called on host arrays every iteration, while the library's own arithmetic (forward step,
  built-in prox, moment updates, ADMM variable updates, norms) still runs as CUDA kernels.
  This is the user's code on the host, not a CPU fallback of the library.

Arithmetic is fp32 on the device; fp64 inputs are converted on upload and results cast back.
"""
import ctypes as C
import logging
from functools import partial

import numpy as np

from . import _ffi
from . import operators
from . import utils

logger = logging.getLogger("proxmin")


# ------------------------------------------------------------------------------------------
# recognition of library callables
# ------------------------------------------------------------------------------------------
def _nmf_grad_target(grad):
    """Y if ``grad`` is partial(nmf.grad_likelihood, Y=Y[, W=1]); else None."""
    from . import nmf as _nmf

    if isinstance(grad, partial) and grad.func is _nmf.grad_likelihood and not grad.args:
        kw = grad.keywords or {}
        W = kw.get("W", 1)
        if "Y" in kw and set(kw) <= {"Y", "W"} and np.ndim(W) == 0 and W == 1 and np.ndim(kw["Y"]) == 2:
            return kw["Y"]
    return None


def _is_step(step, fn_name):
    from . import nmf as _nmf

    target = getattr(_nmf, fn_name)
    if step is target:
        return True
    if isinstance(step, partial) and step.func is target and not step.args:
        kw = step.keywords or {}
        W = kw.get("W", 1)
        return set(kw) <= {"W"} and np.ndim(W) == 0 and W == 1
    return False


def _describe_all(prox, allow_none=False):
    out = []
    for p in prox:
        if p is None and allow_none:
            out.append(None)
            continue
        d = operators.describe(p)
        if d is None or len(d) > _ffi.PMX_MAX_OPS:
            return None
        axes = {a for (o, _, a, _) in d if o == _ffi.OP_UNITY}
        if len(axes) > 1:
            return None
        out.append(d)
    return out


def _is_factor_pair(X):
    return (len(X) == 2 and all(isinstance(x, np.ndarray) and x.ndim == 2 for x in X)
            and X[0].shape[1] == X[1].shape[0])


def _writeback(prob, X):
    prob.get(_ffi.A, out=X[0])
    prob.get(_ffi.S, out=X[1])


# ------------------------------------------------------------------------------------------
# device helpers for the callback loops
# ------------------------------------------------------------------------------------------
def _dev_pgm_update(ops, Xe, G, X, step):
    """X[:] = prox(Xe - step*G) with a built-in chain ``ops``; returns (|X-Xold|^2, |X|^2)."""
    ctx = _ffi.context()
    x32 = np.ascontiguousarray(X, dtype=np.float32)
    rows, cols = (1, x32.size) if x32.ndim != 2 else x32.shape
    if x32.ndim != 2:
        ops = [(o, r, 1 if o == _ffi.OP_UNITY else a, t) for (o, r, a, t) in ops]
    dXe, dG, dX = ctx.upload(Xe), ctx.upload(np.broadcast_to(np.asarray(G, dtype=np.float32), x32.shape)), ctx.upload(x32)
    try:
        nd, nn = C.c_double(0), C.c_double(0)
        prox = _ffi.make_prox(ops)
        _ffi.check(_ffi.lib().pmx_pgm_update(ctx.handle, C.byref(prox), dXe, dG, dX, rows, cols, float(step),
                                             C.byref(nd), C.byref(nn)))
        ctx.d2h(x32, dX)
    finally:
        for p in (dXe, dG, dX):
            ctx.free(p)
    X[...] = x32.reshape(X.shape)
    return nd.value, nn.value


def _scalar_step(s):
    if np.ndim(s) != 0:
        raise NotImplementedError("array-valued step sizes are not supported by the callback loop")
    return float(s)


# ------------------------------------------------------------------------------------------
# PGM
# ------------------------------------------------------------------------------------------
def pgm(
    X,
    grad,
    step,
    prox=None,
    accelerated=False,
    backtracking=False,
    f=None,
    e_rel=1e-6,
    max_iter=1000,
    callback=None,
):
    """Proximal Gradient Method / FISTA / block-simultaneous PGM (algorithms.py:12-144).

    Returns ``(converged, gradient, step)`` of the last iteration; X is updated in place."""
    X = utils._as_tuple(X)
    N = len(X)
    prox = utils._as_tuple(prox)
    if len(prox) == 1:
        prox = prox * N
    assert len(prox) == len(X)
    prox = tuple(p if p is not None else operators.prox_id for p in prox)

    if np.isscalar(e_rel):
        e_rel = (e_rel,) * N
    assert len(e_rel) == len(X)
    assert backtracking is False or f is not None

    Y = _nmf_grad_target(grad)
    chains = _describe_all(prox)
    if (Y is not None and chains is not None and _is_factor_pair(X) and _is_step(step, "step_pgm")
            and not backtracking):
        return _pgm_nmf_device(X, Y, chains, accelerated, e_rel, max_iter, callback)
    return _pgm_callbacks(X, grad, step, prox, accelerated, backtracking, f, e_rel, max_iter, callback)


def _pgm_nmf_device(X, Y, chains, accelerated, e_rel, max_iter, callback):
    from . import nmf as _nmf

    A, S = X
    prob = _nmf.Problem(Y, A, S)
    try:
        prob.pgm_begin(chains[0], chains[1], accelerated=accelerated, e_rel=e_rel)
        converged = (False, False)
        steps = (np.float32(np.nan), np.float32(np.nan))
        done = 0
        if callback is None:
            done, converged, steps = prob.pgm_run(max_iter)
        else:
            for it in range(max_iter):
                try:
                    callback(*X, it=it)
                except StopIteration:
                    break
                prob.set(_ffi.A, A)  # the callback may have touched the factors (in-place contract)
                prob.set(_ffi.S, S)
                n, converged, steps = prob.pgm_run(1)
                done += n
                _writeback(prob, X)
                if all(converged):
                    break
        _writeback(prob, X)
        dt = np.result_type(A.dtype, S.dtype)
        G = (prob.get(_ffi.GA, dtype=dt), prob.get(_ffi.GS, dtype=dt))
    finally:
        prob.close()
    logger.info("Completed {0} iterations".format(done))
    if not all(converged):
        logger.warning("Solution did not converge")
    return tuple(np.bool_(c) for c in converged), G, tuple(dt.type(s) for s in steps)


def _pgm_callbacks(X, grad, step, prox, accelerated, backtracking, f, e_rel, max_iter, callback):
    N = len(X)
    try:  # algorithms.py:73-77: the probe really calls the step function once
        step(*X, it=0, grads=X)
        _step = step
    except TypeError:
        _step = lambda *X, it=None, grads=None: step(*X, it=it)  # noqa: E731

    if callback is None:
        callback = utils.NullCallback()
    chains = [operators.describe(p) for p in prox]
    accel = utils.NesterovAccelerator(accelerated=accelerated)
    T = [1.0] * N
    converged = (False,) * N
    G = S = None
    it = -1
    for it in range(max_iter):
        try:
            callback(*X, it=it)
            omega = accel.omega
            if omega > 0:
                _X = tuple(X[j] + omega * (X[j] - X_[j]) for j in range(N))  # noqa: F821 (host glue on user arrays)
            elif backtracking:
                _X = utils._copy_tuple(X)
            else:
                _X = X
            X_ = utils._copy_tuple(X)
            G = utils._as_tuple(grad(*_X))
            S = utils._as_tuple(_step(*_X, it=it, grads=G))
            norms = [None] * N
            for j in range(N):
                norms[j] = _update_block(chains[j], prox[j], _X[j], G[j], X[j], X_[j], T[j] * S[j])

            if backtracking:  # Beck & Teboulle eq. 3.2 (algorithms.py:110-127); f is the user's function
                f_now = f(*X)
                if it == 0:
                    f_prev = f(*X_)
                while f_now > f_prev + np.sum(
                    [np.sum((X[j] - X_[j]) * G[j]) + 0.5 / (T[j] * S[j]) * np.sum((X[j] - X_[j]) ** 2)
                     for j in range(N)]):
                    jmax = np.argmax([np.max(np.abs(S[j] * G[j])) / np.max(np.abs(X_[j])) for j in range(N)])
                    T[jmax] /= 2
                    norms[jmax] = _update_block(chains[jmax], prox[jmax], _X[jmax], G[jmax], X[jmax], X_[jmax],
                                                T[jmax] * S[jmax])
                    f_now = f(*X)
                f_prev = f_now

            converged = tuple(np.float32(norms[j][0]) <= np.float32(e_rel[j] ** 2) * np.float32(norms[j][1])
                              for j in range(N))
            if all(converged):
                break
        except StopIteration:
            break

    logger.info("Completed {0} iterations".format(it + 1))
    if not all(converged):
        logger.warning("Solution did not converge")
    return converged, G, S


def _update_block(chain, prox, Xe, G, X, Xold, step):
    """One block of algorithms.py:107-108 + the norms of :130-133, forward step on the device."""
    s = _scalar_step(step)
    if chain is not None:
        return _dev_pgm_update(chain, Xe, G, X, s)
    # user prox: forward step on the device, the user's callable on the host, norms on the device
    V = np.array(Xe, dtype=X.dtype, copy=True)
    _dev_pgm_update([], Xe, G, V, s)
    X[:] = prox(V, step)
    return _dev_diff_norms(X, Xold)


def _dev_diff_norms(X, Xold):
    """(|X - Xold|^2, |X|^2) on the device: a forward step with zero gradient scale."""
    tmp = np.array(Xold, dtype=np.float32, copy=True)
    nd, nn = _dev_pgm_update([], X, np.zeros_like(tmp), tmp, 0.0)
    return nd, nn


# ------------------------------------------------------------------------------------------
# placeholders filled in below
# ------------------------------------------------------------------------------------------
def adaprox(X, grad, step, prox=None, scheme="adam", b1=0.9, b2=0.999, eps=1e-8, check_convergence=True,
            p=0.25, e_rel=1e-6, max_iter=1000, prox_max_iter=1000, M=None, V=None, Vhat=None, callback=None):
    raise NotImplementedError


def admm(X, prox_f, step_f, prox_g=None, step_g=None, L=None, e_rel=1e-6, e_abs=0, max_iter=1000, callback=None):
    raise NotImplementedError


def sdmm(X, prox_f, step_f, proxs_g=None, steps_g=None, Ls=None, e_rel=1e-6, e_abs=0, max_iter=1000,
         callback=None):
    raise NotImplementedError


def bsdmm(X, proxs_f, steps_f_cb, proxs_g=None, steps_g=None, Ls=None, update_order=None,
          steps_g_update="steps_f", max_iter=1000, e_rel=1e-6, e_abs=0, callback=None):
    raise NotImplementedError


def _bsdmm_nmf(Y, A, S, W, prox_A, prox_S, max_iter, e_rel, callback, **kw):
    raise NotImplementedError
