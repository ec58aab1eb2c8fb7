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

from . import _dev
from . import _ffi
from . import operators
from . import utils

logger = logging.getLogger("proxmin")


# ------------------------------------------------------------------------------------------
# recognition of library callables
# ------------------------------------------------------------------------------------------
def _nmf_grad_target(grad, weighted=False):
    """Y if ``grad`` is partial(nmf.grad_likelihood, Y=Y[, W=1]); else None.  ``weighted=True``: also accept an M x N
    weight matrix and return ``(Y, W)`` (W is None for W == 1)."""
    from . import nmf as _nmf

    if isinstance(grad, partial) and grad.func is _nmf.grad_likelihood and not grad.args:
        kw = grad.keywords or {}
        W = kw.get("W", 1)
        if "Y" in kw and set(kw) <= {"Y", "W"} and np.ndim(kw["Y"]) == 2:
            if np.ndim(W) == 0 and W == 1:
                return (kw["Y"], None) if weighted else kw["Y"]
            if weighted and np.ndim(W) == 2 and np.shape(W) == np.shape(kw["Y"]):
                return kw["Y"], W
    return (None, None) if weighted else None


class _ResidentGradient(object):
    """grad(*X) of the NMF likelihood for the callback loops: Y (and W) stay on the device in a solver handle, every
    call uploads the two factors, runs the fused gradient kernel and downloads the two gradients (nmf.py:28-41)."""

    def __init__(self, Y, W):
        self.Y, self.W, self.prob = Y, W, None

    def __call__(self, *X):
        from . import nmf as _nmf

        A, S = X
        if self.prob is None:
            self.prob = _nmf.Problem(self.Y, A, S, W=self.W)
        else:
            self.prob.set(_ffi.A, A)
            self.prob.set(_ffi.S, S)
        self.prob.gradient()
        dt = np.result_type(A.dtype, S.dtype)
        return self.prob.get(_ffi.GA, dtype=dt), self.prob.get(_ffi.GS, dtype=dt)

    def close(self):
        if self.prob is not None:
            self.prob.close()
            self.prob = None


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


def _device_chain(p):
    """Primitive-op chain of ``p`` if one fused device chain can express it (built-ins only, at most
    PMX_MAX_OPS primitive ops, prox_unity along a single axis); else None -> the callable runs on the host."""
    d = operators.describe(p)
    if d is None or len(d) > _ffi.PMX_MAX_OPS:
        return None
    axes = {a for (o, _, a, _) in d if o == _ffi.OP_UNITY}
    if len(axes) > 1:
        return None
    return d


def _describe_all(prox, allow_none=False):
    out = []
    for p in prox:
        if p is None and allow_none:
            out.append(None)
            continue
        d = _device_chain(p)
        if d is None:
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
def _dev_pgm_update(ops, Xe, G, X, step, Xold=None):
    """X[:] = prox(Xe - step*G) with a built-in chain ``ops``; returns (|X-Xold|^2, |X|^2).

    ``Xold`` is the iterate the convergence norms compare against (algorithms.py:130-133: the copy ``X_`` taken
    before the update); it defaults to the contents of X, which is the same thing except for the re-update of a
    block inside the backtracking line search (algorithms.py:124), where X already holds the first attempt."""
    ctx = _ffi.context()
    x32 = np.array(X if Xold is None else Xold, dtype=np.float32, order="C", copy=True)
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
        if np.size(s) == 1:      # e.g. BarzilaiBorweinStepper: np.minimum(...) of one variable, shape (1,)
            return float(np.asarray(s).reshape(-1)[0])
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
    Yw, Ww = _nmf_grad_target(grad, weighted=True)
    if Yw is not None and _is_factor_pair(X):
        # user step / prox / backtracking with the library's gradient: Y (and W) stay resident on the device
        resident = _ResidentGradient(Yw, Ww)
        try:
            return _pgm_callbacks(X, resident, step, prox, accelerated, backtracking, f, e_rel, max_iter, callback)
        finally:
            resident.close()
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
    chains = [_device_chain(p) for p in prox]
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
                _X = tuple(_dev.extrapolate(X[j], X_[j], omega) for j in range(N))  # noqa: F821
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
                while f_now > f_prev + _quadratic_model(X, X_, G, T, S):
                    jmax = int(np.argmax([_dev.maxabs(G[j], _scalar_step(S[j])) / _dev.maxabs(X_[j]) for j in range(N)]))
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
        return _dev_pgm_update(chain, Xe, G, X, s, Xold=Xold)
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



def _quadratic_model(X, X_, G, T, S):
    """sum_j <X_j - X_j_old, G_j> + |X_j - X_j_old|^2 / (2 T_j S_j)   (Beck & Teboulle eq. 3.2, algorithms.py:118)"""
    tot = 0.0
    for j in range(len(X)):
        dot, sq = _dev.dot_diff(X[j], X_[j], G[j])
        tot += dot + 0.5 / (T[j] * _scalar_step(S[j])) * sq
    return tot


# ------------------------------------------------------------------------------------------
# adaprox
# ------------------------------------------------------------------------------------------
def adaprox(
    X,
    grad,
    step,
    prox=None,
    scheme="adam",
    b1=0.9,
    b2=0.999,
    eps=1e-8,
    check_convergence=True,
    p=0.25,
    e_rel=1e-6,
    max_iter=1000,
    prox_max_iter=1000,
    M=None,
    V=None,
    Vhat=None,
    callback=None,
):
    """Adaptive proximal gradient method: Adam, NAdam, AMSGrad, PAdam, AdamX, RAdam with proximal
    sub-iterations (algorithms.py:248-423).  Returns ``(converged, M, V, Vhat)``; X is updated in place."""
    X = utils._as_tuple(X)
    N = len(X)
    prox = utils._as_tuple(prox)
    if len(prox) == 1:
        prox = prox * N
    assert len(prox) == len(X)

    if np.isscalar(e_rel):
        e_rel = (e_rel,) * N
    assert len(e_rel) == len(X)

    if not hasattr(b1, "__iter__"):
        b1 = np.array((b1,) * max_iter)
    assert len(b1) == max_iter
    assert (b1 >= 0).all() and (b1 < 1).all()

    assert b2 >= 0 and b2 < 1
    assert eps >= 0
    assert p > 0 and p <= 0.5
    scheme = scheme.lower()
    assert scheme in ["adam", "nadam", "adamx", "amsgrad", "padam", "radam"]

    if M is not None:
        assert len(M) == N and all(m.shape == x.shape for x, m in zip(X, M))
    if V is not None:
        assert len(V) == N and all(v.shape == x.shape for x, v in zip(X, V))
    if Vhat is not None:
        assert len(Vhat) == N and all(vhat.shape == x.shape for x, vhat in zip(X, Vhat))

    Y, W = _nmf_grad_target(grad, weighted=True)
    chains = _describe_all(prox, allow_none=True)
    if (Y is not None and chains is not None and _is_factor_pair(X) and _is_step(step, "step_adaprox")):
        # (step_adaprox does not depend on W, nmf.py:91-93: the weighted likelihood takes the fused loop as well)
        return _adaprox_nmf_device(X, Y, chains, scheme, b1, b2, eps, check_convergence, p, e_rel, max_iter,
                                   prox_max_iter, M, V, Vhat, callback, W=W)
    if Y is not None and _is_factor_pair(X):
        resident = _ResidentGradient(Y, W)
        try:
            return _adaprox_callbacks(X, resident, step, prox, scheme, b1, b2, eps, check_convergence, p, e_rel,
                                      max_iter, prox_max_iter, M, V, Vhat, callback)
        finally:
            resident.close()
    return _adaprox_callbacks(X, grad, step, prox, scheme, b1, b2, eps, check_convergence, p, e_rel, max_iter,
                              prox_max_iter, M, V, Vhat, callback)


def _b1_prev(b1):
    b1 = np.asarray(b1, dtype=np.float64)
    return np.concatenate((b1[-1:], b1[:-1]))  # b1[it - 1] with Python's wrap-around at it = 0 (algorithms.py:213)


def _adaprox_nmf_device(X, Y, chains, scheme, b1, b2, eps, check_convergence, p, e_rel, max_iter, prox_max_iter,
                        M, V, Vhat, callback, W=None):
    from . import nmf as _nmf

    A, S = X
    dt = np.result_type(A.dtype, S.dtype)
    prob = _nmf.Problem(Y, A, S, W=W)
    b1 = np.asarray(b1, dtype=np.float64)
    b1p = _b1_prev(b1)
    try:
        prob.adaprox_begin(chains[0], chains[1], scheme, b2, eps, p, e_rel, check_convergence, prox_max_iter,
                           has_vhat=Vhat is not None)
        if M is not None:
            prob.set(_ffi.MA, M[0]); prob.set(_ffi.MS, M[1])
        if V is not None:
            prob.set(_ffi.VA, V[0]); prob.set(_ffi.VS, V[1])
        if Vhat is not None:
            prob.set(_ffi.VHA, Vhat[0]); prob.set(_ffi.VHS, Vhat[1])
        done, converged, sub = 0, (False, False), (0, 0)
        if callback is None:
            done, converged, sub = prob.adaprox_run(max_iter, b1, b1p)
        else:
            for it in range(max_iter):
                try:
                    callback(*X, it=it)
                except StopIteration:
                    break
                prob.set(_ffi.A, A)
                prob.set(_ffi.S, S)
                n, converged, sub = prob.adaprox_run(1, b1[it:it + 1], b1p[it:it + 1])
                done += n
                _writeback(prob, X)
                if check_convergence and all(converged):
                    break
        _writeback(prob, X)
        if M is None:
            M = (prob.get(_ffi.MA, dtype=dt), prob.get(_ffi.MS, dtype=dt))
        else:
            prob.get(_ffi.MA, out=M[0]); prob.get(_ffi.MS, out=M[1])
        if V is None:
            V = (prob.get(_ffi.VA, dtype=dt), prob.get(_ffi.VS, dtype=dt))
        else:
            prob.get(_ffi.VA, out=V[0]); prob.get(_ffi.VS, out=V[1])
        if Vhat is None:
            Vhat = [None] * 2  # quirk: the running max is never persisted (algorithms.py:176-177, 356-357)
        else:
            prob.get(_ffi.VHA, out=Vhat[0]); prob.get(_ffi.VHS, out=Vhat[1])
    finally:
        prob.close()
    logger.info("Completed {0} iterations and {1} sub-iterations".format(done, [int(sub[0]), int(sub[1])]))
    if check_convergence and not all(converged):
        logger.warning("Solution did not converge")
    conv = tuple(np.bool_(c) for c in converged) if check_convergence else (None,) * 2
    return conv, M, V, Vhat


def _adaprox_callbacks(X, grad, step, prox, scheme, b1, b2, eps, check_convergence, p, e_rel, max_iter,
                       prox_max_iter, M, V, Vhat, callback):
    N = len(X)
    if M is None:
        M = tuple(np.zeros(x.shape, x.dtype) for x in X)
    if V is None:
        V = tuple(np.zeros(x.shape, x.dtype) for x in X)
    if Vhat is None:
        Vhat = [None] * N
    Sub_iter = [0] * N
    if callback is None:
        callback = utils.NullCallback()
    chains = [_device_chain(pj) if pj is not None else None for pj in prox]
    b1 = np.asarray(b1, dtype=np.float64)
    converged = (False,) * N
    it = -1
    for it in range(max_iter):
        try:
            callback(*X, it=it)
            G = utils._as_tuple(grad(*X))
            Alpha = utils._as_tuple(step(*X, it=it))
            if check_convergence:
                X_ = utils._copy_tuple(X)
            for j in range(N):
                Psi, psimax = _dev.adaprox_moments(scheme, G[j], M[j], V[j], Vhat[j], X[j], Alpha[j], b1[it],
                                                   b1[it - 1], b2, eps, p, it + 1)
                if prox[j] is not None:
                    z = X[j].copy()
                    gamma = Alpha[j] / psimax
                    for tau in range(1, prox_max_iter + 1):
                        if chains[j] is not None:
                            z_, nd, nz = _dev.adaprox_sub(chains[j], z, X[j], Psi, Alpha[j], psimax)
                        else:  # user prox: the argument is formed on the device, the callable runs on the host
                            w, _, _ = _dev.adaprox_sub([], z, X[j], Psi, Alpha[j], psimax)
                            z_ = np.asarray(prox[j](w, gamma), dtype=z.dtype)
                            nd, nz = _dev.dot_diff(z_, z, z)[1], _dev.sumsq(z)
                        conv = np.float32(nd) <= np.float32(e_rel[j] ** 2) * np.float32(nz)
                        z = z_
                        if conv:
                            break
                    logger.debug("Proximal sub-iterations for variable {}: {}".format(j, tau))
                    Sub_iter[j] += tau
                    X[j][:] = z
            if check_convergence:
                converged = []
                for j in range(N):
                    nd = _dev.dot_diff(X[j], X_[j], X[j])[1]
                    converged.append(np.float32(nd) <= np.float32(e_rel[j] ** 2) * np.float32(_dev.sumsq(X[j])))
                converged = tuple(converged)
                if all(converged):
                    break
        except StopIteration:
            break

    logger.info("Completed {0} iterations and {1} sub-iterations".format(it + 1, Sub_iter))
    if check_convergence and not all(converged):
        logger.warning("Solution did not converge")
    if not check_convergence:
        converged = (None,) * N
    return converged, M, V, Vhat


# ------------------------------------------------------------------------------------------
# ADMM family
# ------------------------------------------------------------------------------------------
def _linop(L):
    """None for the identity (fast paths below), else a utils.MatrixAdapter holding a dense matrix resident on the
    device (utils.py:38-101).  Sparse operators raise NotImplementedError."""
    if L is None:
        return None
    ad = L if isinstance(L, utils.MatrixAdapter) else utils.MatrixAdapter(L)
    return None if ad.L is None else ad


def _spec(L):
    return 1 if L is None else L.spectral_norm


def _step_g_of(step_f, N=1, M=1, norm_L2=1):
    return step_f * norm_L2 * N * M  # utils.py:279


def _mm(X, Z, U, prox_g, chain_g, step_g, dual_uses_step_g=True, L=None):
    """utils.py:295-304.  Returns (LX, R, S, norms) with norms = (|LX|, |Z'|, |L^T U'(/step_g)|, |R|, |S|).
    L = None is the identity (one fused kernel for R, S, U and the five norms); a dense L adds device GEMMs for
    L X, L^T (Z' - Z) and L^T U."""
    LX = X if L is None else L.dot(X)
    if chain_g is not None:
        Znew = _dev.add(LX, U)
        operators._apply(Znew, step_g, chain_g)
    else:
        Znew = prox_g(_dev.add(LX, U), step_g)
    R, S, norms = _dev.admm_zu(LX, Znew, Z, U, step_g, dual_uses_step_g)
    if L is None:
        return X, R, S, norms
    S = L.T.dot(S)                                       # S = -1/step_g L^T (Z' - Z)   (utils.py:300)
    lS = np.sqrt(np.float32(_dev.sumsq(S)))
    lU = np.sqrt(np.float32(_dev.sumsq(L.T.dot(U))))     # l2(L^T U [/ step_g])          (utils.py:359-362)
    if dual_uses_step_g:
        lU = lU / np.float32(step_g)
    return LX, R, S, (norms[0], norms[1], lU, norms[3], lS)


def _dX_arg(X, Zs, Us, ratios, Ls):
    """X - sum_i ratio_i L_i^T (L_i X - Z_i + U_i)   (utils.py:316-317, 331-338)"""
    if all(l is None for l in Ls):
        return _dev.admm_xarg(X, Zs, Us, ratios)
    dX = None
    for Z, U, r, L in zip(Zs, Us, ratios, Ls):
        LX = X if L is None else L.dot(X)
        t = _dev.ew(_ffi.EW_DX_ACC, LX, Z, U, None, s0=1.0)[0][0]        # L X - Z + U
        if L is not None:
            t = L.T.dot(t)
        dX = _dev.axpy(r, t, dX)
    return _dev.ew(_ffi.EW_SUB, X, dX)[0][0].astype(X.dtype, copy=False)


def _update_variables(X, Z, U, prox_f, step_f, prox_g, chain_g, step_g, dual_uses_step_g=True, L=None):
    """utils.py:307-346.  Returns (LX, R, S, norms) -- lists in the multi-constraint case.  L: None (identity), a
    MatrixAdapter, or a list of those."""
    if not hasattr(prox_g, "__iter__"):
        if prox_g is not None:
            X[:] = prox_f(_dX_arg(X, [Z], [U], [step_f / step_g], [L]), step_f)
            return _mm(X, Z, U, prox_g, chain_g, step_g, dual_uses_step_g, L=L)
        Xold = X.copy()
        X[:] = prox_f(X, step_f)
        Z[:] = X[:]
        R = np.zeros(X.shape, dtype=X.dtype)
        S = _dev.ew(_ffi.EW_SUB, X, Xold)[0][0].astype(X.dtype, copy=False)
        nX = np.sqrt(np.float32(_dev.sumsq(X)))
        LTU = U if L is None else L.T.dot(U)
        norms = (nX, nX, np.sqrt(np.float32(_dev.sumsq(LTU))), np.float32(0), np.sqrt(np.float32(_dev.sumsq(S))))
        return X, R, S, norms
    m = len(prox_g)
    Ls = L if isinstance(L, list) else [L] * m
    X[:] = prox_f(_dX_arg(X, Z, U, [step_f / step_g[i] for i in range(m)], Ls), step_f)
    LX, R, S, norms = [None] * m, [None] * m, [None] * m, [None] * m
    for i in range(m):
        LX[i], R[i], S[i], norms[i] = _mm(X, Z[i], U[i], prox_g[i], chain_g[i], step_g[i], dual_uses_step_g, L=Ls[i])
    return LX, R, S, norms


def _constraint_convergence(size, norms, e_rel, e_abs, p=None, spec=1):
    """utils.py:349-391 from the five device norms of one constraint; size = X.size, p = Z.size, spec = ||L||_s^2
    (the reference divides by L.spectral_norm, the squared norm: utils.py:357-362)."""
    lLX, lZ, lU, lR, lS = norms
    p = size if p is None else p
    e_pri = np.sqrt(p) * e_abs / spec + e_rel * np.max([lLX, lZ])
    e_dual = np.sqrt(size) * e_abs / spec + e_rel * lU
    return (lR <= e_pri) and (lS <= e_dual), (e_pri, e_dual, lR, lS)


def _init_zu(X, m=None, L=None):
    """utils.py:244-254: Z = L X (a copy), U = 0 per constraint."""
    def one(l):
        Z = X.copy() if l is None else np.array(l.dot(X), copy=True)
        return Z, np.zeros(Z.shape, dtype=Z.dtype)
    if m is None:
        return one(L)
    Ls = L if isinstance(L, list) else [L] * m
    pairs = [one(l) for l in Ls]
    return [z for z, _ in pairs], [u for _, u in pairs]


# passes executed / restarts of the last fused admm / sdmm solve (the halved-slack restarts of algorithms.py:503-512
# reset the iteration counter, so a solve can execute more passes than max_iter): diagnostics for bench.py
LAST_ADMM_STATS = {"passes": 0, "restarts": 0}


def _admm_device(X, b, step_value, chains, e_rel, e_abs, max_iter, dual_uses_step_g):
    """Fused device loop (pmx_admm_run): the whole ADMM / SDMM iteration stays on the GPU."""
    ctx = _ffi.context()
    L = _ffi.lib()
    o = _ffi.AdmmOpts()
    o.n_g = len(chains)
    for i, ch in enumerate(chains):
        o.proxs_g[i] = _ffi.make_prox(ch)
    o.e_rel, o.e_abs, o.dual_uses_step_g = float(e_rel), float(e_abs), int(dual_uses_step_g)
    h = C.c_void_p()
    _ffi.check(L.pmx_admm_create(ctx.handle, X.size, C.byref(o), C.byref(h)))
    try:
        x32 = np.ascontiguousarray(X, dtype=np.float32).reshape(-1)
        b32 = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float32), X.shape)).reshape(-1)
        vp = C.c_void_p
        _ffi.check(L.pmx_admm_set(h, x32.ctypes.data_as(vp), b32.ctypes.data_as(vp)))
        it, conv = C.c_int(0), C.c_int(0)
        err = (C.c_double * 16)()
        _ffi.check(L.pmx_admm_run(h, float(step_value), int(max_iter), C.byref(it), C.byref(conv), err))
        passes, restarts = C.c_longlong(0), C.c_int(0)
        _ffi.check(L.pmx_admm_stats(h, C.byref(passes), C.byref(restarts)))
        LAST_ADMM_STATS.update(passes=int(passes.value), restarts=int(restarts.value))
        _ffi.check(L.pmx_admm_get(h, x32.ctypes.data_as(vp)))
        X[...] = x32.reshape(X.shape)
    finally:
        _ffi.check(L.pmx_admm_destroy(h))
    errors = [tuple(X.dtype.type(err[4 * i + k]) for k in range(4)) for i in range(len(chains))]
    return bool(conv.value), errors, it.value


def _fusable_admm(X, prox_f, step_f, chains):
    return (isinstance(prox_f, utils.LeastSquaresProx) and isinstance(step_f, utils.ConstantStep)
            and np.ndim(step_f.value) == 0 and chains is not None and 1 <= len(chains) <= 4
            and all(c is not None and all(o != _ffi.OP_UNITY for (o, _, _, _) in c) for c in chains)
            and isinstance(X, np.ndarray))


def admm(
    X,
    prox_f,
    step_f,
    prox_g=None,
    step_g=None,
    L=None,
    e_rel=1e-6,
    e_abs=0,
    max_iter=1000,
    callback=None,
):
    """Linearised ADMM with one constraint (algorithms.py:426-520).  Returns ``(converged, errors)``.
    ``L``: None (identity), a dense matrix or a ``utils.MatrixAdapter`` (device GEMMs); sparse raises."""
    _L = _linop(L)
    chain_g = _device_chain(prox_g) if prox_g is not None else None

    if (_L is None and prox_g is not None and step_g is None and callback is None
            and _fusable_admm(X, prox_f, step_f, [chain_g])):
        # quirk kept: the tolerances of `admm` use the user's step_g (None) -> no division of U (algorithms.py:494-496)
        converged, errors, logged = _admm_device(X, prox_f.b, step_f.value, [chain_g], e_rel, e_abs, max_iter,
                                                 dual_uses_step_g=False)
        logger.info("Completed {0} iterations".format(logged))
        if not converged:
            logger.warning("Solution did not converge")
        return converged, errors[0]

    Z, U = _init_zu(X, L=_L)
    it = 0
    slack = 1.0
    if callback is None:
        callback = utils.NullCallback()
    converged, error = False, None
    while it < max_iter:
        callback(X, it=it)
        step_f_ = slack * step_f(X, it=it)
        if prox_g is not None and step_g is None:
            step_g_ = _step_g_of(step_f_, norm_L2=_spec(_L))
        else:
            step_g_ = step_g
        LX, R, S, norms = _update_variables(X, Z, U, prox_f, step_f_, prox_g, chain_g, step_g_,
                                            dual_uses_step_g=step_g is not None, L=_L)
        converged, error = _constraint_convergence(X.size, norms, e_rel, e_abs, p=Z.size, spec=_spec(_L))
        if converged:
            break
        it += 1
        if prox_g is not None:
            if it > 1:
                if (X == X_).all() and (R == R_).all():  # noqa: F821
                    slack /= 2
                    it = 0
                    Z, U = _init_zu(X, L=_L)
                    logger.info("Restarting with step size slack = %.3f" % slack)
            X_ = X.copy()
            R_ = R

    logger.info("Completed {0} iterations".format(it + 1))
    if not converged:
        logger.warning("Solution did not converge")
    return converged, error


def sdmm(
    X,
    prox_f,
    step_f,
    proxs_g=None,
    steps_g=None,
    Ls=None,
    e_rel=1e-6,
    e_abs=0,
    max_iter=1000,
    callback=None,
):
    """ADMM with several constraints on one variable (algorithms.py:523-650).  Returns ``converged``."""
    if proxs_g is None or not hasattr(proxs_g, "__iter__"):
        # fall back to admm, dropping e_abs like the reference (algorithms.py:568-579)
        return admm(X, prox_f, step_f, prox_g=proxs_g, step_g=steps_g, L=Ls, e_rel=e_rel, max_iter=max_iter,
                    callback=callback)
    M = len(proxs_g)
    if not hasattr(Ls, "__iter__") or hasattr(Ls, "shape"):   # None or single: M duplicates (algorithms.py:585-587)
        Ls = [Ls] * M
    assert len(Ls) == M
    _L = [_linop(l) for l in Ls]
    identity = all(l is None for l in _L)
    chains = [_device_chain(pg) for pg in proxs_g]

    if identity and steps_g is None and callback is None and _fusable_admm(X, prox_f, step_f, chains):
        converged, _, logged = _admm_device(X, prox_f.b, step_f.value, chains, e_rel, e_abs, max_iter,
                                            dual_uses_step_g=True)
        logger.info("Completed {0} iterations".format(logged))
        if not converged:
            logger.warning("Solution did not converge")
        return converged

    Z, U = _init_zu(X, M, L=_L)
    it = 0
    slack = 1.0
    if callback is None:
        callback = utils.NullCallback()
    converged = False
    while it < max_iter:
        callback(X, it=it)
        step_f_ = slack * step_f(X, it=it)
        if steps_g is None:
            steps_g_ = [_step_g_of(step_f_, M=M, norm_L2=_spec(_L[i])) for i in range(M)]
        else:
            steps_g_ = steps_g
        LX, R, S, norms = _update_variables(X, Z, U, prox_f, step_f_, proxs_g, chains, steps_g_, L=_L)
        converged = True
        for i in range(M):
            c, _ = _constraint_convergence(X.size, norms[i], e_rel, e_abs, p=Z[i].size, spec=_spec(_L[i]))
            converged &= c
        if converged:
            break
        it += 1
        if it > 1:
            if (X == X_).all() and all([(R[i] == R_[i]).all() for i in range(M)]):  # noqa: F821
                slack /= 2
                it = 0
                Z, U = _init_zu(X, M, L=_L)
                logger.info("Restarting with step size slack = %.3f" % slack)
        R_ = R
        X_ = X.copy()

    logger.info("Completed {0} iterations".format(it + 1))
    if not converged:
        logger.warning("Solution did not converge")
    return converged


def bsdmm(
    X,
    proxs_f,
    steps_f_cb,
    proxs_g=None,
    steps_g=None,
    Ls=None,
    update_order=None,
    steps_g_update="steps_f",
    max_iter=1000,
    e_rel=1e-6,
    e_abs=0,
    callback=None,
):
    """Block-SDMM: N variables x M_j constraints, Gauss-Seidel over the blocks (algorithms.py:653-850).

    This is the callback loop (``proxs_f`` / ``steps_f_cb`` are user callables); ``nmf.nmf(..., algorithm=bsdmm)``
    takes the fused device loop instead.  Returns the list ``converged``."""
    N = len(X)
    if proxs_g is None:
        proxs_g = [None] * N
    assert len(proxs_g) == N
    steps_g_update = steps_g_update.lower()
    assert steps_g_update in ["steps_f", "fixed", "relative"]
    if steps_g is not None and steps_g_update != "steps_f":
        raise NotImplementedError("steps_g_update='fixed'/'relative' with explicit steps_g (experts-only option)")

    if np.isscalar(e_rel):
        e_rel = [e_rel] * N
    if np.isscalar(e_abs):
        e_abs = [e_abs] * N
    if update_order is None:
        update_order = range(N)

    proxs_g = list(proxs_g)
    # Ls None or single: N duplicates; per block: M_j duplicates (algorithms.py:754-770)
    if not hasattr(Ls, "__iter__") or hasattr(Ls, "shape"):
        Ls = [Ls] * N
    Ls = list(Ls)
    assert len(Ls) == N
    M = [0] * N
    chains = [None] * N
    _L = [None] * N
    for j in range(N):
        if proxs_g[j] is not None:
            if not hasattr(proxs_g[j], "__iter__"):
                proxs_g[j] = [proxs_g[j]]
            M[j] = len(proxs_g[j])
            chains[j] = [_device_chain(pg) for pg in proxs_g[j]]
            if not hasattr(Ls[j], "__iter__") or hasattr(Ls[j], "shape"):
                Ls[j] = [Ls[j]] * M[j]
            assert len(Ls[j]) == M[j]
            _L[j] = [_linop(l) for l in Ls[j]]

    Z, U = [], []
    for j in range(N):
        z, u = _init_zu(X[j], None if proxs_g[j] is None else M[j], L=_L[j])
        Z.append(z)
        U.append(u)

    converged = [None] * N
    it = 0
    if callback is None:
        callback = utils.NullCallback()

    while it < max_iter:
        callback(*X, it=it)
        for j in update_order:
            proxs_f_j = partial(proxs_f, j=j, Xs=X)
            steps_f_j = steps_f_cb(X, j=j) * 1.0
            if proxs_g[j] is None:
                steps_g_j = None
            else:
                steps_g_j = [_step_g_of(steps_f_j, N=N, M=M[j], norm_L2=_spec(_L[j][i])) for i in range(M[j])]
            LX, R, S, norms = _update_variables(X[j], Z[j], U[j], proxs_f_j, steps_f_j, proxs_g[j], chains[j],
                                                steps_g_j, L=_L[j])
            if proxs_g[j] is None:
                converged[j], _ = _constraint_convergence(X[j].size, norms, e_rel[j], e_abs[j])
            else:
                ok = True
                for i in range(M[j]):
                    c, _ = _constraint_convergence(X[j].size, norms[i], e_rel[j], e_abs[j], p=Z[j][i].size,
                                                   spec=_spec(_L[j][i]))
                    ok &= c
                converged[j] = ok
        it += 1
        if all(converged):
            break

    logger.info("Completed {0} iterations".format(it))
    if not all(converged):
        logger.warning("Solution did not converge")
    return converged


def _bsdmm_nmf(Y, A, S, W, prox_A, prox_S, max_iter, e_rel, callback, proxs_g=None, e_abs=0, **kw):
    """nmf.nmf(..., algorithm=bsdmm): the closures of nmf.py:181-193 driven through algorithms.py:653-850,
    fused on the device (gradient kernel + Lipschitz steps + ADMM variable updates per block)."""
    from . import nmf as _nmf

    if _nmf._check_W(W) is not None:
        # the reference's default step closure evaluates `if W == 1` on the weight matrix (nmf.py:63, 188-193)
        raise ValueError("The truth value of an array with more than one element is ambiguous. "
                         "Use a.any() or a.all()")
    unsupported = set(kw) - {"steps_g", "Ls", "update_order", "steps_g_update"}
    if unsupported:
        raise TypeError("bsdmm() got unexpected keyword arguments %s" % sorted(unsupported))
    if kw.get("steps_g") is not None or kw.get("Ls") is not None or kw.get("update_order") is not None \
            or kw.get("steps_g_update", "steps_f").lower() != "steps_f":
        raise NotImplementedError("nmf(..., algorithm=bsdmm) on the device supports the defaults of steps_g, Ls, "
                                  "update_order and steps_g_update only")
    if proxs_g is None:
        proxs_g = [None, None]
    assert len(proxs_g) == 2
    direct = _describe_all([prox_A, prox_S])
    g_chains = []
    for pg in proxs_g:
        if pg is None:
            g_chains.append([])
            continue
        if not hasattr(pg, "__iter__"):
            pg = [pg]
        d = _describe_all(list(pg))
        if d is None or len(d) > 4:
            direct = None
            break
        g_chains.append(d)
    if direct is None or any(o == _ffi.OP_UNITY for ch in direct for (o, _, _, _) in ch):
        raise NotImplementedError("nmf(..., algorithm=bsdmm): constraints must be built-in proximal operators "
                                  "(prox_unity only inside proxs_g)")
    if np.isscalar(e_rel):
        e_rel = [e_rel] * 2
    if np.isscalar(e_abs):
        e_abs = [e_abs] * 2
    X = [A, S]
    prob = _nmf.Problem(Y, A, S)
    try:
        prob.bsdmm_begin(direct[0], direct[1], g_chains[0], g_chains[1], e_rel, e_abs)
        done, converged = 0, [False, False]
        if callback is None:
            done, converged = prob.bsdmm_run(max_iter)
        else:
            for it in range(max_iter):
                callback(*X, it=it)
                n, converged = prob.bsdmm_run(1)
                done += n
                _writeback(prob, X)
                if all(converged):
                    break
        _writeback(prob, X)
    finally:
        prob.close()
    logger.info("Completed {0} iterations".format(done))
    if not all(converged):
        logger.warning("Solution did not converge")
    return converged
