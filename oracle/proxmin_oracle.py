"""CPU oracle for the proxmin proximal-update hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy restatement of the reference algorithm (pmelchior/proxmin
v0.6.12 @ 66fed49).  It is *not* product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  The shipped path (``proxmin_b200``) never
touches it and fails loudly when its CUDA library is missing.

Parity pin: every function here performs the same NumPy operations, in the same
order and dtype, as the reference lines it cites, so on identical inputs the
results are *bit-identical* to the reference.  That is verified in the build
container by ``tests/test_oracle_pin.py`` (imports the real reference from
/root/reference when present) and on any box by the committed golden vectors
in ``tests/golden/`` (made by ``tests/golden/make_golden.py`` from the real
reference).  The reference itself ships no tests, so the known answers of its
``examples/parabola.py`` / ``examples/unmixing.py`` are part of the goldens.

All ``file:line`` citations are relative to the reference tree.
"""
import logging
from functools import partial

import numpy as np

log = logging.getLogger("proxmin.oracle")

# --------------------------------------------------------------------------
# operators.py  (elementwise proximal maps; all mutate and return X)
# --------------------------------------------------------------------------


def _thr(step, thresh, kind):
    # operators.py:4-14 (_step_gamma) + the "relative"/"absolute" switch that
    # every thresholded operator repeats (operators.py:62-66, 76-80, 119-123, 145-149)
    assert kind in ("relative", "absolute")
    return thresh * step if kind == "relative" else thresh


def prox_id(X, step):  # operators.py:20-23
    return X


def prox_zero(X, step):  # operators.py:26-30
    X[:] = np.zeros(X.shape, dtype=X.dtype)
    return X


def prox_plus(X, step):  # operators.py:33-38
    neg = X < 0
    X[neg] = 0
    return X


def prox_unity(X, step, axis=0):  # operators.py:41-45  (divide by the sum; NOT a simplex projection)
    X[:] = X / np.sum(X, axis=axis, keepdims=True)
    return X


def prox_unity_plus(X, step, axis=0):  # operators.py:48-52
    X[:] = prox_unity(prox_plus(X, step), step, axis=axis)
    return X


def prox_min(X, step, thresh=0, type="relative"):  # operators.py:55-69
    t = _thr(step, thresh, type)
    sel = X - t < 0
    X[sel] = t
    return X


def prox_max(X, step, thresh=0, type="relative"):  # operators.py:72-84
    t = _thr(step, thresh, type)
    sel = X - t > 0
    X[sel] = t
    return X


def prox_hard(X, step, thresh=0, type="relative"):  # operators.py:109-125
    t = _thr(step, thresh, type)
    sel = np.abs(X) < t
    X[sel] = 0
    return X


def prox_hard_plus(X, step, thresh=0, type="relative"):  # operators.py:128-135
    X[:] = prox_plus(prox_hard(X, step, thresh=thresh, type=type), step)
    return X


def prox_soft(X, step, thresh=0, type="relative"):  # operators.py:138-150
    t = _thr(step, thresh, type)
    X[:] = np.sign(X) * prox_plus(np.abs(X) - t, step)
    return X


def prox_soft_plus(X, step, thresh=0, type="relative"):  # operators.py:153-160
    X[:] = prox_plus(prox_soft(X, step, thresh=thresh, type=type), step)
    return X


def prox_max_entropy(X, step, gamma=1, type="relative"):  # operators.py:163-184
    from scipy.special import lambertw

    gamma_ = _thr(step, gamma, type)
    above = X > 0
    X[above] = gamma_ * np.real(lambertw(np.exp(X[above] / gamma_ - 1) / gamma_))
    return X


class AlternatingProjections:  # operators.py:187-224
    def __init__(self, prox_list=None, repeat=1):
        self.operators = list(prox_list) if prox_list is not None else []
        self.repeat = repeat

    def __call__(self, X, step):  # operators.py:203-211: reverse list order, `repeat` rounds
        for _ in range(self.repeat):
            for p in reversed(self.operators):
                X = p(X, step)
        return X


# --------------------------------------------------------------------------
# utils.py  (norms, Lipschitz constant, Nesterov sequence, ADMM primitives, L=None only)
# --------------------------------------------------------------------------


def l2sq(x):  # utils.py:257-260
    return (x ** 2).sum()


def l2(x):  # utils.py:263-266
    return np.sqrt((x ** 2).sum())


def lipschitz(L):
    """max eigenvalue of L^T L = squared spectral norm (utils.py:14-35, dense branch :20,:34)."""
    if L is None:
        return 1
    return np.real(np.linalg.eigvals(L.T.dot(L)).max())


class Nesterov:  # utils.py:193-206; the t-sequence advances on every read of .omega
    def __init__(self, accelerated=False):
        self.t, self.accelerated = 1.0, accelerated

    @property
    def omega(self):
        if not self.accelerated:
            return 0
        t_next = 0.5 * (1 + np.sqrt(4 * self.t * self.t + 1))
        om = (self.t - 1) / t_next
        self.t = t_next
        return om


class Adapter:  # utils.py:38-101 (MatrixAdapter) for L = None or a dense matrix, axis=None
    def __init__(self, L):
        spec = None
        while isinstance(L, Adapter):  # prevent cascade (:44-48)
            spec = L._spec
            L = L.L
        self.L, self._spec = L, spec

    @property
    def spectral_norm(self):  # :53-60
        if self._spec is None:
            self._spec = lipschitz(self.L)
        return self._spec

    @property
    def T(self):  # :62-67
        if self.L is None:
            return self
        return Adapter(self.L.T)

    def dot(self, X):  # :69-77: with L = None the argument itself is returned (not a copy)
        if self.L is None:
            return X
        return self.L.dot(X)


class BarzilaiBorweinStepper:  # utils.py:209-241
    def __init__(self, type=1, init_r=0.1):
        assert type in [1, 2]
        self.r = init_r
        self.type = type

    def step(self, *X, it=None, grads=None):
        N = len(X)
        if it == 0:
            self.Delta = np.array([np.inf, ] * N)
            self.X_ = tuple(x.copy() for x in X)
            self.G_ = grads
            return tuple(self.r * np.max(np.abs(X[j])) / np.max(np.abs(grads[j])) for j in range(N))
        G = grads
        S = tuple(X[j] - self.X_[j] for j in range(N))
        Y = tuple(G[j] - self.G_[j] for j in range(N))
        self.X_ = tuple(x.copy() for x in X)
        self.G_ = grads
        if self.type == 1:
            A = tuple(np.sum(S[j] ** 2) / np.sum(S[j] * Y[j]) for j in range(N))
        else:
            A = tuple(np.sum(S[j] * Y[j]) / np.sum(Y[j] ** 2) for j in range(N))
        if it <= 3:
            self.Delta = np.minimum(self.Delta, tuple(np.sqrt(np.sum(S[j] ** 2)) for j in range(N)))
        Astab = tuple(self.Delta[j] / np.sqrt(np.sum(G[j] ** 2)) for j in range(N))
        return np.minimum(np.abs(A), Astab)


def _tup(X):  # utils.py:8-12
    return X if type(X) in (list, tuple) else (X,)


_ID = Adapter(None)


def _init_zu(X, L=_ID):
    # utils.py:244-254; with MatrixAdapter(None) L.dot(X) is X itself, so Z = X.copy(), U = 0
    if not isinstance(L, list):
        Z = L.dot(X).copy()
        U = np.zeros(Z.shape, dtype=Z.dtype)
    else:
        Z, U = [], []
        for i in range(len(L)):
            Z.append(L[i].dot(X).copy())
            U.append(np.zeros(Z[i].shape, dtype=Z[i].dtype))
    return Z, U


def _mm(X, Z, U, prox_g, step_g, L=_ID):
    # utils.py:295-304
    LX = L.dot(X)
    Znew = prox_g(LX + U, step_g)
    R = LX - Znew
    S = -1 / step_g * L.T.dot(Znew - Z)
    Z[:] = Znew[:]
    U[:] += R
    return LX, R, S


def _update_variables(X, Z, U, prox_f, step_f, prox_g, step_g, L=_ID):
    # utils.py:307-346
    if not hasattr(prox_g, "__iter__"):
        if prox_g is not None:  # utils.py:315-318
            dX = step_f / step_g * L.T.dot(L.dot(X) - Z + U)
            X[:] = prox_f(X - dX, step_f)
            return _mm(X, Z, U, prox_g, step_g, L)
        S = -X.copy()  # utils.py:319-327 (no constraint: fixed-point iteration on f)
        X[:] = prox_f(X, step_f)
        Z[:] = X[:]
        R = np.zeros(X.shape, dtype=X.dtype)
        S += X
        return X, R, S
    m = len(prox_g)  # utils.py:329-345
    dX = np.sum([step_f / step_g[i] * L[i].T.dot(L[i].dot(X) - Z[i] + U[i]) for i in range(m)], axis=0)
    X[:] = prox_f(X - dX, step_f)
    LX, R, S = [None] * m, [None] * m, [None] * m
    for i in range(m):
        LX[i], R[i], S[i] = _mm(X, Z[i], U[i], prox_g[i], step_g[i], L[i])
    return LX, R, S


def _constraint_converged(X, LX, Z, U, R, S, step_g, e_rel, e_abs, L=_ID):
    # utils.py:366-391 (+ get_variable_errors :349-363)
    if isinstance(L, list):
        ok, errs = True, []
        for i in range(len(L)):
            c, e = _constraint_converged(X, LX[i], Z[i], U[i], R[i], S[i], step_g[i], e_rel, e_abs, L[i])
            ok &= c
            errs.append(e)
        return ok, errs
    e_pri = np.sqrt(Z.size) * e_abs / L.spectral_norm + e_rel * np.max([l2(LX), l2(Z)])
    if step_g is not None:
        e_dual = np.sqrt(X.size) * e_abs / L.spectral_norm + e_rel * l2(L.T.dot(U) / step_g)
    else:
        e_dual = np.sqrt(X.size) * e_abs / L.spectral_norm + e_rel * l2(L.T.dot(U))
    lR, lS = l2(R), l2(S)
    return (lR <= e_pri) and (lS <= e_dual), (e_pri, e_dual, lR, lS)


# --------------------------------------------------------------------------
# algorithms.py
# --------------------------------------------------------------------------


def pgm(X, grad, step, prox=None, accelerated=False, backtracking=False, f=None,
        e_rel=1e-6, max_iter=1000, callback=None):
    """algorithms.py:12-144.  Returns (converged, G, S, iterations)."""
    X = _tup(X)
    N = len(X)
    prox = _tup(prox)
    if len(prox) == 1:
        prox = prox * N
    assert len(prox) == N
    prox = tuple(p if p is not None else prox_id for p in prox)  # :63-64
    if np.isscalar(e_rel):
        e_rel = (e_rel,) * N
    assert len(e_rel) == N
    assert backtracking is False or f is not None
    try:  # :73-77 -- the probe really calls the step function once
        step(*X, it=0, grads=X)
        stepper = step
    except TypeError:
        stepper = lambda *X, it=None, grads=None: step(*X, it=it)  # noqa: E731
    accel = Nesterov(accelerated)
    T = [1.0] * N
    it = -1
    converged = (False,) * N
    G = S = None
    for it in range(max_iter):
        try:
            if callback is not None:
                callback(*X, it=it)
            omega = accel.omega  # :93
            if omega > 0:
                Xe = tuple(X[j] + omega * (X[j] - Xold[j]) for j in range(N))  # noqa: F821
            elif backtracking:
                Xe = tuple(x.copy() for x in X)
            else:
                Xe = X  # alias (:99)
            Xold = tuple(x.copy() for x in X)  # :102
            G = _tup(grad(*Xe))  # :105
            S = _tup(stepper(*Xe, it=it, grads=G))  # :106
            for j in range(N):  # :107-108
                X[j][:] = prox[j](Xe[j] - T[j] * S[j] * G[j], T[j] * S[j])
            if backtracking:  # :110-127
                f_now = f(*X)
                if it == 0:
                    f_prev = f(*Xold)
                while f_now > f_prev + np.sum(
                    [np.sum((X[j] - Xold[j]) * G[j]) + 0.5 / (T[j] * S[j]) * np.sum((X[j] - Xold[j]) ** 2)
                     for j in range(N)]):
                    jmax = np.argmax([np.max(np.abs(S[j] * G[j])) / np.max(np.abs(Xold[j])) for j in range(N)])
                    T[jmax] /= 2
                    X[jmax][:] = prox[jmax](Xe[jmax] - T[jmax] * S[jmax] * G[jmax], T[jmax] * S[jmax])
                    f_now = f(*X)
                f_prev = f_now
            converged = tuple(l2sq(X[j] - Xold[j]) <= e_rel[j] ** 2 * l2sq(X[j]) for j in range(N))  # :130-133
            if all(converged):
                break
        except StopIteration:
            break
    return converged, G, S, it + 1


def _moments(it, G, M, V, b1, b2):
    # first two lines of every _*_phi_psi (algorithms.py:149-150, 160-161, 172-173, ...)
    M[:] = (1 - b1[it]) * G + b1[it] * M
    V[:] = (1 - b2) * (G ** 2) + b2 * V


def _phi_psi(scheme, it, G, M, V, Vhat, b1, b2, eps, p):
    _moments(it, G, M, V, b1, b2)
    t = it + 1
    if scheme == "adam":  # :147-156
        return M / (1 - b1[it] ** t), np.sqrt(V / (1 - b2 ** t)) + eps
    if scheme == "nadam":  # :158-167
        return (b1[it] * M[:] + (1 - b1[it]) * G) / (1 - b1[it] ** t), np.sqrt(V / (1 - b2 ** t)) + eps
    if scheme in ("amsgrad", "padam", "adamx"):  # :170-219
        if Vhat is None:
            Vh = V  # local rebinding only -- the running max never persists (quirk, SURVEY 8 a-Q)
        else:
            if scheme == "adamx":
                factor = (1 - b1[it]) ** 2 / (1 - b1[it - 1]) ** 2
                Vhat[:] = np.maximum(factor * Vhat, V)
            else:
                Vhat[:] = np.maximum(Vhat, V)
            Vh = Vhat
        if eps > 0:
            Vh = np.maximum(Vh, eps)
        return M, (Vh ** p if scheme == "padam" else np.sqrt(Vh))
    if scheme == "radam":  # :222-245
        rho_inf = 2 / (1 - b2) - 1
        Phi = M / (1 - b1[it] ** t)
        rho = rho_inf - 2 * t * b2 ** t / (1 - b2 ** t)
        if rho > 4:
            Psi = np.sqrt(V / (1 - b2 ** t))
            r = np.sqrt((rho - 4) * (rho - 2) * rho_inf / (rho_inf - 4) / (rho_inf - 2) / rho)
            Psi /= r
        else:
            Psi = np.ones(G.shape, G.dtype)
        if eps > 0:
            Psi = np.maximum(Psi, np.sqrt(eps))
        return Phi, Psi
    raise AssertionError(scheme)


def adaprox(X, grad, step, prox=None, scheme="adam", b1=0.9, b2=0.999, eps=1e-8,
            check_convergence=True, p=0.25, e_rel=1e-6, max_iter=1000, prox_max_iter=1000,
            M=None, V=None, Vhat=None, callback=None):
    """algorithms.py:248-423.  Returns (converged, M, V, Vhat, iterations, sub_iterations)."""
    X = _tup(X)
    N = len(X)
    prox = _tup(prox)
    if len(prox) == 1:
        prox = prox * N
    assert len(prox) == N
    if np.isscalar(e_rel):
        e_rel = (e_rel,) * N
    assert len(e_rel) == N
    if not hasattr(b1, "__iter__"):
        b1 = np.array((b1,) * max_iter)
    assert len(b1) == max_iter
    assert (b1 >= 0).all() and (b1 < 1).all()
    assert 0 <= b2 < 1 and eps >= 0 and 0 < p <= 0.5
    scheme = scheme.lower()
    assert scheme in ("adam", "nadam", "adamx", "amsgrad", "padam", "radam")
    if M is None:
        M = tuple(np.zeros(x.shape, x.dtype) for x in X)
    if V is None:
        V = tuple(np.zeros(x.shape, x.dtype) for x in X)
    if Vhat is None:
        Vhat = [None] * N
    sub = [0] * N
    it = -1
    converged = (False,) * N
    for it in range(max_iter):
        try:
            if callback is not None:
                callback(*X, it=it)
            G = _tup(grad(*X))  # :369
            Alpha = _tup(step(*X, it=it))  # :370
            if check_convergence:
                Xold = tuple(x.copy() for x in X)
            for j in range(N):
                Phi, Psi = _phi_psi(scheme, it, G[j], M[j], V[j], Vhat[j], b1, b2, eps, p)
                X[j][:] -= Alpha[j] * Phi / Psi  # :378
                if prox[j] is not None:
                    z = X[j].copy()
                    gamma = Alpha[j] / np.max(Psi)  # :384 (global max over the block)
                    for tau in range(1, prox_max_iter + 1):
                        z_new = prox[j](z - gamma / Alpha[j] * Psi * (z - X[j]), gamma)  # :387
                        done = l2sq(z_new - z) <= e_rel[j] ** 2 * l2sq(z)  # :389 (norm of the OLD z)
                        z = z_new
                        if done:
                            break
                    sub[j] += tau
                    X[j][:] = z
            if check_convergence:
                converged = tuple(l2sq(X[j] - Xold[j]) <= e_rel[j] ** 2 * l2sq(X[j]) for j in range(N))
                if all(converged):
                    break
        except StopIteration:
            break
    if not check_convergence:
        converged = (None,) * N
    return converged, M, V, Vhat, it + 1, sub


def admm(X, prox_f, step_f, prox_g=None, step_g=None, L=None, e_rel=1e-6, e_abs=0, max_iter=1000, callback=None):
    """algorithms.py:426-520 (L = None or dense).  Returns (converged, errors, logged_iterations)."""
    _L = Adapter(L)
    Z, U = _init_zu(X, _L)
    it, slack = 0, 1.0
    converged, error = False, None
    while it < max_iter:
        if callback is not None:
            callback(X, it=it)  # un-starred (:480)
        sf = slack * step_f(X, it=it)
        sg = sf * _L.spectral_norm * 1 * 1 if (prox_g is not None and step_g is None) else step_g  # :485-488, utils.py:279
        LX, R, S = _update_variables(X, Z, U, prox_f, sf, prox_g, sg, _L)
        # quirk (:494-496): tolerances use the *user* step_g (None by default), not sg
        converged, error = _constraint_converged(X, LX, Z, U, R, S, step_g, e_rel, e_abs, _L)
        if converged:
            break
        it += 1
        if prox_g is not None:  # :503-514 restart on bit-stall
            if it > 1 and (X == Xprev).all() and (R == Rprev).all():  # noqa: F821
                slack /= 2
                it = 0
                Z, U = _init_zu(X, _L)
            Xprev = X.copy()
            Rprev = R
    return converged, error, it + 1


def sdmm(X, prox_f, step_f, proxs_g=None, steps_g=None, Ls=None, e_rel=1e-6, e_abs=0, max_iter=1000, callback=None):
    """algorithms.py:523-650 (Ls = None or dense).  Returns (converged, logged_iterations)."""
    if proxs_g is None or not hasattr(proxs_g, "__iter__"):  # :568-579 (drops e_abs)
        c, _, n = admm(X, prox_f, step_f, prox_g=proxs_g, step_g=steps_g, L=Ls, e_rel=e_rel,
                       max_iter=max_iter, callback=callback)
        return c, n
    m = len(proxs_g)
    if not hasattr(Ls, "__iter__"):  # :585-587
        Ls = [Ls] * m
    assert len(Ls) == m
    _L = [Adapter(Ls[i]) for i in range(m)]
    Z, U = _init_zu(X, _L)
    it, slack = 0, 1.0
    converged = False
    while it < max_iter:
        if callback is not None:
            callback(X, it=it)
        sf = slack * step_f(X, it=it)
        sg = [sf * _L[i].spectral_norm * 1 * m for i in range(m)] if steps_g is None else steps_g  # :611-616
        LX, R, S = _update_variables(X, Z, U, prox_f, sf, proxs_g, sg, _L)
        converged, _ = _constraint_converged(X, LX, Z, U, R, S, sg, e_rel, e_abs, _L)  # :624-626
        if converged:
            break
        it += 1
        if it > 1 and (X == Xprev).all() and all((R[i] == Rprev[i]).all() for i in range(m)):  # noqa: F821
            slack /= 2
            it = 0
            Z, U = _init_zu(X, _L)
        Rprev = R
        Xprev = X.copy()
    return converged, it + 1


def bsdmm(X, proxs_f, steps_f_cb, proxs_g=None, Ls=None, update_order=None, max_iter=1000,
          e_rel=1e-6, e_abs=0, callback=None):
    """algorithms.py:653-850 for steps_g=None, steps_g_update='steps_f' (Ls = None or dense).

    Returns (converged list, iterations)."""
    N = len(X)
    if proxs_g is None:
        proxs_g = [None] * N
    assert len(proxs_g) == N
    if np.isscalar(e_rel):
        e_rel = [e_rel] * N
    if np.isscalar(e_abs):
        e_abs = [e_abs] * N
    if update_order is None:
        update_order = range(N)
    if not hasattr(Ls, "__iter__"):  # :754-755
        Ls = [Ls] * N
    Ls = list(Ls)
    assert len(Ls) == N
    Mj = [0] * N
    proxs_g = list(proxs_g)
    for j in range(N):
        if proxs_g[j] is not None:
            if not hasattr(proxs_g[j], "__iter__"):
                proxs_g[j] = [proxs_g[j]]
            Mj[j] = len(proxs_g[j])
            if not hasattr(Ls[j], "__iter__"):  # :766-767
                Ls[j] = [Ls[j]] * Mj[j]
            assert len(Ls[j]) == Mj[j]
    _L = []
    for j in range(N):  # :773-781
        if proxs_g[j] is None:
            _L.append(Adapter(None))
        else:
            _L.append([Adapter(Ls[j][m]) for m in range(Mj[j])])
    Z, U = [], []
    for j in range(N):  # :787-790
        z, u = _init_zu(X[j], _L[j])
        Z.append(z)
        U.append(u)
    converged = [None] * N
    it = 0
    while it < max_iter:
        if callback is not None:
            callback(*X, it=it)
        for j in update_order:  # Gauss-Seidel: X is the live list (:806)
            pf = partial(proxs_f, j=j, Xs=X)
            sf = steps_f_cb(X, j=j) * 1.0
            if proxs_g[j] is None:
                sg = None
            else:
                sg = [sf * _L[j][i].spectral_norm * N * Mj[j] for i in range(Mj[j])]  # :815-819, utils.py:279
            LX, R, S = _update_variables(X[j], Z[j], U[j], pf, sf, proxs_g[j], sg, _L[j])
            converged[j], _ = _constraint_converged(X[j], LX, Z[j], U[j], R, S, sg, e_rel[j], e_abs[j], _L[j])
        it += 1
        if all(converged):
            break
    return converged, it


# --------------------------------------------------------------------------
# nmf.py
# --------------------------------------------------------------------------


def log_likelihood(*X, Y=0, W=1):  # nmf.py:13-25
    A, S = X
    return np.sum(W * (Y - A.dot(S)) ** 2) / 2


def grad_likelihood(*X, Y=0, W=1):  # nmf.py:28-41
    A, S = X
    D = W * (A.dot(S) - Y)
    return D.dot(S.T), A.T.dot(D)


def step_pgm(*X, it=None, W=1):  # nmf.py:52-65, W == 1 branch; :44-49
    A, S = X
    return 1 / lipschitz(S.T), 1 / lipschitz(A)


def step_adaprox(*X, it=None):  # nmf.py:91-93
    A, S = X
    return (np.mean(A, axis=0) / 10, S.mean(axis=1)[:, None] / 10)


def nmf(Y, A, S, W=1, prox_A=prox_plus, prox_S=prox_plus, algorithm="pgm", step=None,
        max_iter=1000, e_rel=1e-3, callback=None, **kw):
    """nmf.py:96-203.  ``algorithm`` is one of 'pgm' | 'adaprox' | 'bsdmm'."""
    grad = partial(grad_likelihood, Y=Y, W=W)
    X = [A, S]
    prox = [prox_A, prox_S]
    if algorithm == "pgm":  # :150-162
        return pgm(X, grad, step or partial(step_pgm, W=W), prox=prox, max_iter=max_iter,
                   e_rel=e_rel, callback=callback, **kw)
    if algorithm == "adaprox":  # :164-176
        return adaprox(X, grad, step or step_adaprox, prox=prox, max_iter=max_iter,
                       e_rel=e_rel, callback=callback, **kw)
    if algorithm == "bsdmm":  # :178-203

        def prox_f(Xj, st, Xs=None, j=None):  # :181-185 (both gradients evaluated, one used)
            return prox[j](Xj - st * grad(*Xs)[j], st)

        def step_f(Xs, j=None):  # :190-193
            return step_pgm(*Xs)[j]

        return bsdmm(X, prox_f, step_f, max_iter=max_iter, e_rel=e_rel, callback=callback, **kw)
    raise AssertionError(algorithm)
