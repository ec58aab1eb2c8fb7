"""Seeded test cases, written once and run against any implementation (see apis.py).

Every case takes an ``api`` namespace and returns a flat dict of numpy arrays /
scalars.  ``tests/golden/make_golden.py`` runs them on the real reference and
stores the results; the CPU tests replay them on the oracle (bit-exact), the GPU
tests replay them on the CUDA path (tolerances in test_gpu_parity.py).
"""
from functools import partial

import numpy as np

from proxmin_b200 import workloads

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


# --------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------

def edge_array(dtype=np.float32):
    rng = np.random.default_rng(42)
    x = rng.standard_normal((6, 40)).astype(dtype)
    x[0, :8] = [0.0, -0.0, 0.25, -0.25, 1e-30, -1e-30, 0.5, -0.5]
    x[1, :4] = [np.inf, -np.inf, np.nan, 3.0]
    return x


OPS = [
    ("plus", "prox_plus", {}),
    ("zero", "prox_zero", {}),
    ("id", "prox_id", {}),
    ("min_rel", "prox_min", dict(thresh=0.5)),
    ("min_abs", "prox_min", dict(thresh=-0.3, type="absolute")),
    ("max_rel", "prox_max", dict(thresh=0.5)),
    ("max_abs", "prox_max", dict(thresh=0.25, type="absolute")),
    ("hard_rel", "prox_hard", dict(thresh=0.5)),
    ("hard_abs", "prox_hard", dict(thresh=0.25, type="absolute")),
    ("hard_plus", "prox_hard_plus", dict(thresh=0.5)),
    ("soft_rel", "prox_soft", dict(thresh=0.5)),
    ("soft_abs", "prox_soft", dict(thresh=0.25, type="absolute")),
    ("soft_plus", "prox_soft_plus", dict(thresh=0.5)),
]


@case
def operators_elementwise(api):
    out = {}
    for tag, name, kw in OPS:
        x = edge_array()
        out[tag] = np.array(getattr(api, name)(x, 0.5, **kw))
    return out


@case
def operators_unity(api):
    rng = np.random.default_rng(3)
    base = (rng.random((16, 96)) + 0.05).astype(np.float32)
    signed = rng.standard_normal((16, 96)).astype(np.float32)
    out = {}
    for ax in (0, 1):
        out["unity_ax%d" % ax] = np.array(api.prox_unity(base.copy(), 1.0, axis=ax))
        out["unity_plus_ax%d" % ax] = np.array(api.prox_unity_plus(signed.copy(), 1.0, axis=ax))
    ap = api.AlternatingProjections([api.prox_unity, api.prox_plus])  # plus first, then unity
    out["altproj"] = np.array(ap(signed.copy(), 1.0))
    ap2 = api.AlternatingProjections([partial(api.prox_max, thresh=0.1, type="absolute"), api.prox_plus], repeat=2)
    out["altproj2"] = np.array(ap2(signed.copy(), 1.0))
    return out


# --------------------------------------------------------------------------
# generic solvers on the reference's own example problem (examples/parabola.py)
# --------------------------------------------------------------------------

_dX = np.array([1, 0.5])


def _grad_f(X):
    return 2 * (X - _dX)


def _step_f(X, it=0):
    return 0.1 * 1 / 2


def _prox_circle(X, step):
    phi = np.arctan2(X[1], X[0])
    return 0.5 * np.array([np.cos(phi), np.sin(phi)])


def _prox_line(X, step):
    return np.array([min(X[0], 0.5), min(X[1], -0.75)])


def _prox_gradf(X, step):
    return X - step * _grad_f(X)


def _parabola(api, prox):
    X0 = np.array([-1.0, -1])
    out = {}
    mi = 1000
    X = X0.copy(); api.pgm(X, _grad_f, _step_f, max_iter=mi); out["pgm_free"] = X
    X = X0.copy(); api.pgm(X, _grad_f, _step_f, prox=prox, max_iter=mi); out["pgm"] = X
    X = X0.copy(); api.pgm(X, _grad_f, _step_f, prox=prox, max_iter=mi, accelerated=True); out["apgm"] = X
    for scheme in ["adam", "nadam", "adamx", "amsgrad", "padam", "radam"]:
        X = X0.copy()
        b1 = 0.0
        if scheme != "adam":
            b1 = b1 ** np.arange(1, mi + 1)
        api.adaprox(X, _grad_f, _step_f, prox=prox, b1=b1, b2=0.5, max_iter=mi, scheme=scheme, p=0.125)
        out[scheme] = X
    X = X0.copy(); api.admm(X, _prox_gradf, _step_f, prox_g=prox, max_iter=mi); out["admm"] = X
    X = X0.copy()
    api.admm(X, lambda X, s: prox(_prox_gradf(X, s), s), _step_f, prox_g=None, max_iter=mi)
    out["admm_direct"] = X
    X = X0.copy(); api.sdmm(X, _prox_gradf, _step_f, proxs_g=[prox] * 2, max_iter=mi); out["sdmm"] = X
    return out


@case
def parabola_circle(api):
    return _parabola(api, _prox_circle)


@case
def parabola_line(api):
    return _parabola(api, _prox_line)


# --------------------------------------------------------------------------
# NMF hot path
# --------------------------------------------------------------------------

def _pack(api, A, S, extra=None):
    n, sub = api.iterations()
    out = {"A": A, "S": S, "iterations": np.int64(n if n is not None else -1)}
    if sub is not None:
        out["sub_iterations"] = np.array(sub, dtype=np.int64)
    if extra:
        out.update(extra)
    return out


@case
def nmf_pgm_cfg1(api):
    """BASELINE config 1 (256x512, K=8, plus/plus) for 100 fixed iterations."""
    Y, A, S = workloads.cfg1()
    conv, G, st = api.nmf.nmf(Y, A, S, max_iter=100, e_rel=0)
    return _pack(api, A, S, {"G_A": G[0], "G_S": G[1], "step_A": np.float64(st[0]), "step_S": np.float64(st[1])})


@case
def nmf_pgm_cfg1_converge(api):
    """Same data, default e_rel=1e-3: the stopping iteration is part of parity."""
    Y, A, S = workloads.cfg1()
    conv, G, st = api.nmf.nmf(Y, A, S, max_iter=1000)
    return _pack(api, A, S, {"converged": np.array(conv, dtype=bool)})


@case
def nmf_pgm_unity(api):
    """cfg2 recipe scaled down, prox_A=plus, prox_S=unity_plus (columns of S sum to one)."""
    Y, A, S = workloads.cfg2(128, 384, 16, seed=5)
    api.nmf.nmf(Y, A, S, prox_A=api.prox_plus, prox_S=api.prox_unity_plus, max_iter=60, e_rel=0)
    return _pack(api, A, S)


@case
def nmf_pgm_altproj(api):
    """The AlternatingProjections spelling of unity_plus must give the same path."""
    Y, A, S = workloads.cfg2(128, 384, 16, seed=5)
    pS = api.AlternatingProjections([api.prox_unity, api.prox_plus])
    api.nmf.nmf(Y, A, S, prox_A=api.prox_plus, prox_S=pS, max_iter=60, e_rel=0)
    return _pack(api, A, S)


@case
def nmf_pgm_accel(api):
    """Nesterov-accelerated block PGM (algorithms.py:93-95) through the generic ``pgm`` entry point.

    The reference's accelerated NMF diverges with the plain Lipschitz steps, so the
    steps are halved by a user step function (a user callable: the generic path)."""
    Y, A, S = workloads.cfg1(96, 200, 6, seed=11)

    def step(*X, it=None):
        return tuple(0.5 * s for s in api.nmf.step_pgm(*X))

    api.pgm([A, S], partial(api.nmf.grad_likelihood, Y=Y), step, prox=[api.prox_plus] * 2,
            max_iter=60, e_rel=0, accelerated=True)
    return _pack(api, A, S)


@case
def nmf_pgm_backtracking(api):
    """PGM with the Beck-Teboulle backtracking line search (algorithms.py:110-127) the way examples/unmixing.py
    drives it: f = log_likelihood at the trial point.  The Lipschitz step of A is tripled by a user step function so
    that the line search really has to halve T (five halvings in the reference, none with the plain steps; making
    BOTH steps too long sends the reference into ~120 halvings and an overflow -- not a useful test)."""
    Y, A, S = workloads.cfg1(96, 200, 6, seed=22)

    def step(*X, it=None):
        sA, sS = api.nmf.step_pgm(*X)
        return (3.0 * sA, sS)

    api.pgm([A, S], partial(api.nmf.grad_likelihood, Y=Y), step, prox=[api.prox_plus] * 2,
            max_iter=40, e_rel=0, backtracking=True, f=partial(api.nmf.log_likelihood, Y=Y))
    return _pack(api, A, S, {"loss": np.float64(api.nmf.log_likelihood(A, S, Y=Y))})


@case
def nmf_pgm_soft(api):
    """L1-regularised S (prox_soft_plus, relative threshold scales with the Lipschitz step)."""
    Y, A, S = workloads.cfg1(96, 200, 6, seed=12)
    api.nmf.nmf(Y, A, S, prox_S=partial(api.prox_soft_plus, thresh=0.5), max_iter=60, e_rel=0)
    return _pack(api, A, S)


@case
def nmf_pgm_ragged(api):
    """Shapes that are not multiples of any tile size."""
    Y, A, S = workloads.cfg1(77, 203, 5, seed=13)
    api.nmf.nmf(Y, A, S, max_iter=40, e_rel=0)
    return _pack(api, A, S)


@case
def nmf_adaprox_amsgrad(api):
    """config 3 recipe scaled down: adaprox/AMSGrad, plus/plus, default step_adaprox."""
    Y, A, S = workloads.cfg2(128, 384, 16, seed=6)
    conv, M, V, Vh = api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", max_iter=40,
                                 check_convergence=False)
    return _pack(api, A, S, {"M_A": M[0], "V_S": V[1]})


@case
def nmf_adaprox_amsgrad_unity(api):
    Y, A, S = workloads.cfg2(128, 384, 16, seed=6)
    api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", prox_S=api.prox_unity_plus,
                max_iter=40, check_convergence=False)
    return _pack(api, A, S)


@case
def nmf_adaprox_schemes(api):
    out = {}
    for scheme in ["adam", "nadam", "padam", "adamx"]:
        Y, A, S = workloads.cfg2(64, 192, 8, seed=7)
        api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme=scheme, max_iter=25, e_rel=1e-3)
        n, sub = api.iterations()
        out[scheme + "_A"], out[scheme + "_S"] = A, S
        out[scheme + "_iterations"] = np.int64(n)
        out[scheme + "_sub"] = np.array(sub, dtype=np.int64)
    return out


@case
def nmf_bsdmm(api):
    """config 5 recipe scaled down: bsdmm, direct prox_id, constraints via proxs_g.

    Horizon: 6 outer iterations.  The reference's own bsdmm trajectory on this problem is chaotic: scaling Y
    by (1 + 1e-6) changes the reference's A by 2e-6 after 6, 7e-5 after 12 and 18 % after 20 iterations
    (measured with the oracle), so longer horizons cannot separate implementation error from sensitivity."""
    Y, A, S = workloads.cfg5(96, 256, 8, seed=9)
    proxs_g = [[api.prox_plus, api.prox_unity], [api.prox_plus, partial(api.prox_soft, thresh=0.01)]]
    conv = api.nmf.nmf(Y, A, S, algorithm=api.bsdmm, prox_A=api.prox_id, prox_S=api.prox_id,
                       proxs_g=proxs_g, max_iter=6, e_rel=1e-6)
    return _pack(api, A, S, {"converged": np.array([bool(c) for c in conv])})


@case
def nmf_pgm_cfg1_1000(api):
    """BASELINE config 1 at the SURVEY 7.3 horizon: 1000 fixed iterations."""
    Y, A, S = workloads.cfg1()
    api.nmf.nmf(Y, A, S, max_iter=1000, e_rel=0)
    return _pack(api, A, S)


@case
def nmf_pgm_k96(api):
    """K > 64 (the tcgen05 kernel's two k-panel configuration; SIMT before it existed): plus / unity_plus."""
    Y, A, S = workloads.cfg2(192, 512, 96, seed=31)
    api.nmf.nmf(Y, A, S, prox_A=api.prox_plus, prox_S=api.prox_unity_plus, max_iter=40, e_rel=0)
    return _pack(api, A, S)


@case
def nmf_bsdmm_k128(api):
    """BASELINE config 5 scaled down in M and N only: K = 128 like the real thing (6 outer iterations, see nmf_bsdmm)."""
    Y, A, S = workloads.cfg5(256, 1024, 128, seed=32)
    proxs_g = [[api.prox_plus, api.prox_unity], [api.prox_plus, partial(api.prox_soft, thresh=0.01)]]
    conv = api.nmf.nmf(Y, A, S, algorithm=api.bsdmm, prox_A=api.prox_id, prox_S=api.prox_id,
                       proxs_g=proxs_g, max_iter=6, e_rel=1e-6)
    return _pack(api, A, S, {"converged": np.array([bool(c) for c in conv])})


@case
def nmf_adaprox_soft_kvector(api):
    """Quirk row of SURVEY 8 a-Q: under step_adaprox the step handed to the prox is gamma = Alpha/max(Psi), a K-vector
    for A and a K x 1 column for S, so a *relative* threshold broadcasts per component (algorithms.py:384-387)."""
    Y, A, S = workloads.cfg2(96, 256, 8, seed=33)
    api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", prox_A=partial(api.prox_soft_plus, thresh=0.05),
                prox_S=partial(api.prox_soft_plus, thresh=0.1), max_iter=30, check_convergence=False)
    return _pack(api, A, S)


@case
def nmf_adaprox_zero_column(api):
    """Quirk row: gamma/Alpha is evaluated literally (algorithms.py:384-387) -- an all-zero column of A gives
    Alpha = 0 there and 0/0 = NaN in the proximal sub-iteration; the NaN pattern is part of parity."""
    Y, A, S = workloads.cfg2(64, 128, 6, seed=34)
    A[:, 2] = 0
    with np.errstate(all="ignore"):
        api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", max_iter=1, prox_max_iter=5,
                    check_convergence=False)
    return _pack(api, A, S)


class _Stopper(object):
    """callback that keeps copies of the iterates it sees and raises StopIteration at iteration `stop_at`"""

    def __init__(self, stop_at=None):
        self.trace, self.its, self.stop_at = [], [], stop_at

    def __call__(self, *X, it=None):
        self.trace.append(tuple(np.array(x, copy=True) for x in X))
        self.its.append(it)
        if self.stop_at is not None and it == self.stop_at:
            raise StopIteration


@case
def nmf_callbacks(api):
    """callback(*X, it=it) at the top of every iteration through the fused NMF loops, utils.Traceback,
    and StopIteration ending pgm / adaprox cleanly (algorithms.py:90,137-138,368,412-413,802)."""
    out = {}
    # pgm + Traceback
    Y, A, S = workloads.cfg2(64, 160, 6, seed=35)
    tb = api.utils.Traceback()
    api.nmf.nmf(Y, A, S, prox_S=api.prox_unity_plus, max_iter=8, e_rel=0, callback=tb)
    out["pgm_trace_len"] = np.int64(len(tb.trace))
    out["pgm_trace3_A"], out["pgm_trace3_S"] = tb.trace[3][0], tb.trace[3][1]
    out["pgm_A"], out["pgm_S"] = A, S
    # pgm + StopIteration at it == 5: five iterations are executed
    Y, A, S = workloads.cfg2(64, 160, 6, seed=35)
    st = _Stopper(stop_at=5)
    api.nmf.nmf(Y, A, S, max_iter=20, e_rel=0, callback=st)
    out["pgm_stop_its"] = np.array(st.its, dtype=np.int64)
    out["pgm_stop_A"] = A
    # adaprox + StopIteration at it == 4
    Y, A, S = workloads.cfg2(64, 160, 6, seed=36)
    st = _Stopper(stop_at=4)
    api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", max_iter=20, check_convergence=False, callback=st)
    out["ada_stop_its"] = np.array(st.its, dtype=np.int64)
    out["ada_stop_A"], out["ada_stop_S"] = A, S
    out["ada_trace2_S"] = st.trace[2][1]
    # bsdmm + Traceback
    Y, A, S = workloads.cfg5(64, 160, 6, seed=37)
    tb = api.utils.Traceback()
    proxs_g = [[api.prox_plus, api.prox_unity], [api.prox_plus]]
    api.nmf.nmf(Y, A, S, algorithm=api.bsdmm, prox_A=api.prox_id, prox_S=api.prox_id, proxs_g=proxs_g, max_iter=4,
                e_rel=1e-6, callback=tb)
    out["bsdmm_trace_len"] = np.int64(len(tb.trace))
    out["bsdmm_trace2_A"] = tb.trace[2][0]
    out["bsdmm_A"] = A
    return out


@case
def admm_callback_unstarred(api):
    """Quirk row: admm / sdmm hand the array itself to the callback, callback(X, it=it) (algorithms.py:480, 605),
    while pgm / adaprox / bsdmm star the tuple."""
    b, X = workloads.cfg4(2_000, seed=38)
    prox_f, step_f, prox_g = lasso_callables(api, b)
    seen = []

    def cb(*args, it=None):
        seen.append((len(args), np.ndim(args[0]), np.shape(args[0])[0], it))

    api.admm(X, prox_f, step_f, prox_g=prox_g, max_iter=3, e_rel=1e-12, callback=cb)
    X2 = np.zeros_like(b)
    api.sdmm(X2, prox_f, step_f, proxs_g=[prox_g, api.prox_plus], max_iter=3, e_rel=1e-12, callback=cb)
    return {"seen": np.array(seen, dtype=np.int64), "X": X, "X2": X2}


@case
def grad_and_loss(api):
    Y, A, S = workloads.cfg2(200, 333, 24, seed=21)
    gA, gS = api.nmf.grad_likelihood(A, S, Y=Y)
    return {"G_A": gA, "G_S": gS, "loss": np.float64(api.nmf.log_likelihood(A, S, Y=Y)),
            "step_A": np.float64(api.nmf.step_pgm(A, S)[0]), "step_S": np.float64(api.nmf.step_pgm(A, S)[1])}


# --------------------------------------------------------------------------
# ADMM LASSO (BASELINE config 4 scaled down)
# --------------------------------------------------------------------------

def lasso_callables(api, b):
    def prox_f(X, s):  # README gradient-step pattern for f = 0.5 ||X - b||^2
        return X - s * (X - b)

    def step_f(X, it=None):
        return 0.5

    return prox_f, step_f, partial(api.prox_soft, thresh=0.5)


@case
def admm_lasso(api):
    b, X = workloads.cfg4(10_000, seed=7)
    prox_f, step_f, prox_g = lasso_callables(api, b)
    conv, err = api.admm(X, prox_f, step_f, prox_g=prox_g, max_iter=200, e_rel=1e-6)
    n, _ = api.iterations()
    return {"X": X, "converged": np.bool_(conv), "errors": np.array(err, dtype=np.float64),
            "iterations": np.int64(n)}


@case
def sdmm_lasso_plus(api):
    b, X = workloads.cfg4(10_000, seed=8)
    prox_f, step_f, prox_g = lasso_callables(api, b)
    conv = api.sdmm(X, prox_f, step_f, proxs_g=[prox_g, api.prox_plus], max_iter=100, e_rel=1e-5)
    n, _ = api.iterations()
    return {"X": X, "converged": np.bool_(conv), "iterations": np.int64(n)}


# --------------------------------------------------------------------------
# prox_max_entropy, BarzilaiBorweinStepper, non-identity L, weighted likelihood
# --------------------------------------------------------------------------

@case
def operators_max_entropy(api):
    """operators.py:163-184: Lambert-W prox of gamma sum x ln x; only X > 0 is touched, exp overflows to inf in fp32
    for X / gamma_ > 89 (the reference then returns inf)."""
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((5, 48)) * 2).astype(np.float32)
    x[0, :8] = [0.0, -0.0, 1e-6, 0.5, 1.0, 7.5, 30.0, 60.0]
    x[1, :3] = [np.nan, -np.inf, -3.0]
    out = {}
    with np.errstate(over="ignore", invalid="ignore"):
        out["rel"] = np.array(api.prox_max_entropy(x.copy(), 0.5, gamma=1))
        out["rel_small"] = np.array(api.prox_max_entropy(x.copy(), 0.25, gamma=0.3))
        out["abs"] = np.array(api.prox_max_entropy(x.copy(), 0.5, gamma=2.0, type="absolute"))
        x64 = x.astype(np.float64)
        out["f64"] = np.array(api.prox_max_entropy(x64, 0.5, gamma=1))
    return out


@case
def pgm_barzilai_borwein(api):
    """utils.py:209-241: both BB step types through ``pgm``.  The stepper returns an ndarray from the second
    iteration on, which ``pgm`` wraps as ONE step (utils.py:8-12): the reference supports it for a single
    variable only -- non-negative least squares  min |Y - A S|^2 over S >= 0 with A fixed."""
    out = {}
    for typ in (1, 2):
        Y, A, S = workloads.cfg1(48, 80, 4, seed=21)

        def grad(S_):
            return A.T.dot(A.dot(S_) - Y)

        bb = api.utils.BarzilaiBorweinStepper(type=typ, init_r=0.1)
        conv, G, st = api.pgm(S, grad, bb.step, prox=api.prox_plus, max_iter=25, e_rel=0)
        out["S%d" % typ] = S
        out["step%d" % typ] = np.asarray(st, dtype=np.float64).reshape(-1)
    return out


def _lasso_L(seed, p, n):
    rng = np.random.default_rng(seed)
    L = (rng.standard_normal((p, n)) / np.sqrt(n)).astype(np.float32)
    xs = np.zeros(n, np.float32)
    xs[rng.choice(n, n // 8, replace=False)] = (2 * rng.standard_normal(n // 8)).astype(np.float32)
    b = (xs + 0.05 * rng.standard_normal(n)).astype(np.float32)
    return L, b


@case
def admm_dense_L(api):
    """admm / sdmm / bsdmm with non-identity dense linear operators (utils.py:38-101, :295-346): generalised
    LASSO  min 0.5 |x - b|^2 + lambda |L x|_1  and friends."""
    out = {}
    L, b = _lasso_L(31, 24, 40)

    def prox_f(X, step):
        return X - step * (X - b)

    def step_f(X, it=None):
        return np.float32(0.5)

    X = np.zeros(40, np.float32)
    conv, err = api.admm(X, prox_f, step_f, prox_g=partial(api.prox_soft, thresh=0.1), L=L, max_iter=60, e_rel=1e-4)
    out["admm_X"], out["admm_err"] = X, np.asarray(err, dtype=np.float64)
    out["admm_it"] = np.int64(api.iterations()[0])

    L2, _ = _lasso_L(32, 40, 40)
    X = np.zeros(40, np.float32)
    api.sdmm(X, prox_f, step_f, proxs_g=[partial(api.prox_soft, thresh=0.1), api.prox_plus], Ls=[L, None],
             max_iter=60, e_rel=1e-4)
    out["sdmm_X"] = X
    out["sdmm_it"] = np.int64(api.iterations()[0])

    rng = np.random.default_rng(33)
    b1 = rng.standard_normal(20).astype(np.float32)
    b2 = rng.standard_normal((6, 5)).astype(np.float32)
    bb = [b1, b2]
    L1 = (rng.standard_normal((12, 20)) / 4).astype(np.float32)
    L3 = (rng.standard_normal((6, 6)) / 2).astype(np.float32)

    def proxs_f(X, step, Xs=None, j=None):
        return X - step * (X - bb[j])

    def steps_f(Xs, j=None):
        return np.float32(0.4)

    X1, X2 = np.zeros(20, np.float32), np.zeros((6, 5), np.float32)
    api.bsdmm([X1, X2], proxs_f, steps_f, proxs_g=[[partial(api.prox_soft, thresh=0.05)], [api.prox_plus, api.prox_plus]],
              Ls=[[L1], [L3, None]], max_iter=40, e_rel=1e-4)
    out["bsdmm_X1"], out["bsdmm_X2"] = X1, X2
    out["bsdmm_it"] = np.int64(api.iterations()[0])
    return out


def _weights(shape, seed):
    rng = np.random.default_rng(seed)
    W = (0.25 + 0.75 * rng.random(shape)).astype(np.float32)
    W[rng.random(shape) < 0.05] = 0.0        # masked entries
    return W


@case
def nmf_weighted(api):
    """Weighted likelihood (nmf.py:25, 40): gradient / loss with an M x N weight matrix, adaprox (whose default step
    does not depend on W) and PGM with a user step (the reference's weighted step_pgm is broken)."""
    out = {}
    Y, A, S = workloads.cfg2(96, 224 + 40, 12, seed=41)
    W = _weights(Y.shape, 42)
    gA, gS = api.nmf.grad_likelihood(A, S, Y=Y, W=W)
    out["G_A"], out["G_S"] = gA, gS
    out["loss"] = np.float64(api.nmf.log_likelihood(A, S, Y=Y, W=W))

    A1, S1 = A.copy(), S.copy()
    api.nmf.nmf(Y, A1, S1, W=W, algorithm=api.adaprox, scheme="amsgrad", max_iter=30, check_convergence=False)
    n, sub = api.iterations()
    out["ada_A"], out["ada_S"] = A1, S1
    out["ada_sub"] = np.array(sub, dtype=np.int64)

    def step(*X, it=None):
        return tuple(0.5 * s for s in api.nmf.step_pgm(*X))

    A2, S2 = A.copy(), S.copy()
    api.nmf.nmf(Y, A2, S2, W=W, step=step, prox_S=api.prox_unity_plus, max_iter=40, e_rel=0)
    out["pgm_A"], out["pgm_S"] = A2, S2
    return out
