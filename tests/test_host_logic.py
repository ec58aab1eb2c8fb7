"""Host-side logic of the drop-in layer (no GPU needed): operator translation, recognition of the
library callables that select the fused device loops, API surface / signatures vs the reference."""
import inspect
from functools import partial

import numpy as np
import pytest

import apis
import proxmin_b200 as pmx
from proxmin_b200 import _ffi, algorithms, operators, workloads


def test_describe_builtins_and_partials():
    d = operators.describe
    assert d(None) == [] and d(pmx.prox_id) == []
    assert d(pmx.prox_plus) == [(_ffi.OP_PLUS, 0, 0, 0.0)]
    assert d(pmx.prox_unity_plus) == [(_ffi.OP_PLUS, 0, 0, 0.0), (_ffi.OP_UNITY, 0, 0, 0.0)]
    assert d(partial(pmx.prox_unity, axis=1)) == [(_ffi.OP_UNITY, 0, 1, 0.0)]
    assert d(partial(pmx.prox_soft, thresh=0.3)) == [(_ffi.OP_SOFT, True, 0, 0.3)]
    assert d(partial(pmx.prox_soft_plus, thresh=0.3, type="absolute")) == [(_ffi.OP_SOFT, False, 0, 0.3),
                                                                            (_ffi.OP_PLUS, 0, 0, 0.0)]
    assert d(partial(pmx.prox_hard, thresh=2)) == [(_ffi.OP_HARD, True, 0, 2.0)]
    assert d(lambda X, s: X) is None
    assert d(partial(pmx.prox_soft, thresh=0.3, bogus=1)) is None


def test_describe_alternating_projections_order_and_repeat():
    # operators.py:207-211: the list is applied in REVERSE order, `repeat` times
    ap = pmx.AlternatingProjections([pmx.prox_unity, pmx.prox_plus], repeat=2)
    assert [o for (o, _, _, _) in operators.describe(ap)] == [_ffi.OP_PLUS, _ffi.OP_UNITY] * 2
    assert ap.find(pmx.prox_plus) == 1 and ap.find(pmx.prox_soft) == -1
    ap2 = pmx.AlternatingProjections([partial(pmx.prox_max, thresh=1.0), lambda X, s: X])
    assert operators.describe(ap2) is None
    assert ap2.find(pmx.prox_max) == 0


def test_fast_path_recognition():
    Y = np.zeros((4, 6), np.float32)
    g = partial(pmx.nmf.grad_likelihood, Y=Y)
    assert algorithms._nmf_grad_target(g) is Y
    assert algorithms._nmf_grad_target(partial(pmx.nmf.grad_likelihood, Y=Y, W=1)) is Y
    assert algorithms._nmf_grad_target(partial(pmx.nmf.grad_likelihood, Y=Y, W=np.ones_like(Y))) is None
    assert algorithms._nmf_grad_target(lambda *X: X) is None
    assert algorithms._is_step(pmx.nmf.step_pgm, "step_pgm")
    assert algorithms._is_step(partial(pmx.nmf.step_pgm, W=1), "step_pgm")
    assert not algorithms._is_step(lambda *X, it=None: 1.0, "step_pgm")
    assert algorithms._is_step(pmx.nmf.step_adaprox, "step_adaprox")
    A, S = np.zeros((4, 2)), np.zeros((2, 6))
    assert algorithms._is_factor_pair((A, S)) and not algorithms._is_factor_pair((A, A))


def test_b1_prev_wraps_like_python_indexing():
    b1 = np.array([0.1, 0.2, 0.3])
    assert np.array_equal(algorithms._b1_prev(b1), [0.3, 0.1, 0.2])   # b1[it - 1] at it = 0 is b1[-1]


def test_nesterov_sequence_matches_oracle():
    from oracle import proxmin_oracle as orc

    a, b = pmx.utils.NesterovAccelerator(True), orc.Nesterov(True)
    assert [a.omega for _ in range(20)] == [b.omega for _ in range(20)]
    assert pmx.utils.NesterovAccelerator(False).omega == 0


def test_shard_columns_partition():
    for N, world, align in [(65536, 8, 128), (1000, 3, 128), (203, 2, 1), (131072, 8, 128), (100, 8, 128)]:
        spans = [workloads.shard_columns(N, world, r, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        units = [-(-(hi - lo) // align) for lo, hi in spans]   # stripes of `align` columns (the last may be ragged)
        assert max(units) - min(units) <= 1
        assert all(lo % align == 0 or lo == N for lo, _ in spans)


def test_argument_validation_matches_reference():
    X = np.zeros(3)
    with pytest.raises(AssertionError):
        pmx.pgm(X, lambda x: x, lambda x, it=None: 1.0, prox=[None, None])
    with pytest.raises(AssertionError):
        pmx.pgm(X, lambda x: x, lambda x, it=None: 1.0, backtracking=True)
    with pytest.raises(AssertionError):
        pmx.adaprox(X, lambda x: x, lambda x, it=None: 1.0, scheme="sgd")
    with pytest.raises(AssertionError):
        pmx.adaprox(X, lambda x: x, lambda x, it=None: 1.0, b2=1.0)
    with pytest.raises(AssertionError):
        pmx.adaprox(X, lambda x: x, lambda x, it=None: 1.0, b1=np.array([0.5, 0.5]), max_iter=3)
    with pytest.raises(AssertionError):
        pmx.nmf.nmf(np.zeros((2, 2)), np.zeros((2, 1)), np.zeros((1, 2)), algorithm=pmx.admm)
    with pytest.raises(AssertionError):
        pmx.prox_soft(X, 1.0, type="bogus")
    import scipy.sparse
    with pytest.raises(NotImplementedError):   # dense L runs on the device; sparse operators are out of scope
        pmx.admm(X, lambda x, s: x, lambda x, it=None: 1.0, L=scipy.sparse.eye(3))


REF = apis.reference()


@pytest.mark.skipif(REF is None, reason="reference tree only exists in the build container")
def test_public_names_and_signatures_equal_the_reference():
    import proxmin_ref as ref

    names = ["pgm", "adaprox", "admm", "sdmm", "bsdmm", "prox_id", "prox_zero", "prox_plus", "prox_unity",
             "prox_unity_plus", "prox_min", "prox_max", "prox_hard", "prox_hard_plus", "prox_soft", "prox_soft_plus"]
    for n in names:
        assert str(inspect.signature(getattr(pmx, n))) == str(inspect.signature(getattr(ref, n))), n
    for n in ["nmf", "grad_likelihood", "log_likelihood", "step_pgm", "step_adaprox", "step_A", "step_S"]:
        a, b = inspect.signature(getattr(pmx.nmf, n)), inspect.signature(getattr(ref.nmf, n))
        assert list(a.parameters) == list(b.parameters), n
    assert str(inspect.signature(pmx.AlternatingProjections.__init__)) == \
        str(inspect.signature(ref.AlternatingProjections.__init__))
    for mod in ("nmf", "utils", "algorithms", "operators"):
        assert hasattr(pmx, mod)
    for n in ("Traceback", "NullCallback", "NesterovAccelerator", "get_spectral_norm", "l2sq", "l2", "_as_tuple",
              "_copy_tuple"):
        assert hasattr(pmx.utils, n), n
    # every public name of the reference's utils / operators modules exists with the same parameters
    for modname in ("utils", "operators"):
        rmod, pmod = getattr(ref, modname), getattr(pmx, modname)
        for n in dir(rmod):
            obj = getattr(rmod, n)
            if n.startswith("_") or getattr(obj, "__module__", None) != rmod.__name__:
                continue
            assert hasattr(pmod, n), "%s.%s is missing" % (modname, n)
            if inspect.isfunction(obj):
                assert list(inspect.signature(getattr(pmod, n)).parameters) == list(inspect.signature(obj).parameters), n
            elif inspect.isclass(obj) and "__init__" in vars(obj):
                assert list(inspect.signature(getattr(pmod, n).__init__).parameters)[:len(inspect.signature(obj.__init__).parameters)] \
                    == list(inspect.signature(obj.__init__).parameters), n


def test_recognition_of_round2_callables():
    """operators.describe / algorithms._nmf_grad_target / _linop / _scalar_step on the round-2 additions (host logic
    only: nothing here touches the device)"""
    from functools import partial

    from proxmin_b200 import _ffi, algorithms, operators, utils

    assert operators.describe(operators.prox_max_entropy) == [(_ffi.OP_MAXENT, True, 0, 1.0)]
    assert operators.describe(partial(operators.prox_max_entropy, gamma=2.5, type="absolute")) == \
        [(_ffi.OP_MAXENT, False, 0, 2.5)]
    assert operators.describe(partial(operators.prox_max_entropy, gamma=np.ones(3))) is None   # array parameter: host
    ap = operators.AlternatingProjections([operators.prox_plus, partial(operators.prox_max_entropy, gamma=0.5)])
    assert [o for (o, _, _, _) in operators.describe(ap)] == [_ffi.OP_MAXENT, _ffi.OP_PLUS]   # reverse list order

    Y = np.ones((4, 6), np.float32)
    W = np.full((4, 6), 0.5, np.float32)
    g = partial(pmx.nmf.grad_likelihood, Y=Y)
    gw = partial(pmx.nmf.grad_likelihood, Y=Y, W=W)
    assert algorithms._nmf_grad_target(g) is Y
    assert algorithms._nmf_grad_target(gw) is None                       # the unweighted fused PGM loop must not take it
    y2, w2 = algorithms._nmf_grad_target(gw, weighted=True)
    assert y2 is Y and w2 is W
    assert algorithms._nmf_grad_target(partial(pmx.nmf.grad_likelihood, Y=Y, W=np.ones((3, 3))), weighted=True) == (None, None)
    assert algorithms._nmf_grad_target(lambda *X: X, weighted=True) == (None, None)

    assert algorithms._linop(None) is None
    ident = utils.MatrixAdapter(None)
    assert algorithms._linop(ident) is None and ident.T is ident and ident.spectral_norm == 1
    x = np.arange(3.0)
    assert ident.dot(x) is x                                             # NOT a copy (utils.py:70-74)
    assert algorithms._spec(None) == 1
    nested = utils.MatrixAdapter(utils.MatrixAdapter(np.eye(2, dtype=np.float32)))   # cascade is unwrapped
    assert isinstance(nested.L, np.ndarray) and nested.shape == (2, 2) and len(nested) == 2

    assert algorithms._scalar_step(np.float32(0.25)) == 0.25
    assert algorithms._scalar_step(np.array([0.5])) == 0.5               # BarzilaiBorweinStepper's ndarray of one step
    with pytest.raises(NotImplementedError):
        algorithms._scalar_step(np.array([0.5, 0.25]))
    assert utils.get_step_g(0.5, 4.0, N=2, M=3) == 12.0
    Z, U = utils.initZU(x, ident)
    assert Z is not x and np.array_equal(Z, x) and not U.any()
    with pytest.raises(NameError):
        operators.prox_components(np.zeros(3), 1.0)
    with pytest.raises(ValueError):                                      # `if W == 1` on an array (nmf.py:63)
        pmx.nmf.step_pgm(np.ones((2, 1)), np.ones((1, 2)), W=np.ones((2, 2)))
