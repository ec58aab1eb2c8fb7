"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed
golden vectors of the unmodified reference, on the same seeded inputs.

Tolerances (stated per test): the device computes in fp32 with 3xBF16-split tensor-core GEMMs
(or fp32 SIMT); the north-star bound is 1e-4 relative on the factors; support sets of
plus/hard/soft/max/min outputs must be identical wherever the reference value is not within
1e-6*max|X| of the threshold; iteration and sub-iteration counts must match where stated.
"""
import os
import sys

import numpy as np
import pytest

import apis
import cases
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def product():
    return apis.product()


def rel_fro(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def assert_close(got, want, tol, key):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, key
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want)), key + ": NaN pattern"
    assert np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)]), key + ": inf pattern"
    if fin.any():
        err = rel_fro(got[fin], want[fin])
        assert err <= tol, "%s: relative Frobenius error %.3e > %.1e" % (key, err, tol)


def assert_same_support(got, want, key, guard=1e-6):
    """zero / non-zero pattern identical except where the reference value is within guard*max|X| of zero"""
    got, want = np.asarray(got), np.asarray(want)
    fin = np.isfinite(want)
    scale = np.max(np.abs(want[fin])) if fin.any() else 1.0
    safe = fin & ((want == 0) | (np.abs(want) > guard * scale))
    assert np.array_equal((got != 0)[safe], (want != 0)[safe]), key + ": support set differs"


# ---------------------------------------------------------------- operators: bit-exact
def test_operators_elementwise_bit_exact(product):
    want = load_golden("operators_elementwise")
    got = cases.operators_elementwise(product)
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
        assert np.array_equal(np.signbit(got[k]), np.signbit(want[k])), k + " (sign of zero)"


def test_operators_unity(product):
    want = load_golden("operators_unity")
    got = cases.operators_unity(product)
    for k in want:
        # axis=0 sums run in NumPy's row order on the device as well: bit-exact; axis=1 is a warp tree
        tol = 0.0 if k.endswith("ax0") or k == "altproj" else 5e-7
        if tol == 0.0:
            assert np.array_equal(got[k], want[k]), k
        else:
            assert_close(got[k], want[k], tol, k)
        assert_same_support(got[k], want[k], k)


# ---------------------------------------------------------------- NMF building blocks
def test_grad_loss_lipschitz(product):
    want = load_golden("grad_and_loss")
    got = cases.grad_and_loss(product)
    assert_close(got["G_A"], want["G_A"], 2e-5, "G_A")   # fp32 GEMM vs OpenBLAS fp32 GEMM
    assert_close(got["G_S"], want["G_S"], 2e-5, "G_S")
    assert abs(got["loss"] - want["loss"]) <= 2e-5 * abs(want["loss"])
    assert abs(got["step_A"] - want["step_A"]) <= 5e-6 * want["step_A"]
    assert abs(got["step_S"] - want["step_S"]) <= 5e-6 * want["step_S"]


PGM_CASES = ["nmf_pgm_cfg1", "nmf_pgm_unity", "nmf_pgm_altproj", "nmf_pgm_soft", "nmf_pgm_ragged", "nmf_pgm_accel",
             "nmf_pgm_cfg1_1000", "nmf_pgm_k96"]


@pytest.mark.parametrize("name", PGM_CASES)
def test_nmf_pgm_matches_reference(product, name):
    """factors within 1e-4 relative (north-star tolerance) of the reference's fp32 run, same iteration count"""
    want = load_golden(name)
    got = cases.CASES[name](product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert_close(got["A"], want["A"], 1e-4, name + ":A")
    assert_close(got["S"], want["S"], 1e-4, name + ":S")
    assert_same_support(got["A"], want["A"], name + ":A", guard=1e-4)
    assert_same_support(got["S"], want["S"], name + ":S", guard=1e-4)
    if "G_A" in want:
        assert_close(got["G_A"], want["G_A"], 2e-3, name + ":G_A")  # gradients near the optimum are differences
        assert abs(got["step_A"] - want["step_A"]) <= 1e-4 * want["step_A"]
        assert abs(got["step_S"] - want["step_S"]) <= 1e-4 * want["step_S"]


def test_nmf_pgm_stopping_iteration(product):
    """default e_rel=1e-3: the device-side stop flag must freeze the run at the reference's iteration"""
    want = load_golden("nmf_pgm_cfg1_converge")
    got = cases.nmf_pgm_cfg1_converge(product)
    assert abs(int(got["iterations"]) - int(want["iterations"])) <= 1
    assert np.array_equal(got["converged"], want["converged"])
    assert_close(got["A"], want["A"], 2e-4, "A")
    assert_close(got["S"], want["S"], 2e-4, "S")


@pytest.mark.timeout(300)
def test_nmf_pgm_backtracking(product):
    """PGM with backtracking (algorithms.py:110-127), f = log_likelihood on the device; the same
    five halvings of T as the reference, factors within 2e-4"""
    want = load_golden("nmf_pgm_backtracking")
    got = cases.nmf_pgm_backtracking(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert_close(got["A"], want["A"], 2e-4, "A")
    assert_close(got["S"], want["S"], 2e-4, "S")
    assert abs(float(got["loss"]) - float(want["loss"])) <= 1e-3 * float(want["loss"])


def test_oracle_agrees_live(product):
    """same seeded inputs, oracle run live next to the device (not just the stored vectors)"""
    orc = apis.oracle()
    a = cases.nmf_pgm_ragged(orc)
    b = cases.nmf_pgm_ragged(product)
    assert_close(b["A"], a["A"], 1e-4, "A")
    assert_close(b["S"], a["S"], 1e-4, "S")


# ---------------------------------------------------------------- generic solvers with user callables
@pytest.mark.parametrize("name", ["parabola_circle", "parabola_line"])
def test_parabola_known_answers(product, name):
    """examples/parabola.py through every solver: user grad/step/prox callables on the host, library
    arithmetic on the device (fp32): end points within 2e-5 of the reference's fp64 run."""
    want = load_golden(name)
    got = cases.CASES[name](product)
    for k in want:
        assert np.allclose(got[k], want[k], rtol=0, atol=2e-5), (k, got[k], want[k])


# ---------------------------------------------------------------- adaprox / bsdmm on the NMF path
def test_nmf_adaprox_amsgrad(product):
    want = load_golden("nmf_adaprox_amsgrad")
    got = cases.nmf_adaprox_amsgrad(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert list(got["sub_iterations"]) == list(want["sub_iterations"])
    # adaprox normalises the gradient by sqrt(V): ~100x more sensitive than PGM; the north-star
    # bound 1e-4 still holds with the 3xBF16 split (measured 6e-6 .. 2e-5, profiles/r1_mgpu_check_2gpu.txt)
    assert_close(got["A"], want["A"], 1e-4, "A")
    assert_close(got["S"], want["S"], 1e-4, "S")
    assert_close(got["M_A"], want["M_A"], 1e-3, "M_A")
    assert_close(got["V_S"], want["V_S"], 1e-3, "V_S")


def test_nmf_adaprox_amsgrad_unity(product):
    want = load_golden("nmf_adaprox_amsgrad_unity")
    got = cases.nmf_adaprox_amsgrad_unity(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert list(got["sub_iterations"]) == list(want["sub_iterations"])
    assert_close(got["A"], want["A"], 1e-4, "A")
    assert_close(got["S"], want["S"], 1e-4, "S")


def test_nmf_adaprox_schemes(product):
    want = load_golden("nmf_adaprox_schemes")
    got = cases.nmf_adaprox_schemes(product)
    for scheme in ["adam", "nadam", "padam", "adamx"]:
        assert int(got[scheme + "_iterations"]) == int(want[scheme + "_iterations"]), scheme
        assert list(got[scheme + "_sub"]) == list(want[scheme + "_sub"]), scheme
        assert_close(got[scheme + "_A"], want[scheme + "_A"], 1e-4, scheme + ":A")
        assert_close(got[scheme + "_S"], want[scheme + "_S"], 1e-4, scheme + ":S")


def test_nmf_bsdmm(product):
    want = load_golden("nmf_bsdmm")
    got = cases.nmf_bsdmm(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert np.array_equal(got["converged"], want["converged"])
    assert_close(got["A"], want["A"], 5e-4, "A")   # 6 iterations; see cases.nmf_bsdmm for the horizon
    assert_close(got["S"], want["S"], 5e-4, "S")


def test_nmf_bsdmm_k128(product):
    """BASELINE config 5's K = 128 (scaled down in M, N only): the K > 64 gradient path under bsdmm"""
    want = load_golden("nmf_bsdmm_k128")
    got = cases.nmf_bsdmm_k128(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert np.array_equal(got["converged"], want["converged"])
    assert_close(got["A"], want["A"], 5e-4, "A")
    assert_close(got["S"], want["S"], 5e-4, "S")


def test_nmf_adaprox_soft_kvector(product):
    """relative thresholds under step_adaprox: the prox step is the K-vector gamma (SURVEY 8 a-Q)"""
    want = load_golden("nmf_adaprox_soft_kvector")
    got = cases.nmf_adaprox_soft_kvector(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert list(got["sub_iterations"]) == list(want["sub_iterations"])
    assert_close(got["A"], want["A"], 1e-4, "A")
    assert_close(got["S"], want["S"], 1e-4, "S")
    assert_same_support(got["A"], want["A"], "A", guard=1e-4)
    assert_same_support(got["S"], want["S"], "S", guard=1e-4)


def test_nmf_adaprox_zero_column_nan(product):
    """gamma/Alpha = 0/0 for an all-zero column of A (SURVEY 8 a-Q): same NaN pattern, same sub-iteration counts"""
    want = load_golden("nmf_adaprox_zero_column")
    got = cases.nmf_adaprox_zero_column(product)
    assert np.isnan(want["A"][:, 2]).all() and not np.isnan(want["S"]).any()   # what the reference does
    assert list(got["sub_iterations"]) == list(want["sub_iterations"])
    assert_close(got["A"], want["A"], 1e-5, "A")
    assert_close(got["S"], want["S"], 1e-5, "S")


def test_nmf_callbacks_through_fused_loops(product):
    """callback(*X, it=), utils.Traceback and StopIteration through the fused pgm / adaprox / bsdmm loops"""
    want = load_golden("nmf_callbacks")
    got = cases.nmf_callbacks(product)
    for k in ("pgm_trace_len", "bsdmm_trace_len"):
        assert int(got[k]) == int(want[k]), k
    for k in ("pgm_stop_its", "ada_stop_its"):
        assert list(got[k]) == list(want[k]), k
    for k in ("pgm_trace3_A", "pgm_trace3_S", "pgm_A", "pgm_S", "pgm_stop_A", "ada_stop_A", "ada_stop_S", "ada_trace2_S"):
        assert_close(got[k], want[k], 1e-4, k)
    for k in ("bsdmm_trace2_A", "bsdmm_A"):
        assert_close(got[k], want[k], 5e-4, k)


def test_admm_callback_unstarred(product):
    """admm / sdmm call callback(X, it=it) with the array itself (algorithms.py:480, 605)"""
    want = load_golden("admm_callback_unstarred")
    got = cases.admm_callback_unstarred(product)
    assert np.array_equal(got["seen"], want["seen"])
    assert_close(got["X"], want["X"], 1e-6, "X")
    assert_close(got["X2"], want["X2"], 1e-6, "X2")


# ---------------------------------------------------------------- horizons of SURVEY 7.3 / full-size stripe, oracle run live
def _oracle_vs_product(product, Y, A0, S0, **kw):
    orc = apis.oracle()
    Ao, So = A0.copy(), S0.copy()
    ro = orc.nmf.nmf(Y, Ao, So, **{k: (getattr(orc, v[1:]) if isinstance(v, str) and v.startswith("@") else v)
                                   for k, v in kw.items()})
    no = orc.iterations()
    Ap, Sp = A0.copy(), S0.copy()
    rp = product.nmf.nmf(Y, Ap, Sp, **{k: (getattr(product, v[1:]) if isinstance(v, str) and v.startswith("@") else v)
                                       for k, v in kw.items()})
    return (Ao, So, no, ro), (Ap, Sp, product.iterations(), rp)


def test_cfg2_full_rows_stripe_vs_oracle(product):
    """BASELINE config 2 at its own M and K: a 4096-column stripe of the 8192 x 65536 problem (the fused kernel is
    stripe-local), 3 PGM iterations plus / unity_plus against the oracle run live; factors <= 1e-4"""
    from proxmin_b200 import workloads

    Y, A0, S0 = workloads.cfg2(8192, 4096, 64, seed=1234)
    (Ao, So, no, _), (Ap, Sp, npd, _) = _oracle_vs_product(product, Y, A0, S0, prox_A="@prox_plus",
                                                           prox_S="@prox_unity_plus", max_iter=3, e_rel=0)
    assert no[0] == npd[0] == 3
    assert_close(Ap, Ao, 1e-4, "A")
    assert_close(Sp, So, 1e-4, "S")
    assert_same_support(Sp, So, "S", guard=1e-4)


def test_pgm_horizon_1024x8192_100it(product):
    """SURVEY 7.3 horizon for the config-2 recipe: 1024 x 8192, K = 64, 100 iterations, plus / unity_plus"""
    from proxmin_b200 import workloads

    Y, A0, S0 = workloads.cfg2(1024, 8192, 64, seed=77)
    (Ao, So, no, _), (Ap, Sp, npd, _) = _oracle_vs_product(product, Y, A0, S0, prox_A="@prox_plus",
                                                           prox_S="@prox_unity_plus", max_iter=100, e_rel=0)
    assert no[0] == npd[0] == 100
    assert_close(Ap, Ao, 1e-4, "A")
    assert_close(Sp, So, 1e-4, "S")


def test_amsgrad_horizon_1024x8192_100it(product):
    """SURVEY 7.3 horizon for config 3: adaprox / AMSGrad, plus / plus, 100 iterations; sub-iteration counts equal"""
    from proxmin_b200 import workloads

    Y, A0, S0 = workloads.cfg2(1024, 8192, 64, seed=78)
    (Ao, So, no, _), (Ap, Sp, npd, _) = _oracle_vs_product(product, Y, A0, S0, algorithm="@adaprox", scheme="amsgrad",
                                                           max_iter=100, check_convergence=False)
    assert no[0] == npd[0] == 100
    assert list(no[1]) == list(npd[1]), (no, npd)
    assert_close(Ap, Ao, 1e-4, "A")
    assert_close(Sp, So, 1e-4, "S")


# ---------------------------------------------------------------- ADMM / SDMM
def test_admm_lasso_callback_loop(product):
    """plain closures for prox_f / step_f: callback loop, ADMM variable updates on the device"""
    want = load_golden("admm_lasso")
    got = cases.admm_lasso(product)
    assert int(got["iterations"]) == int(want["iterations"])
    assert bool(got["converged"]) == bool(want["converged"])
    assert_close(got["X"], want["X"], 1e-6, "X")
    assert_same_support(got["X"], want["X"], "X")


def test_admm_lasso_fused_bit_exact(product):
    """LeastSquaresProx + ConstantStep + built-in prox_g: the fused device loop.  The kernel keeps NumPy's
    operation order without FMA contraction, so X is bit-identical to the reference's fp32 result."""
    from functools import partial

    import proxmin_b200 as pmx
    from proxmin_b200 import workloads

    want = load_golden("admm_lasso")
    b, X = workloads.cfg4(10_000, seed=7)
    conv, err = pmx.admm(X, pmx.utils.LeastSquaresProx(b), pmx.utils.ConstantStep(0.5),
                         prox_g=partial(pmx.prox_soft, thresh=0.5), max_iter=200, e_rel=1e-6)
    n, _ = product.iterations()
    assert n == int(want["iterations"])
    assert bool(conv) == bool(want["converged"])
    assert np.array_equal(X, want["X"])
    assert np.array_equal(np.signbit(X), np.signbit(want["X"]))
    assert np.allclose(np.array(err, dtype=np.float64), want["errors"], rtol=1e-5)


def test_sdmm_lasso(product):
    from functools import partial

    import proxmin_b200 as pmx
    from proxmin_b200 import workloads

    want = load_golden("sdmm_lasso_plus")
    got = cases.sdmm_lasso_plus(product)          # callback loop
    assert int(got["iterations"]) == int(want["iterations"])
    assert_close(got["X"], want["X"], 1e-6, "X")
    b, X = workloads.cfg4(10_000, seed=8)          # fused device loop
    conv = pmx.sdmm(X, pmx.utils.LeastSquaresProx(b), pmx.utils.ConstantStep(0.5),
                    proxs_g=[partial(pmx.prox_soft, thresh=0.5), pmx.prox_plus], max_iter=100, e_rel=1e-5)
    n, _ = product.iterations()
    assert n == int(want["iterations"])
    assert bool(conv) == bool(want["converged"])
    assert np.array_equal(X, want["X"])


# ---------------------------------------------------------------- SURVEY 8-f rows
def test_prox_max_entropy(product):
    """operators.py:163-184 (Lambert W on the device in fp64, like scipy's).  Tolerance 2e-6 relative: the argument
    of W is an fp32 exp whose last bit may differ between NumPy and CUDA; untouched elements (X <= 0, NaN) and the
    fp32 overflow to inf are exact."""
    want = load_golden("operators_max_entropy")
    with np.errstate(over="ignore", invalid="ignore"):
        got = cases.operators_max_entropy(product)
    for k in want:
        w, g = np.asarray(want[k]), np.asarray(got[k])
        assert g.dtype == w.dtype, k
        untouched = ~(w > 0) | ~np.isfinite(w)
        assert np.array_equal(g[untouched], w[untouched], equal_nan=True), k + " (untouched / overflow)"
        assert np.array_equal(np.signbit(g[untouched]), np.signbit(w[untouched])), k
        # float64 arrays travel through fp32 device buffers (parity is defined on fp32 inputs)
        assert np.allclose(g[~untouched], w[~untouched], rtol=2e-6 if k != "f64" else 3e-7 * 4, atol=0), k


def test_barzilai_borwein_stepper(product):
    """utils.py:209-241 through pgm: reductions on the device (fp32 products, fp64 accumulation vs NumPy's pairwise
    fp32 sums): the 25-iteration path agrees to 2e-5 relative, the last step to 5e-3."""
    want = load_golden("pgm_barzilai_borwein")
    got = cases.pgm_barzilai_borwein(product)
    for typ in (1, 2):
        assert_close(got["S%d" % typ], want["S%d" % typ], 2e-5, "S%d" % typ)
        assert_same_support(got["S%d" % typ], want["S%d" % typ], "S%d" % typ, guard=1e-4)
        # the BB step is a ratio of sums over DIFFERENCES of consecutive iterates / gradients: fp32 rounding noise of
        # the path is amplified ~100x in it
        assert np.allclose(got["step%d" % typ], want["step%d" % typ], rtol=5e-3), typ


def test_admm_family_dense_L(product):
    """admm / sdmm / bsdmm with non-identity dense linear operators (utils.py:38-101, 295-391): L X, L^T (..) and the
    spectral norm on the device.  fp32 GEMM summation order differs from OpenBLAS: 2e-5 relative on the iterates,
    iteration counts equal."""
    want = load_golden("admm_dense_L")
    got = cases.admm_dense_L(product)
    for k in ("admm_it", "sdmm_it", "bsdmm_it"):
        assert int(got[k]) == int(want[k]), k
    for k in ("admm_X", "sdmm_X", "bsdmm_X1", "bsdmm_X2"):
        assert_close(got[k], want[k], 2e-5, k)
    # (e_pri, e_dual, |R|, |S|): the residual norms at convergence are differences of nearly equal iterates
    assert np.allclose(got["admm_err"][:2], want["admm_err"][:2], rtol=1e-4), "tolerances"
    assert np.allclose(got["admm_err"][2:], want["admm_err"][2:], rtol=2e-2), "residual norms"


def test_matrix_adapter(product):
    """utils.MatrixAdapter: None is the identity (argument returned uncopied), dense dot / T.dot / spectral norm"""
    import proxmin_b200 as pmx

    rng = np.random.default_rng(5)
    L = rng.standard_normal((37, 150)).astype(np.float32)
    x = rng.standard_normal(150).astype(np.float32)
    X2 = rng.standard_normal((150, 9)).astype(np.float32)
    ad = pmx.utils.MatrixAdapter(L)
    assert np.allclose(ad.dot(x), L.dot(x), rtol=2e-5, atol=2e-5)
    assert np.allclose(ad.dot(X2), L.dot(X2), rtol=2e-5, atol=2e-5)
    y = rng.standard_normal(37).astype(np.float32)
    assert np.allclose(ad.T.dot(y), L.T.dot(y), rtol=2e-5, atol=2e-5)
    lam = np.linalg.eigvalsh(L.astype(np.float64) @ L.astype(np.float64).T).max()
    assert abs(float(ad.spectral_norm) - lam) <= 5e-6 * lam
    big = rng.standard_normal((200, 300)).astype(np.float32)       # both dimensions > 128: power iteration
    lam = np.linalg.eigvalsh(big.astype(np.float64) @ big.astype(np.float64).T).max()
    assert abs(float(pmx.utils.get_spectral_norm(big)) - lam) <= 1e-4 * lam
    ident = pmx.utils.MatrixAdapter(None)
    assert ident.dot(x) is x and ident.T is ident and ident.spectral_norm == 1


def test_tiled_upload_in_column_blocks():
    """pmx_nmf_set_Y builds the tiled device copy of Y from any column range (the Python front end uploads fp64 /
    non-contiguous input in column blocks): three ragged ranges must give the same gradients as one upload, for the
    weight matrix too."""
    import ctypes as C

    from proxmin_b200 import _ffi, workloads
    from proxmin_b200 import nmf as pnmf

    Y, A, S = workloads.cfg2(200, 517, 12, seed=4)
    W = (0.5 + np.random.default_rng(1).random(Y.shape)).astype(np.float32)
    ref = pnmf.Problem(Y, A, S, W=W)
    ref.gradient()
    GA0, GS0 = ref.get(_ffi.GA), ref.get(_ffi.GS)
    ref.close()
    prob = pnmf.Problem(np.zeros_like(Y), A, S, W=np.ones_like(W))
    L = _ffi.lib()
    for c0, c1 in ((0, 131), (131, 400), (400, 517)):
        for setter, src in ((L.pmx_nmf_set_Y, Y), (L.pmx_nmf_set_W, W)):
            part = np.ascontiguousarray(src[:, c0:c1])
            _ffi.check(setter(prob.handle, part.ctypes.data_as(C.c_void_p), c1 - c0, c0, c1 - c0))
    prob.gradient()
    # (the gradients are accumulated with red.add in an order that varies from launch to launch: last-bit differences)
    assert np.allclose(prob.get(_ffi.GA), GA0, rtol=2e-6, atol=1e-5) and np.allclose(prob.get(_ffi.GS), GS0, rtol=2e-6, atol=1e-5)
    prob.close()
    Y64 = np.asfortranarray(Y.astype(np.float64))       # fp64, non-contiguous: the column-block path of Problem
    g64 = pnmf.grad_likelihood(A.astype(np.float64), S.astype(np.float64), Y=Y64, W=W.astype(np.float64))
    assert g64[0].dtype == np.float64 and np.allclose(g64[0], GA0, rtol=1e-5, atol=1e-4)


def test_utils_admm_primitives(product):
    """utils.update_variables / do_the_mm / get_variable_errors / check_constraint_convergence / check_convergence
    (utils.py:295-406) as public functions: one ADMM step with a dense L against the same NumPy expressions."""
    from functools import partial

    import proxmin_b200 as pmx

    rng = np.random.default_rng(9)
    L = (rng.standard_normal((10, 16)) / 4).astype(np.float32)
    b = rng.standard_normal(16).astype(np.float32)
    X0 = rng.standard_normal(16).astype(np.float32)
    prox_f = lambda X, s: X - s * (X - b)          # noqa: E731
    prox_g = partial(pmx.prox_soft, thresh=0.1)
    sf, sg = np.float32(0.5), np.float32(0.7)
    # reference expressions (utils.py:295-346) in NumPy
    Xr = X0.copy()
    Zr = L.dot(Xr).copy()
    Ur = np.zeros_like(Zr)
    Ur += 0.05
    dX = sf / sg * L.T.dot(L.dot(Xr) - Zr + Ur)
    Xr[:] = prox_f(Xr - dX, sf)
    LXr = L.dot(Xr)
    t = LXr + Ur
    Zn = np.sign(t) * np.maximum(np.abs(t) - 0.1 * sg, 0)
    Rr = LXr - Zn
    Sr = -1 / sg * L.T.dot(Zn - Zr)
    Ur2 = Ur + Rr
    # product
    Xp = X0.copy()
    ad = pmx.utils.MatrixAdapter(L)
    Zp, Up = pmx.utils.initZU(Xp, ad)
    Up += 0.05
    LXp, Rp, Sp = pmx.utils.update_variables(Xp, Zp, Up, prox_f, sf, prox_g, sg, ad)
    for got, want, name in ((Xp, Xr, "X"), (LXp, LXr, "LX"), (Zp, Zn, "Z"), (Up, Ur2, "U"), (Rp, Rr, "R"), (Sp, Sr, "S")):
        assert np.allclose(got, want, rtol=2e-5, atol=2e-6), name
    conv, (e_pri, e_dual, lR, lS) = pmx.utils.check_constraint_convergence(Xp, ad, LXp, Zp, Up, Rp, Sp, sf, sg, 1e-3, 0.01)
    spec = np.linalg.eigvalsh(L.T.astype(np.float64) @ L.astype(np.float64)).max()
    l2 = lambda x: np.sqrt((x.astype(np.float64) ** 2).sum())   # noqa: E731
    assert np.isclose(e_pri, np.sqrt(Zn.size) * 0.01 / spec + 1e-3 * max(l2(LXr), l2(Zn)), rtol=1e-4)
    assert np.isclose(e_dual, np.sqrt(Xr.size) * 0.01 / spec + 1e-3 * l2(L.T.dot(Ur2) / sg), rtol=1e-4)
    assert np.isclose(lR, l2(Rr), rtol=1e-4) and np.isclose(lS, l2(Sr), rtol=1e-4)
    assert conv == bool(l2(Rr) <= e_pri and l2(Sr) <= e_dual)
    new, old = rng.random((4, 9)).astype(np.float32), rng.random((4, 9)).astype(np.float32)
    c, norms = pmx.utils.check_convergence(new, old, 0.5)
    assert np.allclose(norms, [np.sum(new * old), np.sum(old ** 2)], rtol=1e-5)
    assert c == bool(np.sum(new * old) >= (1 - 0.25) * np.sum(old ** 2))


def test_nmf_weighted_likelihood(product):
    """Weighted likelihood W (nmf.py:25, 40): gradient / loss, fused adaprox loop, PGM with a user step (callback
    loop, Y and W resident on the device)"""
    want = load_golden("nmf_weighted")
    got = cases.nmf_weighted(product)
    assert_close(got["G_A"], want["G_A"], 2e-5, "G_A")
    assert_close(got["G_S"], want["G_S"], 2e-5, "G_S")
    assert abs(float(got["loss"]) - float(want["loss"])) <= 2e-5 * abs(float(want["loss"]))
    assert np.array_equal(got["ada_sub"], want["ada_sub"])
    for k in ("ada_A", "ada_S", "pgm_A", "pgm_S"):
        assert_close(got[k], want[k], 1e-4, k)
        assert_same_support(got[k], want[k], k, guard=1e-4)
    import proxmin_b200 as pmx
    from proxmin_b200 import workloads
    Y, A, S = workloads.cfg2(64, 128, 4, seed=1)
    W = np.ones_like(Y)
    with pytest.raises(ValueError):          # the reference's `if W == 1` on an array (nmf.py:63)
        pmx.nmf.nmf(Y, A, S, W=W, max_iter=2)
    with pytest.raises(ValueError):
        pmx.nmf.nmf(Y, A, S, W=W, algorithm=pmx.bsdmm, max_iter=2)


# ---------------------------------------------------------------- tcgen05 kernel vs the SIMT kernel
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 512, 8), (300, 1000, 20), (77, 204, 5), (1024, 2048, 64), (130, 333, 7),
                                   (257, 129, 33), (256, 512, 96), (300, 700, 128), (130, 333, 65), (1024, 2048, 128)])
def test_tcgen05_gradient_matches_fp64(shape):
    """3xBF16-split tensor-core GEMMs: gradients within 2e-5 (relative Frobenius) of an fp64 evaluation"""
    import ctypes as C

    from proxmin_b200 import _ffi

    M, N, K = shape
    rng = np.random.default_rng(123)
    A = rng.random((M, K), dtype=np.float32)
    S = rng.random((K, N), dtype=np.float32)
    Y = (rng.random((M, K), dtype=np.float32) @ rng.random((K, N), dtype=np.float32)).astype(np.float32)
    R = A.astype(np.float64) @ S.astype(np.float64) - Y
    ctx = _ffi.context()
    dY, dA, dS = ctx.upload(Y), ctx.upload(A), ctx.upload(S)
    dGA, dGS, dl = ctx.malloc(4 * M * K), ctx.malloc(4 * K * N), ctx.malloc(16)
    try:
        _ffi.check(_ffi.lib().pmx_nmf_grad(ctx.handle, dY, dA, dS, M, N, K, dGA, dGS, dl, 2))
        GA, GS, loss = np.empty((M, K), np.float32), np.empty((K, N), np.float32), np.empty(2, np.float64)
        ctx.d2h(GA, dGA)
        ctx.d2h(GS, dGS)
        ctx.d2h(loss, dl)
    finally:
        for p in (dY, dA, dS, dGA, dGS, dl):
            ctx.free(p)
    assert rel_fro(GA, R @ S.T.astype(np.float64)) < 2e-5
    assert rel_fro(GS, A.T.astype(np.float64) @ R) < 2e-5
    assert abs(loss[0] - 0.5 * (R ** 2).sum()) <= 5e-6 * 0.5 * (R ** 2).sum()


def test_full_size_properties():
    """BASELINE config 2 shape (8192 x 65536, K = 64): size-independent properties of the fused path.

    * linearity of the gradient in Y:  G(Y1) + G(Y2) - G(0) == G(Y1 + Y2)   (G is affine in Y)
    * prox_unity_plus idempotence: column sums of S are 1 and S >= 0 after every iteration
    * the loss decreases monotonically under the Lipschitz steps"""
    import proxmin_b200 as pmx
    from proxmin_b200 import _ffi
    from proxmin_b200 import nmf as pnmf

    M, N, K = 8192, 65536, 64
    rng = np.random.default_rng(5)
    A0 = rng.random((M, K), dtype=np.float32)
    S0 = rng.random((K, N), dtype=np.float32)
    Y = A0[:, :8] @ rng.random((8, N), dtype=np.float32)
    prob = pnmf.Problem(Y, A0, S0)
    try:
        chain_A = [(_ffi.OP_PLUS, 0, 0, 0.0)]
        chain_S = [(_ffi.OP_PLUS, 0, 0, 0.0), (_ffi.OP_UNITY, 0, 0, 0.0)]
        prob.pgm_begin(chain_A, chain_S, e_rel=(0.0, 0.0))
        losses = [prob.loss()]
        for _ in range(3):
            prob.pgm_run(2)
            losses.append(prob.loss())
        S = prob.get(_ffi.S)
        A = prob.get(_ffi.A)
    finally:
        prob.close()
    assert np.all(np.diff(losses) < 0), losses
    assert S.min() >= 0 and A.min() >= 0
    assert np.allclose(S.sum(axis=0), 1.0, atol=1e-5)


@pytest.mark.gpu
def test_multi_gpu_sharded():
    """Column-sharded PGM / adaprox / bsdmm over NCCL against the oracle (tests/mgpu_check.py); needs >= 2 GPUs."""
    import subprocess
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(os.path.dirname(__file__), "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_cached_buffers_are_clean_between_solves():
    """Solver handles reuse device blocks of destroyed handles (pmx_dev_alloc): a second solve must not see state
    of the first one, and pmx_ctx_trim must leave the context usable."""
    from proxmin_b200 import _ffi, workloads
    from proxmin_b200 import nmf as pnmf
    import proxmin_b200 as pmx

    def solve(seed):
        Y, A, S = workloads.cfg2(256, 1024, 16, seed=seed)
        pnmf.nmf(Y, A, S, prox_A=pmx.prox_plus, prox_S=pmx.prox_unity_plus, max_iter=20, e_rel=0)
        return A, S

    A1, S1 = solve(3)
    A2, S2 = solve(4)          # same shapes: takes the cached blocks of the first handle
    A3, S3 = solve(3)
    # (the gradient accumulates with floating-point reductions whose order varies: equal to rounding, not bitwise)
    close = lambda a, b: np.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert close(A1, A3) and close(S1, S3)
    assert not close(A1, A2)
    _ffi.context().trim()
    A4, S4 = solve(3)
    assert close(A1, A4) and close(S1, S4)
