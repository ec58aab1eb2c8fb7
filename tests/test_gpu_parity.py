"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed
golden vectors of the unmodified reference, on the same seeded inputs.

Tolerances (stated per test): the device computes in fp32 with 3xBF16-split tensor-core GEMMs
(or fp32 SIMT); the north-star bound is 1e-4 relative on the factors; support sets of
plus/hard/soft/max/min outputs must be identical wherever the reference value is not within
1e-6*max|X| of the threshold; iteration and sub-iteration counts must match where stated.
"""
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


PGM_CASES = ["nmf_pgm_cfg1", "nmf_pgm_unity", "nmf_pgm_altproj", "nmf_pgm_soft", "nmf_pgm_ragged", "nmf_pgm_accel"]


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


def test_oracle_agrees_live(product):
    """same seeded inputs, oracle run live next to the device (not just the stored vectors)"""
    orc = apis.oracle()
    a = cases.nmf_pgm_ragged(orc)
    b = cases.nmf_pgm_ragged(product)
    assert_close(b["A"], a["A"], 1e-4, "A")
    assert_close(b["S"], a["S"], 1e-4, "S")
