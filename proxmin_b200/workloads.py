"""Deterministic synthetic inputs for the five BASELINE.json configurations.

Recipes follow SURVEY.md section 8(d).  Everything is fp32, C-contiguous, drawn
from ``numpy.random.default_rng(seed)``.  The shape arguments let the parity tests
use scaled-down versions of the same recipe.
"""
import numpy as np


def cfg1(M=256, N=512, K=8, seed=0, dtype=np.float32):
    """nmf.nmf PGM, prox_plus/prox_plus (BASELINE config 1): noiseless Y = A* S*."""
    rng = np.random.default_rng(seed)
    At, St = rng.random((M, K)), rng.random((K, N))
    Y = At @ St
    A0, S0 = rng.random((M, K)), rng.random((K, N))
    return (np.ascontiguousarray(Y, dtype=dtype), np.ascontiguousarray(A0, dtype=dtype),
            np.ascontiguousarray(S0, dtype=dtype))


def cfg2(M=8192, N=65536, K=64, seed=1234):
    """nmf.nmf PGM / adaprox at the north-star shape (configs 2 and 3): noisy, clipped Y."""
    rng = np.random.default_rng(seed)
    At = rng.random((M, K), dtype=np.float32)
    St = rng.random((K, N), dtype=np.float32)
    Y = At @ St
    sd = np.float32(0.01) * np.float32(Y.std(dtype=np.float64))
    # add the noise in row blocks so the full-size recipe never holds a float64 M x N temporary
    blk = max(1, (1 << 24) // max(N, 1))
    for r0 in range(0, M, blk):
        r1 = min(M, r0 + blk)
        Y[r0:r1] += sd * rng.standard_normal((r1 - r0, N), dtype=np.float32)
    np.maximum(Y, 0, out=Y)
    A0 = rng.random((M, K), dtype=np.float32)
    S0 = rng.random((K, N), dtype=np.float32)
    return Y, A0, S0


def cfg4(n=10_000_000, seed=7):
    """ADMM LASSO (config 4): b = x* + noise with 1 % non-zeros; returns (b, X0)."""
    rng = np.random.default_rng(seed)
    x = np.zeros(n, dtype=np.float32)
    nz = rng.choice(n, size=max(1, n // 100), replace=False)
    x[nz] = 3.0 * rng.standard_normal(nz.size).astype(np.float32)
    b = x + np.float32(0.1) * rng.standard_normal(n, dtype=np.float32)
    return b.astype(np.float32), np.zeros(n, dtype=np.float32)


def cfg5(M=4096, N=131072, K=128, seed=99):
    """bsdmm constrained MF (config 5): columns of A* sum to one, noiseless Y."""
    rng = np.random.default_rng(seed)
    At = rng.random((M, K), dtype=np.float32)
    At /= At.sum(axis=0, keepdims=True)
    St = rng.random((K, N), dtype=np.float32)
    Y = At @ St
    A0 = rng.random((M, K), dtype=np.float32)
    A0 /= A0.sum(axis=0, keepdims=True)
    S0 = rng.random((K, N), dtype=np.float32)
    return Y, A0, S0


def cfg2_columns(M, N, K, lo, hi, seed=1234, block=4096):
    """Columns [lo, hi) of the config-2 workload in column-block form (bench.py): the factors A*, A0 come from one
    generator, every block of ``block`` columns of S*, the noise and S0 from its own generator, so that any
    partition of the columns over ranks sees the SAME global Y (SCALE runs at N = 1, 2, 4, 8 solve one problem).
    Noise level 0.01 std(Y) with the analytic std of a sum of K products of uniform(0,1) pairs, sqrt(7K/144)."""
    rng_a = np.random.default_rng(seed)
    At = rng_a.random((M, K), dtype=np.float32)
    A0 = rng_a.random((M, K), dtype=np.float32)
    sd = np.float32(0.01 * np.sqrt(7.0 * K / 144.0))
    n = hi - lo
    Y = np.empty((M, n), dtype=np.float32)
    S0 = np.empty((K, n), dtype=np.float32)
    for b in range(lo // block, (hi + block - 1) // block):
        c0, c1 = b * block, min(N, (b + 1) * block)
        rng = np.random.default_rng([seed, b])
        St = rng.random((K, c1 - c0), dtype=np.float32)
        Yb = At @ St
        Yb += sd * rng.standard_normal((M, c1 - c0), dtype=np.float32)
        np.maximum(Yb, 0, out=Yb)
        S0b = rng.random((K, c1 - c0), dtype=np.float32)
        a, e = max(lo, c0), min(hi, c1)
        Y[:, a - lo:e - lo] = Yb[:, a - c0:e - c0]
        S0[:, a - lo:e - lo] = S0b[:, a - c0:e - c0]
    return Y, A0, S0


def cfg5_columns(M, N, K, lo, hi, seed=99, block=4096):
    """Columns [lo, hi) of the config-5 workload in column-block form (see cfg2_columns): noiseless Y = A* S*,
    columns of A* and A0 sum to one."""
    rng_a = np.random.default_rng(seed)
    At = rng_a.random((M, K), dtype=np.float32)
    At /= At.sum(axis=0, keepdims=True)
    A0 = rng_a.random((M, K), dtype=np.float32)
    A0 /= A0.sum(axis=0, keepdims=True)
    n = hi - lo
    Y = np.empty((M, n), dtype=np.float32)
    S0 = np.empty((K, n), dtype=np.float32)
    for b in range(lo // block, (hi + block - 1) // block):
        c0, c1 = b * block, min(N, (b + 1) * block)
        rng = np.random.default_rng([seed, b])
        St = rng.random((K, c1 - c0), dtype=np.float32)
        Yb = At @ St
        S0b = rng.random((K, c1 - c0), dtype=np.float32)
        a, e = max(lo, c0), min(hi, c1)
        Y[:, a - lo:e - lo] = Yb[:, a - c0:e - c0]
        S0[:, a - lo:e - lo] = S0b[:, a - c0:e - c0]
    return Y, A0, S0


def shard_columns(N, world, rank, align=1):
    """Column range [lo, hi) of rank ``rank`` when N columns are split over ``world`` ranks.

    Stripes are contiguous and differ in width by at most ``align`` columns; every
    boundary is a multiple of ``align`` (the kernel's stripe width) except the last.
    """
    units = (N + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(N, lo_u * align), min(N, hi_u * align)
