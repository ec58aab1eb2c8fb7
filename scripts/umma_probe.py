"""GPU probe: tcgen05 gradient kernel vs the SIMT kernel and NumPy fp64 on several shapes."""
import sys

import numpy as np

sys.path.insert(0, ".")
from proxmin_b200 import _ffi  # noqa: E402


def grad(ctx, dY, dA, dS, M, N, K, kernel):
    L = _ffi.lib()
    dGA, dGS, dl = ctx.malloc(4 * M * K), ctx.malloc(4 * K * N), ctx.malloc(16)
    _ffi.check(L.pmx_nmf_grad(ctx.handle, dY, dA, dS, M, N, K, dGA, dGS, dl, kernel))
    ctx.sync()
    GA, GS, loss = np.empty((M, K), np.float32), np.empty((K, N), np.float32), np.empty(2, np.float64)
    ctx.d2h(GA, dGA)
    ctx.d2h(GS, dGS)
    ctx.d2h(loss, dl)
    for p in (dGA, dGS, dl):
        ctx.free(p)
    return GA, GS, loss[0]


def rel(a, b):
    return np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300)


def main():
    ctx = _ffi.context()
    print(ctx.device_info())
    shapes = [(128, 128, 64), (128, 256, 64), (256, 128, 64), (256, 512, 8), (1024, 2048, 64), (300, 1000, 20),
              (77, 204, 5), (2048, 16384, 64)]
    rng = np.random.default_rng(0)
    for (M, N, K) in shapes:
        A = rng.random((M, K), dtype=np.float32)
        S = rng.random((K, N), dtype=np.float32)
        Y = (rng.random((M, K), dtype=np.float32) @ rng.random((K, N), dtype=np.float32)).astype(np.float32)
        dY, dA, dS = ctx.upload(Y), ctx.upload(A), ctx.upload(S)
        R = A.astype(np.float64) @ S.astype(np.float64) - Y
        GA64, GS64, l64 = R @ S.T.astype(np.float64), A.T.astype(np.float64) @ R, 0.5 * (R ** 2).sum()
        out = {}
        for kern in (1, 2):
            GA, GS, loss = grad(ctx, dY, dA, dS, M, N, K, kern)
            out[kern] = (rel(GA, GA64), rel(GS, GS64), abs(loss - l64) / l64)
        print("M=%d N=%d K=%d  simt: GA %.2e GS %.2e loss %.2e | umma: GA %.2e GS %.2e loss %.2e" %
              ((M, N, K) + out[1] + out[2]), flush=True)
        for p in (dY, dA, dS):
            ctx.free(p)


if __name__ == "__main__":
    main()
