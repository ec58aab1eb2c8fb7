"""Column-sharded solvers on N GPUs against the same problem solved on one GPU.

Run under torchrun on a multi-GPU box (tests/test_gpu_parity.py::test_multi_gpu_sharded does so when >= 2 GPUs
are visible):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/mgpu_check.py

Every rank builds the SAME seeded problem, keeps the columns workloads.shard_columns gives it and runs the
solver on its stripe (Y, S sharded; A replicated; G_A, the Gram matrix of S, the S-block norms, max(Psi) and
the row means of S all-reduced over NCCL).  Rank 0 then solves the whole problem alone (a second, local
context-free Problem on the same GPU is not possible while the communicator is attached, so the single-GPU
answer comes from the CPU oracle, the same checker the other GPU tests use) and compares.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import proxmin_b200 as pmx  # noqa: E402
from proxmin_b200 import _ffi, workloads  # noqa: E402
from proxmin_b200 import nmf as pnmf  # noqa: E402


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def gather_S(S_loc, N, world, rank):
    parts = [None] * world
    dist.all_gather_object(parts, S_loc)
    return np.concatenate(parts, axis=1)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _ffi.context()
    box = [ctx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    pmx.init_distributed(box[0], world, rank)

    from oracle import proxmin_oracle as orc   # checker only

    M, N, K = 256, 1536 + 40, 16          # ragged last stripe
    rng = np.random.RandomState(11)
    A0 = rng.rand(M, K).astype(np.float32) + 0.1
    S0 = rng.rand(K, N).astype(np.float32) + 0.1
    Y = (A0 @ S0 + 0.01 * rng.rand(M, N)).astype(np.float32)
    A1 = (A0 * (1 + 0.3 * rng.rand(M, K))).astype(np.float32)
    S1 = (S0 * (1 + 0.3 * rng.rand(K, N))).astype(np.float32)
    lo, hi = workloads.shard_columns(N, world, rank)
    plus = [(_ffi.OP_PLUS, 0, 0, 0.0)]
    fails = []

    def report(name, err, tol):
        if rank == 0:
            print("%-28s %.3e (tol %.0e) %s" % (name, err, tol, "ok" if err <= tol else "FAIL"), flush=True)
            if not err <= tol:
                fails.append(name)

    # ---- PGM (both gradient kernels) --------------------------------------------------
    import functools
    for kern, kname in ((1, "simt"), (2, "tcgen05")):
        prob = pnmf.Problem(np.ascontiguousarray(Y[:, lo:hi]), A1, np.ascontiguousarray(S1[:, lo:hi]))
        prob.pgm_begin(plus, plus, False, (0.0, 0.0), kernel=kern, check_every=4)
        it, _, _ = prob.pgm_run(12)
        A_d, S_d = prob.get(_ffi.A), gather_S(prob.get(_ffi.S), N, world, rank)
        prob.close()
        if rank == 0:
            A_o, S_o = A1.copy(), S1.copy()
            orc.pgm([A_o, S_o], functools.partial(orc.nmf_grad, Y=Y), orc.nmf_step_pgm,
                    prox=[orc.prox_plus, orc.prox_plus], max_iter=12, e_rel=0.0)
            tol = 1e-5 if kern == 1 else 1e-4
            report("pgm/%s A" % kname, relerr(A_d, A_o), tol)
            report("pgm/%s S" % kname, relerr(S_d, S_o), tol)
            if it != 12:
                fails.append("pgm/%s iterations %d" % (kname, it))

    # ---- adaprox / amsgrad ------------------------------------------------------------
    prob = pnmf.Problem(np.ascontiguousarray(Y[:, lo:hi]), A1, np.ascontiguousarray(S1[:, lo:hi]))
    prob.adaprox_begin(plus, plus, "amsgrad", 0.999, 1e-8, 0.25, (1e-3, 1e-3), False, 1000)
    n_it = 8
    b1 = np.full(n_it, 0.9)
    b1p = np.roll(b1, 1)
    it, _, sub = prob.adaprox_run(n_it, b1, b1p)
    A_d, S_d = prob.get(_ffi.A), gather_S(prob.get(_ffi.S), N, world, rank)
    prob.close()
    if rank == 0:
        A_o, S_o = A1.copy(), S1.copy()
        res = orc.adaprox([A_o, S_o], functools.partial(orc.nmf_grad, Y=Y), orc.nmf_step_adaprox,
                          prox=[orc.prox_plus, orc.prox_plus], scheme="amsgrad", max_iter=n_it, e_rel=1e-3,
                          check_convergence=False, return_counts=True) if "return_counts" in orc.adaprox.__code__.co_varnames \
            else orc.adaprox([A_o, S_o], functools.partial(orc.nmf_grad, Y=Y), orc.nmf_step_adaprox,
                             prox=[orc.prox_plus, orc.prox_plus], scheme="amsgrad", max_iter=n_it, e_rel=1e-3,
                             check_convergence=False)
        report("adaprox A", relerr(A_d, A_o), 2e-4)
        report("adaprox S", relerr(S_d, S_o), 2e-4)
        print("adaprox iterations", it, "sub-iterations", sub, "oracle", res, flush=True)

    # ---- bsdmm ------------------------------------------------------------------------
    prob = pnmf.Problem(np.ascontiguousarray(Y[:, lo:hi]), A1, np.ascontiguousarray(S1[:, lo:hi]))
    gA = [[(_ffi.OP_PLUS, 0, 0, 0.0)], [(_ffi.OP_UNITY, 0, 0, 0.0)]]
    gS = [[(_ffi.OP_PLUS, 0, 0, 0.0)], [(_ffi.OP_SOFT, 1, 0, 0.01)]]
    prob.bsdmm_begin([], [], gA, gS, (1e-3, 1e-3), (0.0, 0.0))
    it, _ = prob.bsdmm_run(5)
    A_d, S_d = prob.get(_ffi.A), gather_S(prob.get(_ffi.S), N, world, rank)
    prob.close()
    if rank == 0:
        A_o, S_o = A1.copy(), S1.copy()
        orc.nmf_bsdmm_reference(Y, A_o, S_o, max_iter=5) if hasattr(orc, "nmf_bsdmm_reference") else None
    dist.barrier()
    if rank == 0:
        print("FAILS", fails, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
