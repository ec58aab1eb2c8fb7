"""Column-sharded solvers on N GPUs against the same problem solved by the oracle.

Run under torchrun on a multi-GPU box (tests/test_gpu_parity.py::test_multi_gpu_sharded does so when >= 2 GPUs
are visible):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/mgpu_check.py

Every rank builds the SAME seeded problem, keeps the columns workloads.shard_columns gives it and calls the
public ``nmf.nmf`` on its stripe (Y, S sharded; A replicated; G_A, the Gram matrix of S, the S-block norms,
max(Psi) and the row sums of S all-reduced over NCCL).  Rank 0 solves the whole problem with the CPU oracle
(the checker of every other GPU test) and compares the gathered factors and the iteration counts.
"""
import os
import sys
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import apis  # noqa: E402
import proxmin_b200 as pmx  # noqa: E402
from proxmin_b200 import _ffi, workloads  # noqa: E402


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _ffi.context()
    box = [ctx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    pmx.init_distributed(box[0], world, rank)
    prod, orc = apis.product(), apis.oracle()
    fails = []

    def run(name, make, solve, tol, counts=True):
        """make() -> (Y, A, S); solve(api, Y, A, S) updates A, S in place."""
        Y, A, S = make()
        N = Y.shape[1]
        lo, hi = workloads.shard_columns(N, world, rank)
        A_d, S_d = A.copy(), np.ascontiguousarray(S[:, lo:hi])
        solve(prod, np.ascontiguousarray(Y[:, lo:hi]), A_d, S_d)
        it_d = prod.iterations()
        parts = [None] * world
        dist.all_gather_object(parts, S_d)
        As = [None] * world
        dist.all_gather_object(As, A_d)
        if rank != 0:
            return
        S_all = np.concatenate(parts, axis=1)
        A_o, S_o = A.copy(), S.copy()
        solve(orc, Y, A_o, S_o)
        it_o = orc.iterations()
        eA, eS = relerr(A_d, A_o), relerr(S_all, S_o)
        rep = max(relerr(a, A_d) for a in As)          # A must be identical on all ranks
        ok = eA <= tol and eS <= tol and rep == 0.0 and (not counts or it_d == it_o)
        print("%-26s A %.2e S %.2e (tol %.0e) replicas %.1e iterations %s vs %s %s"
              % (name, eA, eS, tol, rep, it_d, it_o, "ok" if ok else "FAIL"), flush=True)
        if not ok:
            fails.append(name)

    def pgm_unity(api, Y, A, S):
        api.nmf.nmf(Y, A, S, prox_A=api.prox_plus, prox_S=api.prox_unity_plus, max_iter=40, e_rel=0)

    def pgm_converge(api, Y, A, S):
        api.nmf.nmf(Y, A, S, max_iter=1000)

    def ada_ams(api, Y, A, S):
        api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", max_iter=30, check_convergence=False)

    def ada_adam(api, Y, A, S):
        api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="adam", max_iter=25, e_rel=1e-3)

    def ada_unity(api, Y, A, S):
        api.nmf.nmf(Y, A, S, algorithm=api.adaprox, scheme="amsgrad", prox_S=api.prox_unity_plus, max_iter=30,
                    check_convergence=False)

    def bsdmm(api, Y, A, S):
        proxs_g = [[api.prox_plus, api.prox_unity], [api.prox_plus, partial(api.prox_soft, thresh=0.01)]]
        api.nmf.nmf(Y, A, S, algorithm=api.bsdmm, prox_A=api.prox_id, prox_S=api.prox_id, proxs_g=proxs_g,
                    max_iter=6, e_rel=1e-6)

    run("pgm plus/unity_plus", lambda: workloads.cfg2(128, 384 + 72, 16, seed=5), pgm_unity, 1e-4)
    run("pgm converge (cfg1)", lambda: workloads.cfg1(), pgm_converge, 1e-4)
    run("pgm K=64 tcgen05", lambda: workloads.cfg2(512, 2048, 64, seed=8), pgm_unity, 1e-4)
    run("adaprox amsgrad", lambda: workloads.cfg2(128, 384 + 72, 16, seed=6), ada_ams, 2e-4)
    run("adaprox amsgrad unity", lambda: workloads.cfg2(128, 384, 16, seed=6), ada_unity, 2e-4)
    run("adaprox adam converge", lambda: workloads.cfg2(64, 192 + 40, 8, seed=7), ada_adam, 2e-4)
    run("bsdmm", lambda: workloads.cfg5(96, 256 + 24, 8, seed=9), bsdmm, 5e-4)

    dist.barrier()
    if rank == 0:
        print("FAILS", fails, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
