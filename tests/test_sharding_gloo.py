"""Multi-GPU host logic on CPU: world_size-2 gloo run of the column-sharded PGM iteration.

The device loop (nmf_solver.cu: pgm_enqueue_iteration) shards Y and S by columns, replicates A and
exchanges exactly two messages per iteration: sum(G_A partials) after the gradient, and ONE packed buffer
[S S^T partials of the new S (step_A of the next iteration) | the S-block norms] after the S update.  This test
replays that exchange sequence with gloo collectives around the oracle's
NumPy pieces and checks it against the unsharded oracle: it pins the partition (workloads.shard_columns)
and the list of reduced quantities that the device path relies on.  Two spellings of the exchange: an all-reduce
(the NCCL fallback) and an all-gather followed by a sum in rank order in the working precision -- what the
peer-memory kernels of comm.cu do -- for which the replicas of A must stay bit-identical on every rank -- and the
reduce-scatter / all-gather of the fused PGM tail of round 2 (mode "scatter": every rank sums and updates only its row slice
of A, the new rows are gathered, the Gram partials of both new factors are exchanged for the next steps)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, out, mode="allreduce"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import proxmin_oracle as orc
    from proxmin_b200 import workloads

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(x):
        if mode == "gather":   # k_peer_sum: every rank reads every partial and adds them in rank order, same dtype
            x = np.ascontiguousarray(x)
            parts = [torch.empty_like(torch.from_numpy(x)) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(x))
            acc = np.zeros_like(x)
            for p in parts:
                acc = acc + p.numpy()
            return acc
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    M, N, K = 96, 400, 6
    Y, A, S = workloads.cfg2(M, N, K, seed=3)
    lo, hi = workloads.shard_columns(N, world, rank, align=128)
    Yl, Sl = Y[:, lo:hi].copy(), S[:, lo:hi].copy()
    A = A.copy()
    iters = 25
    if mode == "scatter":
        # The fused PGM tail of round 2 (pgm_tail.cu): reduce-scatter of the G_A partials by row slice, the A update on
        # the slice only, all-gather of the new rows; the Gram partials of BOTH new factors (rows of A per rank, columns
        # of S per rank) summed in rank order give the steps of the next iteration.
        def gather_sum(x):
            x = np.ascontiguousarray(x)
            parts = [torch.empty_like(torch.from_numpy(x)) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(x))
            acc = np.zeros_like(x)
            for p in parts:
                acc = acc + p.numpy()
            return acc, parts

        m_lo, m_hi = M * rank // world, M * (rank + 1) // world
        slices = [(M * r // world, M * (r + 1) // world) for r in range(world)]
        gramS, _ = gather_sum(Sl.astype(np.float64).dot(Sl.T))
        gramA, _ = gather_sum(A[m_lo:m_hi].astype(np.float64).T.dot(A[m_lo:m_hi]))
        for it in range(iters):
            R = A.dot(Sl) - Yl
            GA_part = np.ascontiguousarray(R.dot(Sl.T))
            GS = A.T.dot(R)
            step_A = np.float32(1 / np.linalg.eigvalsh(gramS).max())
            step_S = np.float32(1 / np.linalg.eigvalsh(gramA).max())
            _, parts = gather_sum(GA_part)                                    # every rank can read every partial ...
            g_rows = np.zeros((m_hi - m_lo, K), np.float32)
            for p in parts:                                                   # ... but sums only ITS rows, rank order
                g_rows = g_rows + p.numpy()[m_lo:m_hi]
            rows_new = orc.prox_plus(A[m_lo:m_hi] - step_A * g_rows, step_A)
            pad = np.zeros((max(b - a for a, b in slices), K), np.float32)     # all-gather of the new rows
            pad[:m_hi - m_lo] = rows_new
            recv = [torch.empty_like(torch.from_numpy(pad)) for _ in range(world)]
            dist.all_gather(recv, torch.from_numpy(pad))
            A_new = np.concatenate([recv[r].numpy()[:b - a] for r, (a, b) in enumerate(slices)], axis=0)
            S_new = orc.prox_unity_plus(Sl - step_S * GS, step_S)
            gramA, _ = gather_sum(rows_new.astype(np.float64).T.dot(rows_new))
            packed = np.concatenate([S_new.astype(np.float64).dot(S_new.T).ravel(),
                                     [((S_new - Sl) ** 2).sum(), (S_new ** 2).sum(), (Sl ** 2).sum()]])
            packed, _ = gather_sum(packed)
            gramS, nS = packed[:K * K].reshape(K, K), packed[K * K:]
            A, Sl = A_new, S_new
        if rank == 0:
            np.savez(out, A=A, nS=nS)
        np.save(out + ".S%d.npy" % rank, Sl)
        np.save(out + ".A%d.npy" % rank, A)
        dist.barrier()
        dist.destroy_process_group()
        return
    gramS = allreduce(Sl.astype(np.float64).dot(Sl.T))       # before the first iteration: S S^T of the start point
    for it in range(iters):
        R = A.dot(Sl) - Yl                                   # local stripe of the residual
        GA = allreduce(R.dot(Sl.T)).astype(np.float32)       # message 1: G_A partials
        GS = A.T.dot(R)                                      # local
        step_A = np.float32(1 / np.linalg.eigvalsh(gramS).max())
        step_S = np.float32(1 / orc.lipschitz(A))
        A_new = orc.prox_plus(A - step_A * GA, step_A)
        S_new = orc.prox_unity_plus(Sl - step_S * GS, step_S)   # column sums are local to a stripe
        packed = np.concatenate([S_new.astype(np.float64).dot(S_new.T).ravel(),
                                 [((S_new - Sl) ** 2).sum(), (S_new ** 2).sum(), (Sl ** 2).sum()]])
        packed = allreduce(packed)                           # message 2: [S S^T of the new S | S-block norms]
        gramS, nS = packed[:K * K].reshape(K, K), packed[K * K:]
        A, Sl = A_new, S_new
    if rank == 0:
        np.savez(out, A=A, nS=nS)
    np.save(out + ".S%d.npy" % rank, Sl)
    np.save(out + ".A%d.npy" % rank, A)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["allreduce", "gather", "scatter"])
def test_column_sharded_pgm_matches_unsharded(tmp_path, mode):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    from oracle import proxmin_oracle as orc
    from proxmin_b200 import workloads

    world, port = 2, 29611 + os.getpid() % 200 + {"allreduce": 0, "gather": 300, "scatter": 600}[mode]
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, port, out, mode), nprocs=world, join=True)
    replicas = [np.load(out + ".A%d.npy" % r) for r in range(world)]
    if mode in ("gather", "scatter"):   # rank-ordered sums: bit-identical replicas, the property the peer exchange guarantees
        assert all(np.array_equal(replicas[0], a) for a in replicas[1:])
    got = np.load(out)
    S = np.concatenate([np.load(out + ".S%d.npy" % r) for r in range(world)], axis=1)

    M, N, K = 96, 400, 6
    Y, A, S0 = workloads.cfg2(M, N, K, seed=3)
    orc.nmf(Y, A, S0, prox_A=orc.prox_plus, prox_S=orc.prox_unity_plus, algorithm="pgm", max_iter=25, e_rel=0)
    assert np.linalg.norm(got["A"] - A) / np.linalg.norm(A) < 2e-5
    assert np.linalg.norm(S - S0) / np.linalg.norm(S0) < 2e-5
    assert np.allclose(S.sum(axis=0), 1, atol=1e-5)


# ------------------------------------------------------------------------------------------------------------------
# adaprox / AMSGrad column-sharded (BASELINE config 3): the exchange list of nmf_solver.cu's adaprox loop
# ------------------------------------------------------------------------------------------------------------------
def _adaprox_worker(rank, world, port, out):
    """Per iteration the device loop exchanges: sum(G_A partials); the row sums of S (step_adaprox needs the mean over ALL
    columns, nmf.py:91-93); max(Psi) of the S block (algorithms.py:384, a MAX); per sub-iteration of the S block the two
    norms of the stopping rule (:389).  The A block is replicated: its moments, Psi and sub-iterations are computed
    redundantly and must agree bit for bit."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import proxmin_oracle as orc
    from proxmin_b200 import workloads

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(x):
        x = np.ascontiguousarray(x)
        parts = [torch.empty_like(torch.from_numpy(x)) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(x))
        return [p.numpy() for p in parts]

    def rank_sum(x):   # k_peer_sum / k_small_allreduce: rank-ordered sum in the working precision
        acc = np.zeros_like(np.ascontiguousarray(x))
        for p in gather(x):
            acc = acc + p
        return acc

    def rank_max(x):
        return np.max(np.stack(gather(np.atleast_1d(x))), axis=0)[0]

    M, N, K = 64, 300, 5
    Y, A, S = workloads.cfg2(M, N, K, seed=13)
    lo, hi = workloads.shard_columns(N, world, rank, align=128)
    Yl, Sl, A = Y[:, lo:hi].copy(), S[:, lo:hi].copy(), A.copy()
    iters, b2, eps, p, e_rel = 12, 0.999, 1e-8, 0.25, 1e-3
    b1 = np.array((0.9,) * iters)
    MA, VA = np.zeros_like(A), np.zeros_like(A)
    MS, VS = np.zeros_like(Sl), np.zeros_like(Sl)
    sub = [0, 0]
    for it in range(iters):
        R = A.dot(Sl) - Yl
        GA = rank_sum(R.dot(Sl.T))                                        # exchange 1: G_A partials
        GS = A.T.dot(R)
        alphaA = np.mean(A, axis=0) / 10                                   # nmf.py:91-93
        rows = rank_sum(Sl.astype(np.float64).sum(axis=1))                 # exchange 2: row sums of S over all columns
        alphaS = (rows / N).astype(np.float32)[:, None] / 10
        for j, (X, G, Mj, Vj, al) in enumerate(((A, GA, MA, VA, alphaA), (Sl, GS, MS, VS, alphaS))):
            Phi, Psi = orc._phi_psi("amsgrad", it, G, Mj, Vj, None, b1, b2, eps, p)
            X[:] -= al * Phi / Psi
            z = X.copy()
            psimax = np.max(Psi) if j == 0 else rank_max(np.max(Psi))      # exchange 3: max Psi of the sharded block
            gamma = al / psimax
            for tau in range(1, 1001):
                z_new = orc.prox_plus(z - gamma / al * Psi * (z - X), gamma)
                nd, nz = orc.l2sq(z_new - z), orc.l2sq(z)
                if j == 1:                                                 # exchange 4: the sub-iteration norms
                    nd, nz = rank_sum(np.array([nd, nz], dtype=np.float64))
                z = z_new
                if np.float32(nd) <= np.float32(e_rel ** 2) * np.float32(nz):
                    break
            sub[j] += tau
            X[:] = z
    np.save(out + ".S%d.npy" % rank, Sl)
    np.save(out + ".A%d.npy" % rank, A)
    np.save(out + ".sub%d.npy" % rank, np.array(sub))
    dist.barrier()
    dist.destroy_process_group()


def test_column_sharded_adaprox_matches_unsharded(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    from oracle import proxmin_oracle as orc
    from proxmin_b200 import workloads

    world, port = 2, 30311 + os.getpid() % 200
    out = str(tmp_path / "ada")
    mp.spawn(_adaprox_worker, args=(world, port, out), nprocs=world, join=True)
    replicas = [np.load(out + ".A%d.npy" % r) for r in range(world)]
    assert all(np.array_equal(replicas[0], a) for a in replicas[1:])       # rank-ordered sums: bit-identical replicas
    subs = [np.load(out + ".sub%d.npy" % r) for r in range(world)]
    assert all(np.array_equal(subs[0], s_) for s_ in subs[1:])             # same stopping decisions on every rank
    S = np.concatenate([np.load(out + ".S%d.npy" % r) for r in range(world)], axis=1)
    M, N, K = 64, 300, 5
    Y, A, S0 = workloads.cfg2(M, N, K, seed=13)
    res = orc.nmf(Y, A, S0, algorithm="adaprox", scheme="amsgrad", max_iter=12, check_convergence=False)
    assert list(res[5]) == list(subs[0])                                   # sub-iteration counts of the unsharded run
    assert np.linalg.norm(replicas[0] - A) / np.linalg.norm(A) < 2e-5
    assert np.linalg.norm(S - S0) / np.linalg.norm(S0) < 2e-5
