"""Model check of the peer-memory exchange protocol of the sharded solvers (proxmin_b200/csrc/comm.cu).

The CUDA kernels (k_peer_signal / k_peer_sum) cannot run without GPUs; what CAN be checked on a CPU is the protocol
they implement: buffer parity = epoch & 1, a rank signals epoch e only after its partial is complete, the consumer
waits for every rank's flag >= e, sums the partials in rank order and clears ITS OWN other-parity buffer, into which
its next partial is accumulated.  The model runs R ranks as coroutines that may be pre-empted between any two memory
operations and drives them with random schedules (including one rank racing far ahead of a straggler); every rank
must obtain exactly sum_r partial(r, it) in every iteration.  Two deliberately broken variants (clearing the buffer
before the wait, one shared buffer instead of a parity pair) must be caught by the same check.
"""
import random

import numpy as np
import pytest


class Rank:
    def __init__(self, r, world, n):
        self.r, self.world, self.n = r, world, n
        self.arena = np.zeros((2, n), dtype=np.int64)     # the symmetric region: pair of partial buffers
        self.flags = np.zeros(world, dtype=np.int64)      # flags[src]: written by rank src (st.release.sys)
        self.epoch = 0                                    # local device counter, never reset
        self.results = []


def partial(r, it, n):
    rng = np.random.default_rng(1000 * it + r)
    return rng.integers(1, 1 << 20, size=n)


def rank_program(me, ranks, iters, chunks, variant="ok"):
    """One rank's kernel sequence of `iters` iterations; every `yield` is a point where another rank may run."""
    n = me.n
    bounds = np.linspace(0, n, chunks + 1).astype(int)
    for it in range(iters):
        # gradient kernel: red.add of this rank's partial into the local buffer of the NEXT epoch
        par = (me.epoch + 1) & 1 if variant != "single_buffer" else 0
        p = partial(me.r, it, n)
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            me.arena[par, lo:hi] += p[lo:hi]
            yield
        # k_peer_signal: bump the epoch, publish it to every peer
        me.epoch += 1
        e = me.epoch
        for peer in ranks:
            peer.flags[me.r] = e
            yield
        if variant == "clear_before_wait":   # BROKEN on purpose: clears a buffer its readers may still need
            me.arena[(e & 1) ^ 1, :] = 0
        # k_peer_sum: wait until every rank reached this epoch ...
        while any(me.flags[q] < e for q in range(me.world)):
            yield
        par = e & 1 if variant != "single_buffer" else 0
        out = np.zeros(n, dtype=np.int64)
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            for q in range(me.world):            # ... sum in rank order (bit-identical on every rank) ...
                out[lo:hi] += ranks[q].arena[par, lo:hi]
                yield
            if variant == "ok":                  # ... and clear the local buffer of the next epoch
                me.arena[par ^ 1, lo:hi] = 0
            elif variant == "single_buffer":     # BROKEN on purpose: peers may not have read it yet
                me.arena[0, lo:hi] = 0
            yield
        me.results.append(out)
        # (update kernels, finalize: nothing that touches the exchange state)
        yield


def run_schedule(world, iters, chunks, n, seed, variant="ok", straggler=None):
    rnd = random.Random(seed)
    ranks = [Rank(r, world, n) for r in range(world)]
    procs = {r: rank_program(ranks[r], ranks, iters, chunks, variant) for r in range(world)}
    steps = 0
    while procs:
        live = list(procs)
        if straggler is not None and straggler in procs and len(live) > 1 and rnd.random() < 0.97:
            live.remove(straggler)               # everybody else runs as far as the protocol lets them
        r = rnd.choice(live)
        try:
            next(procs[r])
        except StopIteration:
            del procs[r]
        steps += 1
        assert steps < 5_000_000, "deadlock in the exchange protocol model"
    return ranks


def check(ranks, iters, n):
    world = len(ranks)
    for it in range(iters):
        want = sum(partial(q, it, n) for q in range(world))
        for rk in ranks:
            if not np.array_equal(rk.results[it], want):
                return False
    return True


@pytest.mark.parametrize("world", [2, 3, 8])
def test_exchange_protocol_random_schedules(world):
    n, iters, chunks = 24, 6, 3
    for seed in range(40):
        ranks = run_schedule(world, iters, chunks, n, seed)
        assert check(ranks, iters, n), "wrong sum with seed %d" % seed
        for rk in ranks:                         # replicas agree bit for bit (same order of addition)
            assert all(np.array_equal(a, b) for a, b in zip(rk.results, ranks[0].results))


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_protocol_with_straggler(world):
    n, iters, chunks = 16, 8, 4
    for seed in range(20):
        for slow in range(world):
            ranks = run_schedule(world, iters, chunks, n, seed, straggler=slow)
            assert check(ranks, iters, n), "wrong sum with seed %d, straggler %d" % (seed, slow)


@pytest.mark.parametrize("variant", ["clear_before_wait", "single_buffer"])
def test_model_detects_broken_protocols(variant):
    """The checker is not vacuous: protocols without the wait-before-clear rule / the parity pair lose data."""
    n, iters, chunks = 16, 6, 4
    bad = 0
    for seed in range(60):
        ranks = run_schedule(3, iters, chunks, n, seed, variant=variant, straggler=seed % 3)
        bad += not check(ranks, iters, n)
    assert bad > 0
