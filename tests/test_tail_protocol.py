"""Model check of the exchange protocol of the fused PGM tail (proxmin_b200/csrc/pgm_tail.cu, round 2).

The kernels cannot run without GPUs; the protocol they implement can be checked on a CPU.  Per iteration and rank:

  gradient kernel   reads the replicated A (rows pushed by ALL ranks), accumulates its G_A partial into the local
                    buffer of parity (par_ctr + 1) & 1;  next to it runs k_tail_final of the PREVIOUS iteration:
                    push the Gram partial into slot `rank` of every rank's inbox (parity of that iteration), signal
                    (set 1), wait for every rank, sum the slots in rank order
  k_peer_signal     "gradient done" (set 0)
  k_pgm_tail        A blocks: wait for set 0 of every rank, REDUCE-SCATTER (sum the partials of all ranks for the own
                    row slice, rank order), update the rows, ALL-GATHER (push the new rows into every rank's copy of
                    A), zero-fill the own other-parity partial buffer, signal "rows in place" (set 2);
                    last block: wait for set 2 of every rank, bump par_ctr

Ranks are coroutines that may be pre-empted between any two memory operations; random schedules (including a
straggler) must give every rank the serially computed factors and Gram sums in every iteration.  Broken variants
(no wait for "rows in place", one partial buffer instead of a parity pair) must be caught by the same check.  (A single
Gram inbox instead of the parity pair is NOT caught: the "rows in place" barrier at the end of every tail already keeps a
rank from running k_tail_final of iteration i + 1 before every rank has finished the one of iteration i; the kernels
keep the pair anyway.)
"""
import random

import numpy as np
import pytest

M, KG = 12, 5          # rows of A (elements stand for rows), length of the Gram partial


class Rank:
    def __init__(self, r, world):
        self.r, self.world = r, world
        self.GA = np.zeros((2, M), dtype=np.int64)            # G_A partial pair (arena)
        self.A = np.arange(1, M + 1, dtype=np.int64)          # replicated factor (arena copy, all rows)
        self.inbox = np.zeros((2, world, KG), dtype=np.int64)  # Gram inbox [parity][source rank]
        self.flags = np.zeros((4, world), dtype=np.int64)     # flags[set][source rank]
        self.epoch = np.zeros(4, dtype=np.int64)
        self.par_ctr = 0
        self.hist_A, self.hist_gram = [], []
        lo = M * r // world
        hi = M * (r + 1) // world
        self.rows = (lo, hi)


def grad_partial(A, r, it):
    """stand-in for R S^T: depends on EVERY row of A the rank sees, on the rank and on the iteration"""
    return (A * (r + 2) + 7 * it + r) % 1000003


def gram_partial(A_rows_new, r, it):
    return (np.resize(A_rows_new, KG) * (r + 1) + it) % 1000003


def update(a, g):
    return (a * 3 + g) % 1000003


def serial_reference(world, iters):
    A = np.arange(1, M + 1, dtype=np.int64)
    out_A, out_G = [], []
    for it in range(iters):
        g = sum(grad_partial(A, r, it) for r in range(world))
        A = update(A, g)
        out_A.append(A.copy())
        gp = 0
        for r in range(world):
            lo, hi = M * r // world, M * (r + 1) // world
            gp = gp + gram_partial(A[lo:hi], r, it)
        out_G.append(gp)
    return out_A, out_G


def grad_steps(me, it, variant):
    par = (me.par_ctr + 1) & 1 if variant != "single_partial" else 0
    snapshot = []
    for m in range(M):                       # the kernel reads A while it runs (rows in any order of time)
        snapshot.append(me.A[m])
        if m % 3 == 2:
            yield
    p = grad_partial(np.array(snapshot, dtype=np.int64), me.r, it)
    for m in range(0, M, 4):
        me.GA[par, m:m + 4] += p[m:m + 4]
        yield


def final_steps(me, ranks, it, variant):
    """k_tail_final of iteration `it` (called while the gradient kernel of it + 1 runs)"""
    pc = it
    par = (pc + 1) & 1 if variant != "single_inbox" else 0
    lo, hi = me.rows
    gp = gram_partial(me.hist_A[it][lo:hi], me.r, it)
    for peer in ranks:
        peer.inbox[par, me.r, :] = gp
        yield
    me.epoch[1] += 1
    e = me.epoch[1]
    for peer in ranks:
        peer.flags[1, me.r] = e
        yield
    while any(me.flags[1, q] < e for q in range(me.world)):
        yield
    tot = np.zeros(KG, dtype=np.int64)
    for q in range(me.world):
        tot += me.inbox[par, q]
        yield
    me.hist_gram.append(tot)


def tail_steps(me, ranks, it, variant):
    pc = me.par_ctr
    par = (pc + 1) & 1 if variant != "single_partial" else 0
    # k_peer_signal: gradient done
    me.epoch[0] += 1
    e0 = me.epoch[0]
    for peer in ranks:
        peer.flags[0, me.r] = e0
        yield
    # A blocks
    while any(me.flags[0, q] < e0 for q in range(me.world)):
        yield
    lo, hi = me.rows
    g = np.zeros(hi - lo, dtype=np.int64)
    for q in range(me.world):                # reduce-scatter in rank order
        g += ranks[q].GA[par, lo:hi]
        yield
    new = update(me.A[lo:hi].copy(), g)
    for peer in ranks:                       # all-gather: push the rows into every rank's copy
        peer.A[lo:hi] = new
        yield
    for m in range(0, M, 4):                 # zero-fill of the own other-parity buffer
        me.GA[par ^ 1 if variant != "single_partial" else 0, m:m + 4] = 0
        yield
    me.epoch[2] += 1
    e2 = me.epoch[2]
    for peer in ranks:
        peer.flags[2, me.r] = e2
        yield
    if variant != "no_rows_wait":
        while any(me.flags[2, q] < e2 for q in range(me.world)):
            yield
    me.par_ctr = pc + 1
    me.hist_A.append(me.A.copy())            # what the next gradient kernel of this rank is entitled to see
    yield


def rank_program(me, ranks, iters, variant, rnd):
    for it in range(iters):
        # graph of iteration `it`: [k_tail_final(it - 1)] runs next to [gradient kernel(it)], both before the tail
        streams = [grad_steps(me, it, variant)]
        if it > 0:
            streams.append(final_steps(me, ranks, it - 1, variant))
        while streams:
            s = rnd.choice(streams)
            try:
                next(s)
            except StopIteration:
                streams.remove(s)
            yield
        for _ in tail_steps(me, ranks, it, variant):
            yield
    for _ in final_steps(me, ranks, iters - 1, variant):   # tail_close before the host reads the control block
        yield


def run_schedule(world, iters, seed, variant="ok", straggler=None):
    rnd = random.Random(seed)
    ranks = [Rank(r, world) for r in range(world)]
    procs = {r: rank_program(ranks[r], ranks, iters, variant, rnd) for r in range(world)}
    steps = 0
    while procs:
        live = list(procs)
        if straggler is not None and straggler in procs and len(live) > 1 and rnd.random() < 0.95:
            live.remove(straggler)
        r = rnd.choice(live)
        try:
            next(procs[r])
        except StopIteration:
            del procs[r]
        steps += 1
        assert steps < 3_000_000, "deadlock in the fused-tail exchange model"
    return ranks


def check(ranks, iters):
    want_A, want_G = serial_reference(len(ranks), iters)
    for rk in ranks:
        for it in range(iters):
            if not np.array_equal(rk.hist_A[it], want_A[it]):
                return False
            if not np.array_equal(rk.hist_gram[it], want_G[it]):
                return False
    return True


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_tail_protocol_random_schedules(world):
    iters = 5
    for seed in range(25):
        ranks = run_schedule(world, iters, seed)
        assert check(ranks, iters), "wrong factors / Gram sums with seed %d" % seed


@pytest.mark.parametrize("world", [2, 4])
def test_tail_protocol_with_straggler(world):
    iters = 6
    for seed in range(12):
        for slow in range(world):
            ranks = run_schedule(world, iters, seed, straggler=slow)
            assert check(ranks, iters), "wrong result with seed %d, straggler %d" % (seed, slow)


@pytest.mark.parametrize("variant", ["no_rows_wait", "single_partial"])
def test_model_detects_broken_tail_protocols(variant):
    """The checker is not vacuous: without the wait for "rows in place" the next gradient reads stale rows, a single
    partial buffer is zero-filled under its readers."""
    iters = 6
    bad = 0
    for seed in range(60):
        try:
            ranks = run_schedule(3, iters, seed, variant=variant, straggler=seed % 3)
            bad += not check(ranks, iters)
        except (AssertionError, IndexError):
            bad += 1
    assert bad > 0
