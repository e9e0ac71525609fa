"""Randomised model check of the slab halo protocol (K5; montecarlo_simulation_b200/csrc/group_kernels.cuh,
sweep_kernel_p.cuh `halo_fused`).  CPU only, no GPU and no library: the protocol is restated as a tiny
event simulation and run under thousands of random interleavings of the ranks.

The protocol: every rank counts the same sync points (index 5k for the re-bin of sweep k, 5k+1+p for colour pass
p).  After a sync point a rank stores the boundary column(s) that changed into the neighbour's ghost column and
then stores the sync index + 1 into BOTH neighbours' flag words.  Before reading a ghost column it spins until the
flag word from that side has reached the index of the sync point it is about to execute:
  * a pass over even columns reads the LEFT ghost and pushes its FIRST column to the LEFT neighbour,
  * a pass over odd columns reads the RIGHT ghost and pushes its LAST column to the RIGHT neighbour,
  * a re-bin reads both ghosts of the old buffer and pushes both columns into the other ping-pong buffer.
No handshake goes the other way.  The properties checked here are the ones that argument rests on:
  (1) a ghost column read at sync point t contains exactly the neighbour's modifications with index < t
      (not stale, not from the future),
  (2) a ghost column is never written while its owner reads it.
A deliberately broken variant (waiting on the wrong side) must be caught."""
import random

import pytest


class Violation(AssertionError):
    pass


class Rank:
    def __init__(self, r, n):
        self.r, self.n = r, n
        self.flag = [0, 0]  # [from left, from right]
        # ghost[buf][side] = version of the neighbour's column it holds; side 0 = left ghost, 1 = right ghost
        self.ghost = [[0, 0], [0, 0]]
        self.col = [0, 0]  # modification count of my first / last column
        self.reading = set()  # (buf, side) being read right now


def expected_version(orders, t, column):
    """number of modifications with index < t of a rank's first (0) / last (1) column"""
    v = 0
    for idx in range(t):
        k, ph = divmod(idx, 5)
        if ph == 0:
            v += 1  # the re-bin rewrites both columns
        else:
            even = (orders[k][ph - 1] & 1) == 0
            v += 1 if (even and column == 0) or (not even and column == 1) else 0
    return v


def program(rk, ranks, orders, nsweeps, broken):
    n = rk.n
    left, right = ranks[(rk.r - 1) % n], ranks[(rk.r + 1) % n]
    nb = [left, right]

    def wait(side, seq):
        while rk.flag[side] < seq:
            yield

    def read(buf, side, t):
        rk.reading.add((buf, side))
        yield  # the read takes time
        # my left ghost mirrors the left neighbour's LAST column, my right ghost the right neighbour's FIRST one
        want = expected_version(orders, t, 1 if side == 0 else 0)
        if rk.ghost[buf][side] != want:
            raise Violation("rank %d sync %d: ghost side %d holds version %d, wants %d" % (rk.r, t, side, rk.ghost[buf][side], want))
        rk.reading.discard((buf, side))

    def push(buf, side):
        """my first column -> left neighbour's right ghost (side 0); my last column -> right neighbour's left ghost"""
        target, tside = nb[side], 1 - side
        if (buf, tside) in target.reading:
            raise Violation("rank %d writes ghost %d of rank %d while it is being read" % (rk.r, tside, target.r))
        target.ghost[buf][tside] = rk.col[side]

    def signal(seq):
        left.flag[1] = max(left.flag[1], seq)  # I am the left neighbour's RIGHT neighbour
        right.flag[0] = max(right.flag[0], seq)

    for k in range(nsweeps):
        s, old, new = 5 * k, k % 2, (k + 1) % 2
        # ---- re-bin ----
        yield from wait(0, s)
        yield from wait(1, s)
        yield from read(old, 0, s)
        yield from read(old, 1, s)
        rk.col[0] += 1
        rk.col[1] += 1
        yield
        push(new, 0)
        yield
        push(new, 1)
        signal(s + 1)
        # ---- four colour passes ----
        for p in range(4):
            t = s + 1 + p
            even = (orders[k][p] & 1) == 0
            side = 0 if even else 1
            yield from wait((1 - side) if broken else side, t)
            yield from read(new, side, t)
            rk.col[side] += 1
            yield
            push(new, side)
            signal(t + 1)
            yield
    # what follows a batch (energy, download) reads both ghosts
    t = 5 * nsweeps
    yield from wait(0, t)
    yield from wait(1, t)
    yield from read(nsweeps % 2, 0, t)
    yield from read(nsweeps % 2, 1, t)


def simulate(n, nsweeps, seed, broken=False):
    rng = random.Random(seed)
    orders = [rng.sample(range(4), 4) for _ in range(nsweeps)]
    ranks = [Rank(r, n) for r in range(n)]
    # initial upload: every ghost holds version 0 of the neighbour's column in buffer 0, nothing signalled yet
    gens = [program(rk, ranks, orders, nsweeps, broken) for rk in ranks]
    alive = list(range(n))
    idle = 0
    while alive:
        # a skewed scheduler: one rank is sometimes much faster than the others
        i = rng.choice(alive) if rng.random() < 0.7 else alive[0]
        before = [(rk.flag[:], rk.col[:]) for rk in ranks]
        try:
            next(gens[i])
        except StopIteration:
            alive.remove(i)
        idle = idle + 1 if before == [(rk.flag[:], rk.col[:]) for rk in ranks] else 0
        if idle > 20000:
            raise Violation("deadlock: no rank makes progress")
    return True


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_protocol_holds_under_random_schedules(n):
    for seed in range(300):
        assert simulate(n, nsweeps=4, seed=1000 * n + seed)


def test_the_checker_catches_a_broken_protocol():
    caught = 0
    for seed in range(200):
        try:
            simulate(4, nsweeps=4, seed=seed, broken=True)
        except Violation:
            caught += 1
    assert caught > 100


# ---- re-upload into live storage (mc2d_upload with nranks > 1) ---------------------------------------------------------
# After the collective that agrees on the cell capacity every rank converts ITS cells and pushes its two boundary
# columns; nothing orders one rank's conversion against a neighbour's push.  The conversion must therefore leave the
# ghost columns alone (k_csr_to_slots / k_order_cells write owned cells only).  The broken variant is what the first
# version did: the conversion also reset the ghost counts.

def upload_program(rk, ranks, seq0, broken):
    left, right = ranks[(rk.r - 1) % rk.n], ranks[(rk.r + 1) % rk.n]
    yield  # ... arrives from the collective at its own time
    rk.col = [1, 1]  # the new configuration of my boundary columns
    if broken:
        yield
        rk.ghost[0] = [0, 0]  # conversion kernel writes the ghost counts too
    yield
    left.ghost[0][1] = rk.col[0]
    yield
    right.ghost[0][0] = rk.col[1]
    left.flag[1] = max(left.flag[1], seq0 + 1)
    right.flag[0] = max(right.flag[0], seq0 + 1)
    while rk.flag[0] < seq0 + 1 or rk.flag[1] < seq0 + 1:
        yield
    if rk.ghost[0] != [1, 1]:
        raise Violation("rank %d reads ghosts %r after the upload" % (rk.r, rk.ghost[0]))


def simulate_upload(n, seed, broken=False):
    rng = random.Random(seed)
    ranks = [Rank(r, n) for r in range(n)]
    gens = [upload_program(rk, ranks, 0, broken) for rk in ranks]
    alive = list(range(n))
    steps = 0
    while alive:
        i = rng.choice(alive) if rng.random() < 0.6 else alive[-1]  # the last rank is often far ahead
        try:
            next(gens[i])
        except StopIteration:
            alive.remove(i)
        steps += 1
        if steps > 100000:
            raise Violation("deadlock")
    return True


@pytest.mark.parametrize("n", [2, 3, 8])
def test_upload_leaves_ghost_columns_to_the_neighbour(n):
    for seed in range(300):
        assert simulate_upload(n, seed)
    caught = 0
    for seed in range(300):
        try:
            simulate_upload(n, seed, broken=True)
        except Violation:
            caught += 1
    assert caught > 30


# ---- slab pipeline: no grid barrier between the colour passes (k_sweep_p, pipeline + halo_fused) ------------------------
# Without the barrier the slab-edge items of a rank's passes form two chains that run concurrently: the items of the
# FIRST column pair (edge items of the even passes; they read the left ghost) and those of the LAST pair (edge items of
# the odd passes; right ghost).  Inside a chain the passes stay ordered (an item waits for the previous pass's items of
# its own pair), between the chains nothing orders them.  The flag value t + 1 must nevertheless mean "every edge item of
# my sync points <= t is finished", because the neighbour takes it as the licence to overwrite the ghost column I read
# AND as the promise that the ghost column it reads is complete.  So signals go out in pass order: whoever completes a
# pass moves a frontier over all passes that are complete by then (max-merge, two warps may signal back to back).  The
# broken variant signals t + 1 as soon as the edge item of pass t is done — the naive port of the barrier version.

def pipelined_rank(rk, ranks, orders, nsweeps, broken):
    """-> the generators (threads) of one rank: a sequencer (re-bin, then releases the two chains of the sweep) and
    the chains themselves are created per sweep by the sequencer."""
    n = rk.n
    left, right = ranks[(rk.r - 1) % n], ranks[(rk.r + 1) % n]
    nb = [left, right]
    state = {"threads": []}

    def signal(seq):
        left.flag[1] = max(left.flag[1], seq)
        right.flag[0] = max(right.flag[0], seq)

    def read(buf, side, t):
        rk.reading.add((buf, side))
        yield
        want = expected_version(orders, t, 1 if side == 0 else 0)
        if rk.ghost[buf][side] != want:
            raise Violation("rank %d sync %d: ghost side %d holds version %d, wants %d" % (rk.r, t, side, rk.ghost[buf][side], want))
        rk.reading.discard((buf, side))

    def push(buf, side):
        target, tside = nb[side], 1 - side
        if (buf, tside) in target.reading:
            raise Violation("rank %d writes ghost %d of rank %d while it is being read" % (rk.r, tside, target.r))
        target.ghost[buf][tside] = rk.col[side]

    def chain(k, side, done, frontier):
        s, new = 5 * k, (k + 1) % 2
        for p in range(4):
            if ((orders[k][p] & 1) == 0) != (side == 0):
                continue
            t = s + 1 + p
            while rk.flag[side] < t:
                yield
            yield from read(new, side, t)
            rk.col[side] += 1
            yield
            push(new, side)
            yield
            done[p] = True
            if broken:
                signal(t + 1)
            else:
                while frontier[0] < 4 and done[frontier[0]]:
                    frontier[0] += 1
                    signal(s + 1 + frontier[0])
            yield

    def sequencer():
        for k in range(nsweeps):
            s, old, new = 5 * k, k % 2, (k + 1) % 2
            for side in (0, 1):
                while rk.flag[side] < s:
                    yield
            yield from read(old, 0, s)
            yield from read(old, 1, s)
            rk.col[0] += 1
            rk.col[1] += 1
            yield
            push(new, 0)
            yield
            push(new, 1)
            signal(s + 1)
            done, frontier = [False] * 4, [0]
            chains = [chain(k, 0, done, frontier), chain(k, 1, done, frontier)]
            state["threads"] = chains[:]
            while state["threads"]:  # the kernel ends when both chains are through
                yield
        t = 5 * nsweeps
        for side in (0, 1):
            while rk.flag[side] < t:
                yield
        yield from read(nsweeps % 2, 0, t)
        yield from read(nsweeps % 2, 1, t)

    return sequencer(), state


def simulate_pipelined(n, nsweeps, seed, broken=False):
    rng = random.Random(seed)
    orders = [rng.sample(range(4), 4) for _ in range(nsweeps)]
    ranks = [Rank(r, n) for r in range(n)]
    seqs = [pipelined_rank(rk, ranks, orders, nsweeps, broken) for rk in ranks]
    alive = list(range(n))
    idle = 0
    while alive:
        i = rng.choice(alive) if rng.random() < 0.7 else alive[0]
        gen, state = seqs[i]
        before = [(rk.flag[:], rk.col[:]) for rk in ranks]
        # one step of this rank: one of its chains (a chain is often far ahead of the other) or its sequencer
        th = state["threads"]
        try:
            if th and rng.random() < 0.9:
                c = th[0] if rng.random() < 0.5 else rng.choice(th)
                try:
                    next(c)
                except StopIteration:
                    th.remove(c)
            else:
                next(gen)
        except StopIteration:
            alive.remove(i)
        idle = idle + 1 if before == [(rk.flag[:], rk.col[:]) for rk in ranks] else 0
        if idle > 40000:
            raise Violation("deadlock: no rank makes progress")
    return True


@pytest.mark.parametrize("n", [2, 3, 8])
def test_pipelined_protocol_signals_in_pass_order(n):
    for seed in range(300):
        assert simulate_pipelined(n, nsweeps=4, seed=7000 * n + seed)


def test_the_checker_catches_out_of_order_signals():
    caught = 0
    for seed in range(300):
        try:
            simulate_pipelined(4, nsweeps=4, seed=seed, broken=True)
        except Violation:
            caught += 1
    assert caught > 30
