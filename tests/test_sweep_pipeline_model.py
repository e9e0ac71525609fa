"""Item-level model of ONE cooperative sweep launch in slab mode (k_sweep_p with `pipeline` and `halo_fused`,
montecarlo_simulation_b200/csrc/sweep_kernel_p.cuh): CPU only, no GPU and no library.  The existing
tests/test_halo_protocol_model.py models the ranks and their sync points; this one models what happens INSIDE the kernel:

  * the items of the four colour passes form one ticket list (pass-major, then column pair, then rank in the pair); for
    a pass over odd columns the slab-edge pair (the last one) is rotated to the front;
  * W persistent warps per rank draw tickets in order; an item of pass p > 0 in pair q waits until ALL items of pass
    p - 1 in the pairs q-1, q, q+1 are finished (per-pair completion counters; the pairs do not wrap in a slab);
  * a slab-edge item also waits for the neighbour's flag word >= pass index ("the neighbour's edge items of all earlier
    passes are finished"), reads the ghost column, writes its cells and pushes them into the neighbour's ghost column;
  * whoever completes the edge items of a pass takes a lock and moves a frontier over every pass complete by then,
    storing frontier into both neighbours' flag words — BEFORE the warp waits for its next item.

Checked under thousands of random interleavings, for tiny grids and one or two warps as well:
  (1) no deadlock;
  (2) an item never reads cells that an unfinished item of the previous pass or a running item of a later pass writes;
  (3) a ghost column is read only when it holds exactly the neighbour's pushes of the earlier passes, and is never
      pushed into while it is read.
Two deliberately broken variants must be caught: signals out of pass order, and signalling after the blocking wait."""
import random

import pytest


class Violation(AssertionError):
    pass


class Rank:
    def __init__(self, r, n, P, I, W):
        self.r, self.n, self.P, self.I, self.W = r, n, P, I, W
        self.flag = [0, 0]  # [from left, from right]; 0 = the re-bin's signal has arrived
        self.next_ticket = 0
        self.done = [[0] * P for _ in range(4)]
        self.edge_done = [0] * 4
        self.frontier = 0
        self.lock = False
        self.reading = {}  # warp -> (pass, pair)
        self.writing = {}  # warp -> (pass, pair)
        self.ghost_readers = [set(), set()]  # side -> passes being read
        self.pushes_done = [[0] * 4, [0] * 4]  # side -> pushes of my pass ps into the neighbour on that side
        self.pushing = [set(), set()]


def ticket_list(P, I, colours, rotate=True):
    items = []
    for ps in range(4):
        px = colours[ps] & 1
        for q in range(P):
            p = q
            if rotate and px == 1:
                p = P - 1 if q == 0 else q - 1
            for r in range(I):
                items.append((ps, p, r))
    return items


def warp(rk, w, ranks, colours, items, variant):
    n, P, I = rk.n, rk.P, rk.I
    nb = [ranks[(rk.r - 1) % n], ranks[(rk.r + 1) % n]]

    def deps_ok(ps, p):
        if ps > 0:
            for q in (max(p - 1, 0), p, min(p + 1, P - 1)):
                if rk.done[ps - 1][q] < I:
                    return False
        px = colours[ps] & 1
        if p == (0 if px == 0 else P - 1) and rk.flag[px] < ps:
            return False
        return True

    def signal_frontier(ps):
        rk.edge_done[ps] += 1
        if rk.edge_done[ps] < I:
            return
        while rk.lock:
            yield
        rk.lock = True
        if variant == "unordered":  # announces its own pass, whatever the earlier ones are doing
            nb[0].flag[1] = max(nb[0].flag[1], ps + 1)
            nb[1].flag[0] = max(nb[1].flag[0], ps + 1)
        else:
            while rk.frontier < 4 and rk.edge_done[rk.frontier] == I:
                rk.frontier += 1
                yield  # (a fence between the counter read and the stores)
                nb[0].flag[1] = rk.frontier  # plain stores stay monotonic under the lock
                nb[1].flag[0] = rk.frontier
        rk.lock = False

    pending_signal = None
    while True:
        t = rk.next_ticket
        rk.next_ticket += 1
        if t >= len(items):
            if pending_signal is not None:
                yield from signal_frontier(pending_signal)
            return
        ps, p, r = items[t]
        if variant != "late_signal" and pending_signal is not None:
            yield from signal_frontier(pending_signal)
            pending_signal = None
        while not deps_ok(ps, p):
            yield
        if pending_signal is not None:  # broken variant: the signal of the previous item goes out after the wait
            yield from signal_frontier(pending_signal)
            pending_signal = None
        px = colours[ps] & 1
        edge = p == (0 if px == 0 else P - 1)
        side = px  # even columns: first pair, left ghost, pushed to the left neighbour; odd: last pair, right
        # ---- read the neighbourhood (and the ghost column) ----
        rk.reading[w] = (ps, p)
        if edge:
            rk.ghost_readers[side].add(ps)
        for _ in range(2):  # the checks hold at the start and at the end of the read
            for w2, (ps2, p2) in rk.writing.items():
                if abs(p2 - p) <= 1 and ps2 != ps:
                    raise Violation("rank %d: item (%d,%d) reads while item (%d,%d) writes next to it" % (rk.r, ps, p, ps2, p2))
            if ps > 0:
                for q in (max(p - 1, 0), p, min(p + 1, P - 1)):
                    if rk.done[ps - 1][q] < I:
                        raise Violation("rank %d: item (%d,%d) started before pair %d of the previous pass was complete" % (rk.r, ps, p, q))
            if edge:
                src, sside = nb[side], 1 - side  # the neighbour pushes towards me on ITS other side
                for ps2 in range(4):
                    if (colours[ps2] & 1) != sside:
                        continue
                    have = src.pushes_done[sside][ps2]
                    want = src.I if ps2 < ps else 0
                    if have != want or ps2 in src.pushing[sside]:
                        raise Violation("rank %d pass %d: ghost side %d holds %d pushes of the neighbour's pass %d, wants %d" %
                                        (rk.r, ps, side, have, ps2, want))
            yield
        del rk.reading[w]
        if edge:
            rk.ghost_readers[side].discard(ps)
        # ---- write the cells back (and push them into the neighbour's ghost column) ----
        rk.writing[w] = (ps, p)
        for w2, (ps2, p2) in rk.reading.items():
            if abs(p2 - p) <= 1 and ps2 != ps:
                raise Violation("rank %d: item (%d,%d) writes while item (%d,%d) reads next to it" % (rk.r, ps, p, ps2, p2))
        if edge:
            target, tside = nb[side], 1 - side
            if target.ghost_readers[tside]:
                raise Violation("rank %d pass %d pushes into ghost %d of rank %d while it is being read" % (rk.r, ps, tside, target.r))
            rk.pushing[side].add(ps)
        yield
        if edge:
            rk.pushes_done[side][ps] += 1
            if rk.pushes_done[side][ps] == I:
                rk.pushing[side].discard(ps)
        del rk.writing[w]
        rk.done[ps][p] += 1
        if edge:
            pending_signal = ps
        yield


def simulate(n, P, I, W, seed, variant="ok", rotate=True, idle_limit=20000):
    rng = random.Random(seed)
    colours = rng.sample(range(4), 4)
    ranks = [Rank(r, n, P, I, W) for r in range(n)]
    items = ticket_list(P, I, colours, rotate)
    threads = [(rk, warp(rk, w, ranks, colours, items, variant)) for rk in ranks for w in range(W)]
    idle = 0
    while threads:
        # skewed scheduling: one rank, or one warp, often runs far ahead of the others
        k = rng.randrange(len(threads)) if rng.random() < 0.6 else rng.randrange(min(len(threads), 1 + len(threads) // 3))
        rk, g = threads[k]
        before = [(x.next_ticket, tuple(x.flag), tuple(x.edge_done), tuple(map(tuple, x.done)), x.frontier, len(x.reading),
                   len(x.writing)) for x in ranks]
        try:
            next(g)
        except StopIteration:
            threads.pop(k)
        after = [(x.next_ticket, tuple(x.flag), tuple(x.edge_done), tuple(map(tuple, x.done)), x.frontier, len(x.reading),
                  len(x.writing)) for x in ranks]
        idle = idle + 1 if before == after else 0
        if idle > idle_limit:
            raise Violation("deadlock: no warp makes progress")
    for rk in ranks:
        assert all(c == I for row in rk.done for c in row)
        assert variant == "unordered" or rk.frontier == 4
    return True


@pytest.mark.parametrize("n,P,I,W", [(2, 1, 1, 1), (2, 2, 1, 1), (2, 3, 2, 2), (3, 2, 3, 1), (3, 5, 2, 3), (4, 3, 1, 7), (2, 4, 3, 5)])
def test_no_deadlock_and_no_race_under_random_schedules(n, P, I, W):
    for seed in range(150):
        assert simulate(n, P, I, W, seed=1000 * n + 100 * P + seed)
        assert simulate(n, P, I, W, seed=1000 * n + 100 * P + seed, rotate=False)  # the order before the rotation holds too


def test_out_of_order_signals_are_caught():
    caught = 0
    for seed in range(300):
        try:
            simulate(3, 3, 2, 4, seed=seed, variant="unordered")
        except Violation:
            caught += 1
    assert caught > 10


def test_signalling_after_the_blocking_wait_deadlocks():
    """one warp per rank and one column pair (every item is a slab-edge item): the next item depends, through the
    neighbour, on the signal the warp still owes"""
    caught = 0
    for seed in range(60):
        try:
            simulate(2, 1, 1, 1, seed=seed, variant="late_signal", idle_limit=3000)
        except Violation as e:
            caught += "deadlock" in str(e)
    assert caught > 10
