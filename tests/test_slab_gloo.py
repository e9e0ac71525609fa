"""world_size-2 (and 4) CPU tests of the N>1 host logic over the gloo backend: particle partition,
schedule agreement without a broadcast, halo routing of the four colour passes, and the
cross-rank sum.  The device data path (NCCL inside libmc2d.so) is covered on real GPUs by
tools/check_multi_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from montecarlo_simulation_b200 import api
    import slab_routing_model as slab

    Lx, Ly, rcut = 321.0, 200.0, 2.5
    xy = np.random.default_rng(5).uniform(0, 1, (20000, 2)) * [Lx, Ly]
    mine = slab.partition(xy, Lx, Ly, rcut, world, rank)
    # 1. partition: every particle owned exactly once; owners agree with the C ABI's slab columns
    counts = [None] * world
    dist.all_gather_object(counts, len(mine))
    nx, ny, dx, dy = api.sweep_grid(Lx, Ly, rcut, world)
    col0, ncol = api.slab_columns(nx, world, rank)
    inside = (mine[:, 0] >= col0 * dx * (1 - 1e-15)) & (mine[:, 0] < (col0 + ncol) * dx * (1 + 1e-15))
    # 2. schedule: identical on every rank without any communication
    sched = [api.sweep_schedule(77, s, dx, dy) for s in range(50)]
    all_sched = [None] * world
    dist.all_gather_object(all_sched, sched)
    # 3. halo protocol: columns = [left ghost, owned..., right ghost]; a pass "modifies" its active
    #    boundary column (stamp = pass number); afterwards every ghost must equal the neighbour's column
    cols = torch.zeros(ncol + 2, 8, dtype=torch.float64)
    cols[1:-1] = torch.arange(col0, col0 + ncol, dtype=torch.float64)[:, None]
    left_ghost_src, right_ghost_src = (rank - 1) % world, (rank + 1) % world
    ok_halo = True
    for p, colour in enumerate(sched[0][0]):
        first = (colour % 2 == 0)
        (cols[1] if first else cols[-2]).add_(1000.0 * (p + 1))  # the pass touched this column
        lg, rg = slab.exchange_columns(dist, colour, cols[1].clone(), cols[-2].clone())
        if lg is not None:
            cols[0] = lg
        if rg is not None:
            cols[-1] = rg
    boundary = [None] * world
    dist.all_gather_object(boundary, (cols[1].tolist(), cols[-2].tolist()))
    ok_halo = (cols[0].tolist() == boundary[left_ghost_src][1]) and (cols[-1].tolist() == boundary[right_ghost_src][0])
    # 4. the cross-rank sum
    U, W, N = slab.reduce_observables(dist, float(rank + 1), 2.0 * (rank + 1), len(mine))
    q.put((rank, sum(counts), bool(inside.all()), all(s == all_sched[0] for s in all_sched), ok_halo, U, W, N))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_host_logic_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tri = world * (world + 1) / 2
    for rank, total, inside, same_sched, ok_halo, U, W, N in res:
        assert total == 20000 and N == 20000
        assert inside and same_sched and ok_halo
        assert U == tri and W == 2 * tri
