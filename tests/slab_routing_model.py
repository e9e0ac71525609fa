"""TEST SCAFFOLDING (used by tests/test_slab_gloo.py only): the routing arithmetic of the 1-D slab decomposition
restated over torch.distributed tensors, so that it can be exercised with the gloo backend on a CPU-only host.

The data path itself (halo columns, migrants, the cross-GPU sum) runs inside libmc2d.so over NVLink peer memory /
NCCL and is tested on GPUs by tests/test_gpu_multi.py and bench.py's `parity` object; nothing here is product code.

Replaces: mpi_src/simulatin_box.cpp Decomposition (Px = nranks, Py = 1), the neighbour bookkeeping
of mpi_src/particle_exchange.cpp:96-112, and bcast_seed_ (MC_parallel.cpp:218-230; not needed: the
schedule is a pure function of (seed, sweep))."""
from __future__ import annotations

import numpy as np

from montecarlo_simulation_b200 import api


def slab_of_x(x, Lx, Ly, rcut, nranks):
    """rank that owns coordinate(s) x on the unshifted grid (what mc2d_upload keeps)."""
    nx, _, dx, _ = api.sweep_grid(Lx, Ly, rcut, nranks)
    col = np.floor(np.asarray(x, dtype=np.float64) * (1.0 / dx)).astype(np.int64)
    col = np.where(col >= nx, col - nx, col)
    col = np.where(col < 0, col + nx, col)
    return (col // (nx // nranks)).astype(np.int64)


def partition(xy, Lx, Ly, rcut, nranks, rank):
    """the particles rank `rank` owns (every particle belongs to exactly one rank)."""
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    return xy[slab_of_x(xy[:, 0], Lx, Ly, rcut, nranks) == rank]


def exchange_columns(dist, colour, first_col, last_col):
    """One halo step of a colour pass over torch.distributed tensors (CPU/gloo or CUDA/nccl):
    the pass modified the first owned column when the colour's x-parity is 0, the last owned
    column otherwise (mc2d_halo_route); that column goes to the neighbour that reads it.
    Returns (new_left_ghost, new_right_ghost), one of them None."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    send_left, to, frm = api.halo_route(colour, rank, world)
    out = (first_col if send_left else last_col).contiguous()
    inc = torch.empty_like(out)
    if world == 1:
        inc.copy_(out)
    else:
        reqs = [dist.isend(out, dst=to), dist.irecv(inc, src=frm)]
        for r in reqs:
            r.wait()
    # what came from the right neighbour is my right ghost (its first column), and vice versa
    return (None, inc) if send_left else (inc, None)


def reduce_observables(dist, U, W, N):
    """sum of {U, W, N} over ranks — the single cross-GPU sum (replaces the MPI_Allreduce sites of
    thermodynamic_calculator_parallel.cpp:41-131) for callers that hold per-rank values."""
    import torch

    t = torch.tensor([U, W, float(N)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return float(t[0]), float(t[1]), int(round(float(t[2])))
