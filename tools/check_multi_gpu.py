#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multi_gpu.py

The Philox streams are keyed by the GLOBAL cell id and the colour order / grid shift are pure
functions of (seed, sweep), so a slab-decomposed run must reproduce the single-GPU trajectory
BIT FOR BIT (positions, N, counters) and its energy to summation-order accuracy.  Rank 0 runs the
single-GPU reference on its own device and compares."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlo_simulation_b200 as m
from montecarlo_simulation_b200 import api


def lattice(n, L, seed):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2).astype(float)
    return (g + 0.5) * (L / n) + rng.uniform(-0.2, 0.2, g.shape)


def sort_rows(a):
    return a[np.lexsort((a[:, 1], a[:, 0]))]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_side = 227
    L = 321.0  # int(L/2.5) = 128 cell columns: the same grid for 1, 2, 4 and 8 slabs; rho = 0.5
    nx1 = api.sweep_grid(L, L, 2.5, 1)[0]
    nxw = api.sweep_grid(L, L, 2.5, world)[0]
    assert nx1 == nxw, "pick a box whose cell count is a multiple of 2*world (%d vs %d)" % (nx1, nxw)
    xy = lattice(n_side, L, 7)
    ok = True
    for label, kw, plan in (("NVT", dict(activity=0.0), dict(gc_per_cell=0)),
                            ("GCMC", dict(activity=0.497), dict(gc_per_cell=1))):
        ctx = m.Context(L, L, 2.5, T=1.0, delta=0.1, seed=11, device=local, rank=rank, nranks=world, **kw)
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0])
        ctx.upload(xy)  # every rank passes the full set; the library keeps its slab
        # Re-uploads into live storage (the end-to-end pattern) with the odd ranks held back after the collective:
        # the neighbours' ghost pushes land before the late rank converts its own cells, which must leave the
        # ghost columns alone.  NVT pass: direct binning; GCMC pass: the sorted path.
        os.environ["MC2D_TEST_UPLOAD_LAG_US"] = "3000"
        if label == "GCMC":
            os.environ["MC2D_UPLOAD_SORT"] = "1"
        for _ in range(3):
            ctx.upload(xy)
        os.environ.pop("MC2D_TEST_UPLOAD_LAG_US")
        os.environ.pop("MC2D_UPLOAD_SORT", None)
        U0, W0, N0 = ctx.total_energy_virial(allreduce=True)
        st = ctx.run_sweeps(25, **plan)
        st_all = ctx.allreduce_stats(st)
        U1, W1, N1 = ctx.total_energy_virial(allreduce=True)
        mine = ctx.download()
        sizes = [None] * world
        dist.all_gather_object(sizes, len(mine))
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        ctx.close()
        if rank == 0:
            allp = np.concatenate(parts)
            with m.Context(L, L, 2.5, T=1.0, delta=0.1, seed=11, device=local, **kw) as one:
                one.upload(xy)
                u0, w0, n0 = one.total_energy_virial()
                s1 = one.run_sweeps(25, **plan)
                u1, w1, n1 = one.total_energy_virial()
                ref = one.download()
            same_pos = allp.shape == ref.shape and bool((sort_rows(allp) == sort_rows(ref)).all())
            same_cnt = st_all["attempted"] == s1["attempted"] and st_all["accepted"] == s1["accepted"]
            rel = lambda a, b: abs(a - b) / max(1.0, abs(b))
            good = (same_pos and same_cnt and N0 == n0 and N1 == n1 and rel(U0, u0) < 1e-12 and rel(U1, u1) < 1e-12
                    and rel(W1, w1) < 1e-12 and rel(st_all["energy"], s1["energy"]) < 1e-9)
            print("%s world=%d: positions bitwise equal=%s counters equal=%s N %d/%d U %.12g/%.12g running %.12g/%.12g -> %s"
                  % (label, world, same_pos, same_cnt, N1, n1, U1, u1, st_all["energy"], s1["energy"],
                     "OK" if good else "MISMATCH"), flush=True)
            ok = ok and good
    # a position outside the box on ONE rank must fail the upload on EVERY rank (no rank left waiting in a collective)
    ctx = m.Context(L, L, 2.5, T=1.0, delta=0.1, seed=11, device=local, rank=rank, nranks=world)
    ids = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0])
    for attempt in range(2):  # first upload: sorted path; second: direct binning
        bad = xy.copy()
        if rank == world - 1:
            bad[5, 0] = L + 1.0
        try:
            ctx.upload(bad if attempt == 1 else xy)
            raised = False
        except m.MC2DError as e:
            raised = e.status == 6
        want = attempt == 1
        if raised != want:
            print("rank %d: upload attempt %d raised=%s, expected %s" % (rank, attempt, raised, want), flush=True)
            ok = False
    ctx.close()
    bad_any = torch.tensor([0 if (ok or rank == 0) else 1], device="cuda")
    dist.all_reduce(bad_any)
    if rank == 0:
        ok = ok and int(bad_any.item()) == 0
        print("out-of-box on one rank fails everywhere -> %s" % ("OK" if int(bad_any.item()) == 0 else "MISMATCH"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
