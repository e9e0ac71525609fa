#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multi_gpu.py

The Philox streams are keyed by the GLOBAL cell id and the colour order / grid shift are pure
functions of (seed, sweep), so a slab-decomposed run must reproduce the single-GPU trajectory
BIT FOR BIT (positions, N, counters) and its energy to summation-order accuracy.  Rank 0 runs the
single-GPU reference on its own device and compares (bench.slab_bitwise_check — the same check
bench.py puts into its `parity` object at every N).  Cases: NVT and GCMC; the peer-memory halo and the
ncclSend/ncclRecv fallback; re-uploads into live storage, with the odd ranks held back after the
capacity collective when the library was built with -DMC2D_FAULT_INJECTION (MC2D_LIB=.../libmc2d_fi.so);
a position outside the box on ONE rank must fail the upload on EVERY rank.  tests/test_gpu_multi.py
runs this file under torchrun when the box has at least two GPUs."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m
from montecarlo_simulation_b200 import api


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [("GCMC plain", dict(gcmc=True)),
             ("NVT", dict(gcmc=False, reuploads=3, tuning=dict(test_upload_lag_us=3000))),
             ("GCMC", dict(gcmc=True, reuploads=3, tuning=dict(test_upload_lag_us=3000, upload_sort=1))),
             ("GCMC nccl halo", dict(gcmc=True, tuning=dict(halo_nccl=1))),
             ("GCMC no overlap", dict(gcmc=True, tuning=dict(halo_overlap=0))),
             ("GCMC unfused", dict(gcmc=True, tuning=dict(fuse_passes=0))),
             ("GCMC grid barrier", dict(gcmc=True, tuning=dict(pipeline=0)))]
    for label, kw in cases:
        kw["tuning"] = dict(kw.get("tuning") or {}, spin_timeout_ms=3000)  # a protocol bug shows up as status 9 within seconds
        res = bench.slab_bitwise_check(m, api, dist, rank, world, local, **kw)
        if rank == 0:
            print("%-16s world=%d: %s -> %s" % (label, world, json.dumps(res), "OK" if res["ok"] else "MISMATCH"), flush=True)
            ok = ok and res["ok"]
    # a position outside the box on ONE rank must fail the upload on EVERY rank (no rank left waiting in a collective)
    n_side, L = 227, 321.0
    rng = np.random.default_rng(7)
    g = np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij"), -1).reshape(-1, 2).astype(float)
    xy = (g + 0.5) * (L / n_side) + rng.uniform(-0.2, 0.2, g.shape)
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
            raised = e.status == api.ERR_OUT_OF_BOX
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
