#!/usr/bin/env python
"""Per-kernel timing of the slab path under torchrun (2 M particles per GPU, GCMC LJ): where the N-GPU sweep differs from
the 1-GPU sweep at the same load.  Usage: torchrun --nproc-per-node 2 tools/ab_slab.py "" "pipeline=0" ...
Each rank prints its own line (events are per device)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import montecarlo_simulation_b200 as m
from montecarlo_simulation_b200 import api

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
variants = sys.argv[1:] or [""]
sweeps = int(os.environ.get("AB_SWEEPS", "20"))
wl = bench.workload(world, rank, 2000000)
if os.environ.get("AB_SHAPE"):  # "Lx,Ly": one rank on a box of another shape (e.g. the slab's), same density
    import numpy as np
    Lx, Ly = (float(v) for v in os.environ["AB_SHAPE"].split(","))
    nx_, ny_ = int(round((0.5 * Lx * Ly) ** 0.5 * (Lx / Ly) ** 0.5)), int(round((0.5 * Lx * Ly) ** 0.5 * (Ly / Lx) ** 0.5))
    g_ = np.stack(np.meshgrid(np.arange(nx_), np.arange(ny_), indexing="ij"), -1).reshape(-1, 2).astype(float)
    wl = dict(Lx=Lx, Ly=Ly, xy=(g_ + 0.5) * [Lx / nx_, Ly / ny_] + np.random.default_rng(5).uniform(-0.2, 0.2, g_.shape))
for v in variants:
    tuning = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in v.split()}
    tuning.setdefault("spin_timeout_ms", 3000)
    with m.Context(wl["Lx"] if "Lx" in wl else wl["L"], wl["Ly"] if "Ly" in wl else wl["L"], bench.RCUT, T=bench.T_BENCH,
                   activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=43, device=local, rank=rank, nranks=world,
                   tuning=tuning) as ctx:
        if world > 1:
            api.bootstrap_comm(ctx, dist)
        ctx.upload(wl["xy"])
        ctx.run_sweeps(40, disp_per_particle=1, gc_per_cell=1)
        ctx.set_kernel_timing(True)
        ctx.run_sweeps(3, disp_per_particle=1, gc_per_cell=1)
        best = None
        for rep in range(3):
            ctx.run_sweeps(sweeps, disp_per_particle=1, gc_per_cell=1)
            i = ctx.info()
            row = (i.last_sweep_ms, i.k_sweep_ms / max(1, i.k_sweep_launches) * 1e3, i.k_rebin_ms / max(1, i.k_rebin_launches) * 1e3,
                   i.k_halo_ms / max(1, i.k_halo_launches) * 1e3, i.k_halo_launches)
            if best is None or row[0] < best[0]:
                best = row
        ctx.set_kernel_timing(False)
        ctx.run_sweeps(3, disp_per_particle=1, gc_per_cell=1)
        ctx.run_sweeps(sweeps, disp_per_particle=1, gc_per_cell=1)
        untimed = ctx.info().last_sweep_ms
        U, W, N = ctx.total_energy_virial(resync=True, allreduce=world > 1)
        print("rank %d %-28s %d sweeps: %.3f ms (no per-kernel events: %.3f ms = %.1f us/sweep)  k_sweep %.1f us/pass  k_rebin %.1f us  halo %.1f us x%d  energy %.1f us  N=%d" %
              (rank, v or "(defaults)", sweeps, best[0], untimed, untimed / sweeps * 1e3, best[1], best[2], best[3], best[4],
               ctx.info().last_energy_ms * 1e3, N), flush=True)
if world > 1:
    dist.destroy_process_group()
