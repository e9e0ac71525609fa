#!/usr/bin/env python
"""A/B timing of the sweep loop on config #4 (1M particles, GCMC LJ) under tuning switches (mc2d_tuning fields).
Usage: python tools/ab_sweep.py "" "pipeline=0" "sweep_lanes=2 pipeline=0" ...
Each argument is one variant (space-separated field=value pairs; "" = defaults).  Every variant
starts from the same equilibrated configuration and seed, so final N and U must agree between
variants that only change scheduling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

variants = sys.argv[1:] or [""]
sweeps = int(os.environ.get("AB_SWEEPS", "20"))
npart = int(os.environ.get("AB_PARTICLES", "1000000"))
wl = bench.workload(1, 0, npart)
base = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in os.environ.get("AB_BASE_TUNING", "").split()}  # for variant libraries
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=42,
               tuning=base) as ctx:
    ctx.upload(wl["xy"])
    ctx.run_sweeps(200, disp_per_particle=1, gc_per_cell=1)
    start = ctx.download()
for v in variants:
    tuning = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in v.split()}
    with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=43,
                   tuning=tuning) as ctx:
        ctx.set_kernel_timing(True)
        ctx.upload(start)
        ctx.run_sweeps(3, disp_per_particle=1, gc_per_cell=1)
        best = None
        for rep in range(3):
            st = ctx.run_sweeps(sweeps, disp_per_particle=1, gc_per_cell=1)
            i = ctx.info()
            row = (i.last_sweep_ms, i.k_sweep_ms / max(1, i.k_sweep_launches) * 1e3, i.k_rebin_ms / max(1, i.k_rebin_launches) * 1e3)
            if best is None or row[0] < best[0]:
                best = row
        ctx.set_kernel_timing(False)
        ctx.run_sweeps(3, disp_per_particle=1, gc_per_cell=1)
        st = ctx.run_sweeps(sweeps, disp_per_particle=1, gc_per_cell=1)
        untimed = ctx.info().last_sweep_ms
        U, W, N = ctx.total_energy_virial(resync=True)
        print("%-50s %d sweeps: %.3f ms (no per-kernel events: %.3f ms)  k_sweep %.1f us  k_rebin %.1f us  energy %.1f us  N=%d U=%.9f" %
              (v or "(defaults)", sweeps, best[0], untimed, best[1], best[2], ctx.info().last_energy_ms * 1e3, N, U), flush=True)
