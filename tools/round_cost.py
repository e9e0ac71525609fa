#!/usr/bin/env python
"""How the colour-pass time of k_sweep_p splits into a fixed part (launch, index chain, staging, write-back) and a
per-round part: time per pass for m displacement trials per particle and g GC draws per cell (config #4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

wl = bench.workload(1, 0, int(os.environ.get("AB_PARTICLES", "1000000")))
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=42) as ctx:
    ctx.upload(wl["xy"])
    ctx.run_sweeps(200, disp_per_particle=1, gc_per_cell=1)
    for disp, gc in ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (4, 0), (4, 1), (8, 0)):
        ctx.set_kernel_timing(True)
        ctx.run_sweeps(3, disp_per_particle=disp, gc_per_cell=gc)
        best = 1e9
        for _ in range(3):
            st = ctx.run_sweeps(10, disp_per_particle=disp, gc_per_cell=gc)
            i = ctx.info()
            best = min(best, i.k_sweep_ms / max(1, i.k_sweep_launches) * 1e3)
        print("disp_per_particle=%d gc_per_cell=%d: %.1f us per pass" % (disp, gc, best), flush=True)
        ctx.set_kernel_timing(False)
