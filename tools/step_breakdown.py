#!/usr/bin/env python
"""Where a bench step (5 sweeps + 1 reduction, config #4) spends its time: device time of the sweep batch
(events inside mc2d_run_sweeps), device time of the reduction, and the wall clock of the two calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import montecarlo_simulation_b200 as m

wl = bench.workload(1, 0, 1000000)
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=42) as ctx:
    ctx.upload(wl["xy"])
    ctx.run_sweeps(100, disp_per_particle=1, gc_per_cell=1)
    for with_stats in (True,):
        rows = []
        for step in range(30):
            t0 = time.perf_counter()
            ctx.run_sweeps(5, disp_per_particle=1, gc_per_cell=1)
            t1 = time.perf_counter()
            ctx.total_energy_virial(resync=True)
            t2 = time.perf_counter()
            i = ctx.info()
            rows.append(((t1 - t0) * 1e3, i.last_sweep_ms, (t2 - t1) * 1e3, i.last_energy_ms))
        r = np.array(rows[5:])
        print("stats=%s  run_sweeps wall %.3f ms (device %.3f)  energy wall %.3f ms (device %.3f)  step %.3f ms" %
              (with_stats, r[:, 0].mean(), r[:, 1].mean(), r[:, 2].mean(), r[:, 3].mean(), (r[:, 0] + r[:, 2]).mean()))
