#!/usr/bin/env python
"""Soak test: many GCMC sweeps of config #4; the running energy must track a full recomputation, no insertion may be
refused, N must stay near its equilibrium value."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

nsweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
wl = bench.workload(1, 0, 1_000_000)
t0 = time.time()
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=7) as ctx:
    ctx.upload(wl["xy"])
    worst = 0.0
    for k in range(nsweeps // 500):
        st = ctx.run_sweeps(500, disp_per_particle=1, gc_per_cell=1)
        U, W, N = ctx.total_energy_virial(resync=False)
        rel = abs(st["energy"] - U) / abs(U)
        worst = max(worst, rel)
        print("sweeps %5d  N=%d  U/N=%.6f  running-vs-recomputed %.2e  overflow %d  acc %.3f  (%.1f s)" %
              ((k + 1) * 500, N, U / N, rel, st["overflow"], st["accepted"][0] / st["attempted"][0], time.time() - t0), flush=True)
    assert worst < 1e-9 and st["overflow"] == 0
    print("OK")
