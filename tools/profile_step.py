#!/usr/bin/env python
"""Short driver for ncu: config #4 (1M particles, GCMC LJ): upload, a few sweeps, one energy reduction, download
and a re-upload (the end-to-end pattern; the second upload takes the sort-free path).
Usage: python tools/profile_step.py [sweeps]   (PROFILE_TUNING="fuse_passes=0 ..." sets mc2d_tuning fields: with
fuse_passes=0 one k_sweep_p launch is one colour pass, the unit of bench.py's roofline)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dpc = int(os.environ.get("MC2D_DISP_PER_CELL", "0"))
tuning = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in os.environ.get("PROFILE_TUNING", "").split()}
wl = bench.workload(1, 0, 1_000_000)
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=42,
               tuning=tuning) as ctx:
    ctx.upload(wl["xy"])
    st = ctx.run_sweeps(sweeps, disp_per_particle=1, gc_per_cell=1, disp_per_cell=dpc)
    U, W, N = ctx.total_energy_virial(resync=True)
    xy = ctx.download()
    ctx.upload(xy)
    print("N=%d U=%.6f moves=%d sweep_ms=%.3f energy_ms=%.3f" % (N, U, sum(st["attempted"]), ctx.info().last_sweep_ms, ctx.info().last_energy_ms))
