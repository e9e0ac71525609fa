#!/usr/bin/env python
"""Runs one sweep of config #4 on a library built with -DMC2D_PROFILE_WAIT (pass it as MC2D_LIB): a few warps print
their total cycles, the cycles spent waiting for the cp.async staging and for the previous colour (grid barrier or
completion counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

wl = bench.workload(1, 0, 1_000_000)
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=42) as ctx:
    ctx.upload(wl["xy"])
    ctx.run_sweeps(1, disp_per_particle=1, gc_per_cell=1)
