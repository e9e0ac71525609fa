#!/usr/bin/env python
"""Wall-clock breakdown of one end-to-end step (config #4): upload / sweeps / energy / download."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

wl = bench.workload(1, 0, 1_000_000)
with m.Context(wl["L"], wl["L"], bench.RCUT, T=bench.T_BENCH, activity=bench.Z_LJ_T1_RHO05, delta=bench.DELTA, seed=42) as ctx:
    host = torch.empty((1_300_000, 2), dtype=torch.float64, pin_memory=True).numpy()
    n = len(wl["xy"])
    host[:n] = wl["xy"]
    for it in range(4):
        t = [time.perf_counter()]
        ctx.upload(host[:n]); t.append(time.perf_counter())
        st = ctx.run_sweeps(5, gc_per_cell=1); t.append(time.perf_counter())
        U, W, N = ctx.total_energy_virial(resync=True); t.append(time.perf_counter())
        n = ctx.download_into(host); t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        print("upload %.2f ms  sweeps %.2f ms  energy %.2f ms  download %.2f ms  total %.2f ms  N=%d" % (*d, d.sum(), n))
