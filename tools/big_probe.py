#!/usr/bin/env python
"""Robustness probes at sizes / densities the parity tests do not reach: N particles at density rho on one GPU,
GCMC sweeps, running energy vs recomputation, timing.  Usage: big_probe.py N rho [potential]"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import montecarlo_simulation_b200 as m

n = int(float(sys.argv[1])); rho = float(sys.argv[2]); pot = sys.argv[3] if len(sys.argv) > 3 else "LennardJones"
side = int(round(math.sqrt(n))); n = side * side
L = math.sqrt(n / rho)
rng = np.random.default_rng(1)
g = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2).astype(float)
a = L / side
xy = (g + 0.5) * a + rng.uniform(-0.2 * min(a, 1.0), 0.2 * min(a, 1.0), g.shape)
kw = dict(f=8.0, alpha=0.3) if "Star" in pot else {}
t0 = time.time()
with m.Context(L, L, 2.5, potential=pot, T=1.0, activity=0.497, delta=0.1, seed=5, **kw) as ctx:
    ctx.upload(xy)
    i = ctx.info()
    st = ctx.run_sweeps(5, gc_per_cell=1)
    st = ctx.run_sweeps(10, gc_per_cell=1)
    ms = ctx.info().last_sweep_ms
    U, W, N = ctx.total_energy_virial(resync=False)
    out = ctx.download()
    print("%s N0=%d rho=%.2f grid %dx%d K=%d lanes=%d: 10 sweeps %.2f ms (%.3g moves/s), N=%d, U=%.6f running=%.6f rel=%.1e, "
          "energy %.2f ms, download %d rows, wall %.1f s" %
          (pot, n, rho, i.nx, i.ny, i.cell_capacity, ctx.info().sweep_kernel_lanes, ms, sum(st["attempted"]) / 10 * 10 / (ms * 1e-3) / 1.5,
           N, U, st["energy"], abs(U - st["energy"]) / max(1, abs(U)), ctx.info().last_energy_ms, len(out), time.time() - t0))
