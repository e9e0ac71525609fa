#!/usr/bin/env python
"""Finds the activity z at which GCMC 2D LJ (rcut 2.5, unshifted, T given) holds <rho> = target.
Run on a GPU box: `python tools/calibrate_z.py`.  Bisection on log z over short runs of a 90k-particle
system; prints the density reached for every z tried.  Used once to fix Z_LJ_T1_RHO05 in bench.py."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import montecarlo_simulation_b200 as m

POT = sys.argv[sys.argv.index("--potential") + 1] if "--potential" in sys.argv else "LennardJones"
RC = {"LennardJones": 2.5, "WCA": 2.0 ** (1.0 / 6.0)}[POT]


def run(z, T=1.0, rho=0.5, n_side=300, sweeps=600):
    L = n_side * math.sqrt(1 / rho)
    g = np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij"), -1).reshape(-1, 2).astype(float)
    xy = (g + 0.5) * (L / n_side) + np.random.default_rng(1).uniform(-0.2, 0.2, g.shape)
    with m.Context(L, L, RC, potential=POT, T=T, activity=z, delta=0.1, seed=9, cell_margin=-1.0 if POT == "WCA" else 0.0) as ctx:
        ctx.upload(xy)
        ctx.run_sweeps(sweeps, gc_per_cell=1)
        ns = [ctx.run_sweeps(20, gc_per_cell=1)["n_particles"] for _ in range(10)]
    return np.mean(ns) / (L * L)

if __name__ == "__main__":
    lo, hi = (0.02, 1.0) if POT == "LennardJones" else (0.5, 60.0)
    for it in range(11):
        z = math.sqrt(lo * hi)
        r = run(z)
        print("z=%.5f rho=%.5f" % (z, r), flush=True)
        if r > 0.5: hi = z
        else: lo = z
