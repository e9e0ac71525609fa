#!/usr/bin/env python
"""Prints the worst relative error of the sweep kernel's per-trial dE against the CPU oracle
(same replay as tests/test_gpu_parity.py::test_per_trial_dE_matches_oracle) — used to decide how
many Newton steps the sweep kernel's reciprocal needs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import montecarlo_simulation_b200 as m
import oracle_lib as ol
from test_gpu_parity import lattice, replay_trace

oracle = ol.Oracle()
Lx, Ly = 27.0, 24.0
xy = lattice(18, 16, Lx, Ly, 0.2, 77)
for name, ptype, rcut, z in (("LennardJones", ol.LJ, 2.5, 0.05), ("WCA", ol.WCA, 2 ** (1 / 6), 0.4)):
    calc = oracle.calc(0.9, 0.0, ptype, rcut, z)
    with m.Context(Lx, Ly, rcut, potential=name, T=0.9, activity=z, delta=0.12, seed=5) as ctx:
        ctx.upload(xy)
        ctx.run_sweeps(3, gc_per_cell=1)
        cfg = ctx.download()
        trials, st = ctx.run_sweep_traced(gc_per_cell=2)
        worst, checked, final = replay_trace(oracle, calc, Lx, Ly, cfg, trials, rcut)
        U, W = ctx.calc_energy_virial(cfg)
        Uo = oracle.total_energy_celllist(calc, Lx, Ly, cfg)
        print("%s: %d trials checked, worst per-trial dE error %.3e (bar 1e-10); U rel err %.3e" %
              (name, checked, worst, abs(U - Uo) / max(1, abs(Uo))))
