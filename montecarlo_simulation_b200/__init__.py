"""montecarlo_simulation_b200 — B200-native (sm_100a) implementation of the 2D Monte Carlo hot path
of ReyhanehFarimani/MonteCarlo_simulation: checkerboard trial displacements, GCMC insert/delete
energy evaluation against the cell list, and full-system energy/virial reduction.

Layout:
  csrc/   hand-written CUDA kernels + the C ABI (include/mc2d.h) -> libmc2d.so
  host/   C++ shims that keep the reference's class interfaces (CellList, ThermodynamicCalculator,
          MonteCarlo, Input, Logging) on top of the C ABI, and the input.txt driver
  api.py  ctypes binding used by tests/ and bench.py (plumbing only)
"""
from .api import (ATHERMAL_STAR, IDEAL, LENNARD_JONES, THERMAL_STAR, WCA, YUKAWA, Context, MC2DError,  # noqa: F401
                  load_library)
