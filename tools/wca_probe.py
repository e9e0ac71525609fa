#!/usr/bin/env python
"""Config #4, WCA variant: 1M particles at rho = 0.5, rcut = 2^(1/6), NVT + GCMC sweeps; prints timing and the
energy bookkeeping check (running energy vs full recomputation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import montecarlo_simulation_b200 as m

n = int(os.environ.get("AB_PARTICLES", "1000000"))
z = float(os.environ.get("WCA_Z", "1.0"))
wl = bench.workload(1, 0, n)
rc = 2.0 ** (1.0 / 6.0)
with m.Context(wl["L"], wl["L"], rc, potential="WCA", T=1.0, activity=z, delta=0.1, seed=42,
               cell_margin=float(os.environ.get("WCA_MARGIN", "0"))) as ctx:
    ctx.upload(wl["xy"])
    for label, plan in (("NVT", dict(gc_per_cell=0)), ("GCMC", dict(gc_per_cell=1))):
        ctx.run_sweeps(50, **plan)
        ctx.set_kernel_timing(True)
        ctx.reset_counters()
        st = ctx.run_sweeps(20, **plan)
        i = ctx.info()
        ctx.set_kernel_timing(False)
        st2 = ctx.run_sweeps(20, **plan)
        ms = ctx.info().last_sweep_ms
        U, W, N = ctx.total_energy_virial(resync=False)
        st3 = ctx.get_stats() if hasattr(ctx, "get_stats") else st2
        print("%s: grid %dx%d K=%d G=%d  20 sweeps %.3f ms  k_sweep %.1f us  k_rebin %.1f us  energy %.1f us  "
              "moves/s %.3g  N=%d  U=%.6f running=%.6f acc=%s" %
              (label, i.nx, i.ny, i.cell_capacity, i.sweep_kernel_lanes, ms, i.k_sweep_ms / max(1, i.k_sweep_launches) * 1e3,
               i.k_rebin_ms / max(1, i.k_rebin_launches) * 1e3, ctx.info().last_energy_ms * 1e3,
               sum(st["attempted"]) / 20 * 20 / (ms * 1e-3), N, U, st2["energy"],
               [round(a / max(1, b), 3) for a, b in zip(st["accepted"], st["attempted"])]))
