#!/usr/bin/env python
"""bench.py — throughput of the 2D Monte Carlo hot path on B200 (BASELINE.json metric:
attempted MC moves/s and pair-evals/s in fp64, with achieved HBM GB/s).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU implementation

One STEP = `--sweeps-per-step` full GCMC sweeps (4 colour passes each: one displacement trial per
particle + one grand-canonical draw per cell) followed by one full-system energy+virial reduction
(the reference's per-sample work, MC.cpp:31-78 + logging.cpp:80-90).
Workload: N=1 -> BASELINE config #4 (GCMC 2D LJ, 1M particles, rho=0.5, rcut 2.5, T=1);
N>1 -> config #5 family (2M particles per GPU, square box, 1-D slabs in x, NCCL halo), weak scaling.
Inputs (1M-particle slot arrays, ~80 MB touched per pass plus the double buffer) stay resident in
HBM for `value`; `e2e` re-uploads the configuration from pinned host memory and downloads it every step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# activity that keeps <rho> ~ 0.5 for 2D LJ (rcut 2.5, unshifted) at T = 1.0; calibrated on the
# GPU with tools/calibrate_z.py
Z_LJ_T1_RHO05 = 0.497
# the same for WCA (rcut = 2^(1/6), the reference's literal r^2 < 1.2599 test): tools/calibrate_z.py --potential WCA
Z_WCA_T1_RHO05 = 10.87
T_BENCH, RCUT, RHO, DELTA = 1.0, 2.5, 0.5, 0.1
POTENTIALS = {
    # name -> (rcut, activity, cell_margin): WCA cells of exactly rcut hold 0.63 particles; -1 lets the library widen
    # them at the first upload (single rank) so that a cell holds ~2.5
    "LennardJones": (RCUT, Z_LJ_T1_RHO05, 0.0),
    "WCA": (2.0 ** (1.0 / 6.0), Z_WCA_T1_RHO05, -1.0),
}


def lattice_columns(i0, i1, n_y, a_x, a_y, seed, jitter=0.2):
    """jittered square lattice, generated column by column so that every rank produces identical
    particles for the same lattice column (SURVEY.md §8d config #4/#5 start recipe)."""
    cols = []
    for i in range(i0, i1):
        rng = np.random.default_rng([seed, i])
        j = np.arange(n_y)
        x = (i + 0.5) * a_x + rng.uniform(-jitter, jitter, n_y)
        y = (j + 0.5) * a_y + rng.uniform(-jitter, jitter, n_y)
        cols.append(np.stack([x, y], 1))
    return np.concatenate(cols) if cols else np.zeros((0, 2))


def default_per_gpu(n_gpus):
    return 1_000_000 if n_gpus == 1 else 2_000_000


def workload(n_gpus, rank, particles_per_gpu=None, rcut=RCUT):
    """Jittered square lattice at rho = 0.5.  A slab rank generates only the lattice
    columns of its slab (+1 on each side); the slab bounds come from the library's pure-host geometry.  The single-rank
    case needs no geometry call at all, so the reference arm builds its input without loading libmc2d.so."""
    per_gpu = particles_per_gpu or default_per_gpu(n_gpus)
    n_total = per_gpu * n_gpus
    L = math.sqrt(n_total / RHO)
    n_side = int(round(math.sqrt(n_total)))
    a = L / n_side
    i0, i1 = 0, n_side
    if n_gpus > 1:
        from montecarlo_simulation_b200 import api

        nx, ny, dx, dy = api.sweep_grid(L, L, rcut, n_gpus)
        col0, ncol = api.slab_columns(nx, n_gpus, rank)
        x0, x1 = col0 * dx, (col0 + ncol) * dx
        i0 = max(0, int(math.floor(x0 / a)) - 1)
        i1 = min(n_side, int(math.ceil(x1 / a)) + 1)
    xy = lattice_columns(i0, i1, n_side, a, a, 12345)
    xy[:, 0] %= L
    xy[:, 1] %= L
    return dict(L=L, n_total=n_side * n_side, xy=xy)


def config_dict(n_gpus, per_gpu, potential, sweeps_per_step):
    """The workload both arms are measured on: the SAME dict (keys and values) in the line of this repo's arm and in
    the line of the reference arm, so that the driver's same_config comparison holds."""
    rcut, z, _ = POTENTIALS[potential]
    n_side = int(round(math.sqrt(per_gpu * n_gpus)))
    tag = {"LennardJones": "lj", "WCA": "wca"}[potential]
    return {"workload": "gcmc_%s_%dk_per_gpu" % (tag, per_gpu // 1000), "particles": n_side * n_side,
            "particles_per_gpu": per_gpu, "potential": potential, "rcut": rcut, "T": T_BENCH, "rho": RHO,
            "activity": z, "delta": DELTA, "ensemble": "GCMC", "sweeps_per_step": sweeps_per_step,
            "step": "%d GCMC sweeps + 1 energy/virial reduction" % sweeps_per_step,
            "sweep": "N displacement trials + grand-canonical insert/delete trials (MC.cpp:36-52); on the GPU 4 colour "
                     "passes, per cell visit 1 displacement trial per particle + 1 GC draw",
            "decomposition": "1-D slabs in x, %d rank(s)" % n_gpus,
            "start": "jittered square lattice (seed 12345), equilibrated before timing",
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB write on the launch stream, outside the per-step "
                  "event pairs)"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML every 2 ms, or
    nvidia-smi every 0.2 s when pynvml is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def _nvml(self):
        """NVML handle for fast sampling (one query takes microseconds; nvidia-smi takes tens of milliseconds,
        longer than a timed step)"""
        try:
            import pynvml as nv

            nv.nvmlInit()
            return nv, nv.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None, None

    def run(self):
        nv, h = self._nvml()
        while not self.stop_flag.is_set():
            try:
                if nv is not None:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    act = lambda bit: "Active" if (r & bit) else "Not Active"
                    self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                      str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)),
                                      str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0),
                                      act(nv.nvmlClocksThrottleReasonHwSlowdown),
                                      act(nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                                      act(nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                                      act(nv.nvmlClocksThrottleReasonSwPowerCap)])
                    self.stop_flag.wait(0.002)
                    continue
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([f.strip() for f in out.split(",")])
            except Exception:
                nv = None
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class _native_stdout_to_stderr:
    """The reference's classes print progress and warnings with std::cout ("cell list update frequency is too high!");
    while its code runs, file descriptor 1 points at stderr so that stdout carries nothing but the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes

            ctypes.CDLL(None).fflush(None)  # the C++ side buffers when stdout is not a terminal
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref = unmodified serial_src compiled in place; the
    oracle port if that .so is absent) on one host core — the serial reference is single-threaded.
    Nothing of this repo's product is loaded here: the input lattice is pure numpy, the timed code is the reference's.
    `value` = the reference's move functions alone (N x stepCellList + grandCanonicalMove + cell rebuild every 10
    sweeps: no logging, no energy re-sync — the most favourable reading for it); `run_stock` = the stock
    MonteCarlo::run loop (MC.cpp:31-78) on the same sample; `mpi` = the reference's MPI code (mpi_src, unmodified)
    on the threads-as-ranks shim of oracle/mpi_shim, when that binary was built."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with _native_stdout_to_stderr():
        line = _reference_arm_line(args)
    print(json.dumps(line))


def _reference_arm_line(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    rcut, z, _ = POTENTIALS[args.potential]
    ptype = {"LennardJones": ol.LJ, "WCA": ol.WCA}[args.potential]
    per_gpu = args.particles_per_gpu or default_per_gpu(args.gpus)
    sample_n = args.ref_sample_particles
    wl = workload(1, 0, sample_n)
    L, xy = wl["L"], wl["xy"]
    run_stock = None
    if ol.have_ref():
        ref = ol.Ref()
        mc = ref.mc(L, L, xy, T=T_BENCH, ptype=ptype, rcut=rcut, mu=z, delta=DELTA, ensemble=ol.GCMC, seed=42)
        kind = "reference"

        def step():
            return ref.time_moves(mc, 1, 10, True)
    else:
        orc = ol.Oracle()
        mc = orc.mc(L, L, xy, orc.calc(T_BENCH, 0.0, ptype, rcut, z), rcut, DELTA, ol.GCMC, 42)
        kind = "port"

        def step():
            t0 = time.perf_counter()
            n = mc.n
            for _ in range(n):
                mc.step_celllist()
            mc.gc_move()
            mc.update_celllist()
            return time.perf_counter() - t0, n + 1
    for _ in range(args.warmup):
        step()
    secs = moves = 0
    for _ in range(args.steps):
        s, m = step()
        secs += s
        moves += m
    value = moves / secs
    if kind == "reference":  # the stock step loop: MonteCarlo::run (trial moves, cell rebuild every 10 sweeps, final sample)
        s2, m2 = ref.time_mc_run(mc, 3, 10)
        run_stock = {"value": m2 / s2, "unit": "moves/s", "sweeps": 3,
                     "what": "MonteCarlo::run(3, no intermediate output, cell update every 10) on the same sample; moves = "
                             "increase of the reference's own trial counter (MC.cpp:40,47)"}
    sample = ("1 serial sweep per step over a %d-particle cut of the workload (same density, lattice and seeds): N "
              "displacement trials + 1 GC trial, cell rebuild every 10 sweeps; move functions only" % len(xy))
    line = {
        "impl": "reference", "metric": "attempted_mc_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, per_gpu, args.potential, args.sweeps_per_step),
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "run_stock": run_stock,
        "cpu_baseline_mpi": {"skipped": "--no-mpi-baseline"} if args.no_mpi_baseline else mpi_baseline(args),
        "gpu_launches": 0,
    }
    return line


def mpi_baseline(args, ranks=(1, 2, 4, 8), n_part=10_000, sweeps=2):
    """The reference's MPI code path (mpi_src/*.cpp, unmodified) compiled against the test-only threads-as-ranks mpi.h of
    oracle/mpi_shim (this image has no MPI): NVT LJ, the only ensemble mpi_src implements, at 1/2/4/8 "ranks" = threads
    of one process, through the reference's own main() and input file.  A 10 000-particle system: the reference's MPI
    step rebuilds cells, migrates and refreshes ghosts every `cellUpdateFreq` ATTEMPTS (MC_parallel.cpp:88-99), O(N)
    work per attempt, so its cost per sweep grows with N^2 — 1 M particles would take hours per sweep.  Reported, never
    used as ground truth (SURVEY.md §7 defect 1: with Py = 1 its energies are wrong).  moves = 4 N per sweep
    (MC_parallel.cpp:69-104: N "attempts" of 4 parity micro-steps, one displacement trial each)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mc2d_ref_mpi")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/mc2d_ref_mpi not built (make -C oracle ref_mpi needs /root/reference)"}
    import tempfile

    out = {"kind": "reference MPI code (mpi_src, unmodified) on the threads-as-ranks shim oracle/mpi_shim, N threads",
           "ensemble": "NVT LennardJones rcut 2.5 T 1 rho 0.5 delta 0.1 (random start, the reference's initializer)",
           "particles": n_part, "sweeps": sweeps, "host_cores_available": os.cpu_count(), "runs": []}
    ncpu = os.cpu_count() or 1
    L = math.sqrt(n_part / RHO)
    for r in ranks:
        if r > ncpu:
            out["runs"].append({"ranks": r, "skipped": "more ranks than host cores (%d)" % ncpu})
            continue
        with tempfile.TemporaryDirectory() as td:
            inp = os.path.join(td, "input.txt")
            with open(inp, "w") as fh:
                fh.write("Lx %.6f\nLy %.6f\nN %d\nrcut %g\nT %g\nnSteps %d\neSteps 0\noutputFreq %d\n"
                         "cellUpdateFreq 10\npotential LennardJones\nout_xyz run.xyz\nout_data run.dat\ndelta %g\n"
                         "seed 42\nensemble NVT\n" % (L, L, n_part, RCUT, T_BENCH, sweeps, sweeps, DELTA))
            try:
                res = subprocess.run([exe, str(r), inp], cwd=td, capture_output=True, text=True, timeout=900)
            except subprocess.TimeoutExpired:
                out["runs"].append({"ranks": r, "error": "timeout"})
                continue
            rec = {"ranks": r, "rc": res.returncode}
            for ln in res.stdout.splitlines():
                if ln.startswith("MC2D_MPI_TIMING"):
                    kv = dict(f.split("=") for f in ln.split()[1:])
                    rec.update(run_s=float(kv["run_s"]), attempted=int(kv["attempted"]),
                               value=float(kv["attempted"]) / float(kv["run_s"]), unit="moves/s", cores=r)
            if res.returncode != 0:
                rec["stderr"] = res.stderr[-300:]
            out["runs"].append(rec)
    return out


def sort_rows(a):
    return a[np.lexsort((a[:, 1], a[:, 0]))]


def slab_bitwise_check(m, api, dist, rank, world, local, sweeps=25, gcmc=True, reuploads=0, tuning=None):
    """Driver-visible multi-GPU parity (the contract of unit_test_mpi/test_mc_parallel.cpp:136-187 and
    test_thermodynamic_calculator_parallel.cpp:378-514, tightened to bit equality): a 51 529-particle LJ system on a
    128-column grid is run for `sweeps` GCMC (or NVT) sweeps on `world` slabs — halos, migration through the ghost
    columns, the all-reduced energy — and, on rank 0's GPU, on ONE rank.  The Philox streams are keyed by the GLOBAL cell
    id and the colour order / grid shift are functions of (seed, sweep), so positions and counters must be EQUAL BIT FOR
    BIT and U, W equal to summation order.  world == 1: the production kernel k_sweep_p against the
    one-warp-per-cell kernel k_sweep."""
    n_side, L = 227, 321.0  # int(321 / 2.5) = 128 cell columns: the same grid for 1, 2, 4 and 8 slabs; rho = 0.5
    rng = np.random.default_rng(7)
    g = np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij"), -1).reshape(-1, 2).astype(float)
    xy = (g + 0.5) * (L / n_side) + rng.uniform(-0.2, 0.2, g.shape)
    kw = dict(activity=Z_LJ_T1_RHO05 if gcmc else 0.0)
    plan = dict(gc_per_cell=1 if gcmc else 0)

    def one_rank(tun):
        with m.Context(L, L, RCUT, T=T_BENCH, delta=DELTA, seed=11, device=local, tuning=tun, **kw) as one:
            one.upload(xy)
            u0 = one.total_energy_virial()[0]
            s1 = one.run_sweeps(sweeps, **plan)
            u1, w1, n1 = one.total_energy_virial()
            return dict(U0=u0, U1=u1, W1=w1, N1=n1, st=s1, xy=one.download())

    if world == 1:
        a = one_rank(tuning)
        b = one_rank(dict(sweep_lanes=0))
        got, ref, what = a, b, "k_sweep_p (production) vs k_sweep (one warp per cell), 1 GPU"
        st_all = a["st"]
        allp = a["xy"]
    else:
        ctx = m.Context(L, L, RCUT, T=T_BENCH, delta=DELTA, seed=11, device=local, rank=rank, nranks=world,
                        tuning=tuning, **kw)
        api.bootstrap_comm(ctx, dist)
        ctx.upload(xy)  # every rank passes the full set; the library keeps its slab
        for _ in range(reuploads):  # re-uploads into live storage (the end-to-end pattern)
            ctx.upload(xy)
        U0 = ctx.total_energy_virial(allreduce=True)[0]
        st = ctx.run_sweeps(sweeps, **plan)
        st_all = ctx.allreduce_stats(st)
        U1, W1, N1 = ctx.total_energy_virial(allreduce=True)
        halo_mode = ctx.info().halo_mode
        mine = ctx.download()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        ctx.close()
        if rank != 0:
            return None
        allp = np.concatenate(parts)
        got = dict(U0=U0, U1=U1, W1=W1, N1=N1, st=st_all, xy=allp)
        ref = one_rank(None)
        what = "%d slabs (halo mode %d) vs 1 GPU" % (world, halo_mode)
    rel = lambda x, y: abs(x - y) / max(1.0, abs(y))
    same_pos = got["xy"].shape == ref["xy"].shape and bool((sort_rows(got["xy"]) == sort_rows(ref["xy"])).all())
    same_cnt = got["st"]["attempted"] == ref["st"]["attempted"] and got["st"]["accepted"] == ref["st"]["accepted"]
    res = {"what": what, "particles": int(len(xy)), "sweeps": sweeps, "ensemble": "GCMC" if gcmc else "NVT",
           "positions_bitwise_equal": same_pos, "counters_equal": same_cnt, "N": [int(got["N1"]), int(ref["N1"])],
           "U_rel": rel(got["U1"], ref["U1"]), "W_rel": rel(got["W1"], ref["W1"]),
           "U_start_rel": rel(got["U0"], ref["U0"]),
           "running_energy_rel": rel(got["st"]["energy"], ref["st"]["energy"])}
    res["ok"] = bool(same_pos and same_cnt and got["N1"] == ref["N1"] and res["U_rel"] < 1e-12 and res["W_rel"] < 1e-12
                     and res["U_start_rel"] < 1e-12 and res["running_energy_rel"] < 1e-9)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sweeps-per-step", type=int, default=5)
    ap.add_argument("--particles-per-gpu", type=int, default=0)
    ap.add_argument("--potential", default="LennardJones", choices=sorted(POTENTIALS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip nselect2 / weak_ref_2000k / WCA side measurements")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--equil-sweeps", type=int, default=300)
    ap.add_argument("--ref-sample-particles", type=int, default=1_000_000,
                    help="--impl reference: particles in the cut of the workload one step sweeps (tests use a small one)")
    ap.add_argument("--no-mpi-baseline", action="store_true", help="skip the reference-MPI-on-shim side measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch

    import montecarlo_simulation_b200 as m
    from montecarlo_simulation_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
        args.gpus = world
    if api.device_count() < 1:
        sys.exit("bench.py needs a CUDA device; libmc2d has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def allmax(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- parity, before anything is timed: the slab path against one rank, bit for bit ------------------------------
    small = slab_bitwise_check(m, api, dist, rank, world, local)

    rcut, z_act, margin = POTENTIALS[args.potential]
    per_gpu = args.particles_per_gpu or default_per_gpu(world)
    wl = workload(world, rank, per_gpu, rcut)
    L = wl["L"]
    S = args.sweeps_per_step
    dev = torch.device("cuda", local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def make_ctx(Lbox, potential=args.potential, nranks=world, rk=rank):
        rc, za, mg = POTENTIALS[potential]
        c = m.Context(Lbox, Lbox, rc, potential=potential, T=T_BENCH, activity=za, delta=DELTA, seed=42, device=local,
                      rank=rk, nranks=nranks, cell_margin=mg if nranks == 1 else 0.0)
        if nranks > 1:
            api.bootstrap_comm(c, dist)
        return c

    def timed_steps(c, nsteps, allreduce, plan):
        """nsteps x (S sweeps + 1 reduction), each step inside its own event pair, L2 flushed in between; -> ms (max over
        ranks), last stats, last (U, W, N), summed pair counters"""
        stream = torch.cuda.ExternalStream(c.stream(), device=dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        acc = dict(evals=0, tests=0, energy_ms=0.0)
        barrier()
        st = uwn = None
        for k in range(nsteps):
            with torch.cuda.stream(stream):
                flush.zero_()
            ev[k][0].record(stream)
            st = c.run_sweeps(S, **plan)
            uwn = c.total_energy_virial(resync=True, allreduce=allreduce)
            ev[k][1].record(stream)
            info = c.info()
            acc["evals"] += info.last_pair_evals
            acc["tests"] += info.last_pair_tests
            acc["energy_ms"] += info.last_energy_ms
        barrier()
        return allmax(sum(a.elapsed_time(b) for a, b in ev)), st, uwn, acc

    plan1 = dict(disp_per_particle=1, gc_per_cell=1)
    ctx = make_ctx(L)
    ctx.upload(wl["xy"])
    # untimed equilibration: the jittered lattice start relaxes (N dips by ~10 % and recovers) before
    # anything is measured, so the timed steps run on a fluid at <rho> ~ 0.5
    if args.equil_sweeps > 0:
        ctx.run_sweeps(args.equil_sweeps, **plan1)
    for _ in range(args.warmup):
        ctx.run_sweeps(S, **plan1)
        ctx.total_energy_virial(resync=True, allreduce=world > 1)
    ctx.reset_counters()
    launches0 = ctx.info().launches
    sampler = ClockSampler(local)
    sampler.start()
    ms, st, (U, W, N), acc = timed_steps(ctx, args.steps, world > 1, plan1)
    clocks = sampler.summary()
    launches = ctx.info().launches - launches0
    info = ctx.info()
    st_sum = ctx.allreduce_stats(st) if world > 1 else st
    moves = float(sum(st_sum["attempted"]))
    n_now = st_sum["n_particles"]
    value = moves / (ms * 1e-3)
    # half-shell reduction: every undirected in-cutoff pair is evaluated once
    pair_rate = allsum(acc["evals"]) / (allmax(acc["energy_ms"]) * 1e-3) if acc["energy_ms"] > 0 else None
    pair_evals, pair_tests = acc["evals"], acc["tests"]

    # ---- parity on the timed system itself (every N): the running energy of the last timed step (S sweeps of accepted
    # dE on top of the previous resync) against the all-reduced recomputation, the particle count of the move counters
    # against the count of the reduction, and no refused insertion --------------------------------------------------------
    parity = {"small_case": small if rank == 0 else None,
              "running_vs_recomputed_energy_rel": abs(st_sum["energy"] - U) / max(1.0, abs(U)),
              "n_particles_counters_vs_reduction": [int(n_now), int(N)],
              "insertions_refused": int(st_sum["overflow"])}
    if world > 1:
        parity["bitwise_vs_1gpu"] = bool(small["ok"]) if rank == 0 else None
    else:
        parity["bitwise_vs_warp_per_cell_kernel"] = bool(small["ok"])

    extra = {}
    if not args.no_extra_legs:
        # ---- the same step with two displacement trials per particle and cell visit (HOOMD's nselect = 2) -----------
        # A colour pass has a fixed cost (launch, index chain, staging of the 9 cells) that 1 trial per particle pays
        # in full; more trials per visit are a legal schedule of the same ensemble (detailed balance holds per trial).
        # Reported next to `value`, which stays at the reference's sweep: N displacement trials + the GC draws.
        plan2 = dict(disp_per_particle=2, gc_per_cell=1)
        for _ in range(2):
            ctx.run_sweeps(S, **plan2)
        ctx.reset_counters()
        n2 = max(3, args.steps // 2)
        ms2, st2, _, _ = timed_steps(ctx, n2, world > 1, plan2)
        extra["nselect2"] = {"value": allsum(sum(st2["attempted"])) / (ms2 * 1e-3), "unit": "moves/s", "steps": n2,
                             "ms_per_step": ms2 / n2,
                             "what": "same step with disp_per_particle = 2 (two displacement trials per particle per cell visit)"}
        ctx.run_sweeps(S, **plan1)

    # ---- roofline of the dominant kernel (k_sweep), CUDA events around every launch ----------
    ctx.set_kernel_timing(True)
    ctx.reset_counters()
    st_r = ctx.run_sweeps(S, **plan1)
    info = ctx.info()
    ctx.set_kernel_timing(False)
    nl = max(info.k_sweep_launches, 1)
    k_ms = info.k_sweep_ms / nl
    cells_per_pass = (info.ncol // 2) * (info.ny // 2)
    # algorithmic bytes: 32 B per attempted displacement (read + write 16 B),
    # 20 B per attempted insert/delete (16 B slot + 4 B count), 8 B per active cell (start + count)
    alg_bytes = (32.0 * st_r["attempted"][0] + 20.0 * (st_r["attempted"][1] + st_r["attempted"][2])) / nl + 8.0 * cells_per_pass
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_sweep_p<%s, G=%d> (one colour pass)" % (args.potential, info.sweep_kernel_lanes),
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes, "ms_per_launch": k_ms,
                "share_of_sweep_time": info.k_sweep_ms / max(info.last_sweep_ms, 1e-9),
                "moves_per_s_kernel_only": sum(st_r["attempted"]) / nl / (k_ms * 1e-3),
                "rebin_ms_per_launch": info.k_rebin_ms / max(info.k_rebin_launches, 1),
                "halo_ms_per_sync": (info.k_halo_ms / info.k_halo_launches) if info.k_halo_launches else None,
                "note": "path is fp64-ALU/latency bound, not HBM bound (DESIGN.md §4): low HBM fraction expected"}
    # ---- compute roofline (fp64): algorithmic flops of one colour pass against the MEASURED fp64 FMA peak ----
    # Algorithmic flops: a displacement evaluates two positions, an insertion / deletion one; a
    # position costs (distance tests of the 3x3 neighbourhood) x 8 flop + (pairs inside the cutoff) x 14 flop.  The
    # per-particle test and pair counts come from the energy kernel's counters of the timed steps (its half shell
    # visits every unordered pair once, so a particle's full neighbourhood is twice the per-particle average).
    fp64_roof = None
    try:
        peak64 = ctx.measure_fp64_peak()
        n_loc_now = max(ctx.num_particles(), 1)
        tests_pp = 2.0 * pair_tests / max(args.steps, 1) / n_loc_now
        evals_pp = 2.0 * pair_evals / max(args.steps, 1) / n_loc_now
        flop_pos = 8.0 * tests_pp + 14.0 * evals_pp
        flops_launch = (2.0 * flop_pos * st_r["attempted"][0] + flop_pos * (st_r["attempted"][1] + st_r["attempted"][2])) / nl
        ach64 = flops_launch / (k_ms * 1e-3) / 1e12
        fp64_roof = {"bound": "fp64", "achieved": ach64, "peak": peak64, "unit": "TFLOP/s", "frac": ach64 / peak64,
                     "peak_source": "measured here (mc2d_measure_fp64_peak: independent DFMA chains on every SM)",
                     "flop_per_position": flop_pos, "distance_tests_per_position": tests_pp,
                     "pairs_in_cutoff_per_position": evals_pp,
                     "note": "algorithmic flops only; masked lanes, the second Newton step, Philox and the Metropolis "
                             "test are not counted"}
    except Exception as e:  # the HBM roofline above is the contract; this one is an explanation
        fp64_roof = {"error": str(e)}

    # DRAM bytes per launch of the sweep kernel from the newest committed `ncu --set full` capture
    import glob
    trs = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")))  # rNNx names sort by round and capture
    if trs:
        try:
            roofline["traffic"] = json.load(open(trs[-1])).get("k_sweep_dram_bytes_per_launch")
            roofline["traffic_source"] = "profiles/" + os.path.basename(trs[-1]) + " (config #4, 1 GPU)"
        except Exception:
            pass

    # ---- e2e: through the C ABI with HOST buffers, copies inside the timed region ------------------
    n_loc = ctx.num_particles()
    host = torch.empty((int(n_loc * 1.2) + 1024, 2), dtype=torch.float64, pin_memory=True).numpy()
    n_host = ctx.download_into(host)
    h2d = d2h = 0
    moves_e2e = 0
    ctx.upload(host[:n_host])  # one untimed round trip (first-touch of the pinned buffer, allocator warm-up)
    ctx.download_into(host)
    barrier()
    allocs0 = ctx.info().slot_allocs
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ctx.reset_counters()
        # pinned host -> device (chunked, binning overlaps the copy), cell build; no initial-energy pass: the step
        # re-syncs the energy after its sweeps (mc2d_upload_ex, MC2D_UPLOAD_NO_ENERGY)
        ctx.upload(host[:n_host], initial_energy=False)
        h2d += 16 * n_host
        st_e = ctx.run_sweeps(S, **plan1)
        # one sample point: energy/virial reduction (re-sync) + device -> pinned host (the caller's particle vector); the
        # reduction runs while the positions are on the bus (mc2d_sample)
        Ue, We, Ne, n_host = ctx.sample_into(host, resync=True, allreduce=world > 1)
        d2h += 16 * n_host + 24 + 80  # positions + {U, W, N} + the stats block
        moves_e2e += sum(st_e["attempted"])
    barrier()
    dt = allmax(time.perf_counter() - t0)
    e2e = {"value": allsum(moves_e2e) / dt, "unit": "moves/s", "h2d_bytes_per_step": h2d // args.e2e_steps,
           "d2h_bytes_per_step": d2h // args.e2e_steps, "steps": args.e2e_steps,
           "ms_per_step": 1e3 * dt / args.e2e_steps, "slot_reallocations": ctx.info().slot_allocs - allocs0,
           "what": "mc2d_upload_ex(host xy, NO_ENERGY) + %d sweeps + mc2d_sample (energy/virial with re-sync + download to host xy) "
                   "per step; copies chunked and overlapped with the binning / compaction / reduction kernels" % S}
    halo_mode = ctx.info().halo_mode
    grid_now = [info.nx, info.ny]
    cellcap = info.cell_capacity
    slot_mb = info.ncol * info.ny * info.cell_capacity * 16 / 1e6
    ctx.close()

    # ---- side measurements at N = 1: the per-GPU load of the multi-GPU runs on ONE GPU (the like-for-like anchor of the
    # weak-scaling efficiency), and the WCA variant of config #4 -----------------------------------------------------------
    if world == 1 and not args.no_extra_legs:
        def side_leg(potential, npart, nsteps=5):
            rc_, _, _ = POTENTIALS[potential]
            w2 = workload(1, 0, npart, rc_)
            c2 = make_ctx(w2["L"], potential, 1, 0)
            c2.upload(w2["xy"])
            c2.run_sweeps(args.equil_sweeps, **plan1)
            for _ in range(3):
                c2.run_sweeps(S, **plan1)
                c2.total_energy_virial(resync=True)
            c2.reset_counters()
            ms_, st_, (U_, W_, N_), _ = timed_steps(c2, nsteps, False, plan1)
            i2 = c2.info()
            out = {"value": sum(st_["attempted"]) / (ms_ * 1e-3), "unit": "moves/s", "ms_per_step": ms_ / nsteps,
                   "steps": nsteps, "particles_now": int(N_), "grid": [i2.nx, i2.ny], "cell_capacity": i2.cell_capacity,
                   "running_vs_recomputed_energy_rel": abs(st_["energy"] - U_) / max(1.0, abs(U_)),
                   "config": config_dict(1, npart, potential, S)}
            c2.close()
            return out

        if args.potential == "LennardJones" and per_gpu != 2_000_000:
            extra["weak_ref_2000k"] = side_leg("LennardJones", 2_000_000)
            extra["weak_ref_2000k"]["what"] = ("1 GPU at the per-GPU load of the N > 1 runs (2 M particles): divide the "
                                               "N-GPU ms_per_step by this one for the weak-scaling efficiency")
        if args.potential == "LennardJones":
            extra["wca_1000k"] = side_leg("WCA", 1_000_000)
            extra["wca_1000k"]["what"] = "config #4 with WCA (rcut = 2^(1/6), literal r^2 < 1.2599 test), same step"

    # ---- CPU baseline: the reference itself on this box's host cores (rank 0, N=1 only) -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol

        ptype = {"LennardJones": ol.LJ, "WCA": ol.WCA}[args.potential]
        sample_n = 200_000
        wl_s = workload(1, 0, sample_n)
        if ol.have_ref():
            with _native_stdout_to_stderr():
                ref = ol.Ref()
                mcr = ref.mc(wl_s["L"], wl_s["L"], wl_s["xy"], T=T_BENCH, ptype=ptype, rcut=rcut, mu=z_act,
                             delta=DELTA, ensemble=ol.GCMC, seed=42)
                secs, mv = ref.time_moves(mcr, 25, 10, True)
                es, Ur, Wr = ref.time_energy_virial(wl_s["L"], wl_s["L"], wl_s["xy"], T=T_BENCH, ptype=ptype, rcut=rcut)
            kind = "reference"
            # the GPU calculator hook on the same 200k-particle sample against the reference's numbers
            with m.Context(wl_s["L"], wl_s["L"], rcut, potential=args.potential, T=T_BENCH, device=local) as cc:
                Ug, Wg = cc.calc_energy_virial(wl_s["xy"])
            parity["calculator_vs_reference_rel"] = {"U": abs(Ug - Ur) / max(1.0, abs(Ur)), "W": abs(Wg - Wr) / max(1.0, abs(Wr)),
                                                     "particles": sample_n}
        else:
            orc = ol.Oracle()
            mco = orc.mc(wl_s["L"], wl_s["L"], wl_s["xy"], orc.calc(T_BENCH, 0.0, ptype, rcut, z_act), rcut,
                         DELTA, ol.GCMC, 42)
            t0 = time.perf_counter()
            mco.run(25, 1 << 40, 1 << 40)
            secs, mv = time.perf_counter() - t0, 25 * (sample_n + 1)
            es, kind = None, "port"
        cpu = {"value": mv / secs, "unit": "moves/s", "cores": 1, "kind": kind,
               "sample": "25 serial sweeps (N displacement trials + 1 GC trial each) of a %d-particle cut of the same "
                         "workload, move functions only (no logging / energy re-sync)" % sample_n,
               "energy_virial_particles_per_s": (sample_n / es) if es else None, "host_cores_available": os.cpu_count()}

    parity["ok"] = bool((small["ok"] if rank == 0 else True) and parity["running_vs_recomputed_energy_rel"] < 1e-9
                        and parity["n_particles_counters_vs_reduction"][0] == parity["n_particles_counters_vs_reduction"][1]
                        and parity["insertions_refused"] == 0
                        and all(v < 1e-10 for k, v in parity.get("calculator_vs_reference_rel", {}).items() if k in "UW"))
    if rank == 0:
        n_disp, n_gc = st_sum["attempted"][0], st_sum["attempted"][1] + st_sum["attempted"][2]
        line = {
            "metric": "attempted_mc_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world, per_gpu, args.potential, S),
            "run": {"particles_now": int(n_now), "grid": grid_now, "cell_capacity": cellcap,
                    "halo": {0: "none (single rank)", 1: "ncclSend/ncclRecv of whole slot columns",
                             2: "live slots stored into the neighbours' ghost columns over NVLink (CUDA IPC peer "
                                "memory) from inside the sweep kernel, flag-word sync"}[halo_mode],
                    "slot_array_mb_per_buffer": slot_mb},
            # `value` counts displacement AND grand-canonical trials (a GC trial evaluates one position, a displacement
            # two); the reference's sweep is N displacements + ONE GC trial, ours N + ~0.6 per cell
            "displacement_moves_per_s": value * n_disp / max(n_disp + n_gc, 1),
            "gc_moves_per_s": value * n_gc / max(n_disp + n_gc, 1),
            "trials_per_step": {"displacement": int(n_disp), "insert_delete": int(n_gc)},
            "pair_evals_per_s": pair_rate, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "roofline_fp64": fp64_roof, "parity": parity,
            "acceptance": {"displacement": st_sum["accepted"][0] / max(st_sum["attempted"][0], 1),
                           "insertion": st_sum["accepted"][1] / max(st_sum["attempted"][1], 1),
                           "deletion": st_sum["accepted"][2] / max(st_sum["attempted"][2], 1)},
        }
        line.update(extra)
        if cpu is not None and not args.no_cpu_baseline:
            line["cpu_baseline_mpi"] = mpi_baseline(args)
        print(json.dumps(line))
    ok = torch.tensor([1 if (parity["ok"] or rank != 0) else 0], device="cuda")
    if dist:
        dist.broadcast(ok, src=0)
        dist.destroy_process_group()
    if int(ok.item()) == 0:
        sys.exit("bench.py: PARITY CHECK FAILED (see the `parity` object of the line above)")


if __name__ == "__main__":
    main()
