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
T_BENCH, RCUT, RHO, DELTA = 1.0, 2.5, 0.5, 0.1


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


def workload(n_gpus, rank, particles_per_gpu=None):
    from montecarlo_simulation_b200 import api

    per_gpu = particles_per_gpu or (1_000_000 if n_gpus == 1 else 2_000_000)
    n_total = per_gpu * n_gpus
    L = math.sqrt(n_total / RHO)
    n_side = int(round(math.sqrt(n_total)))
    a = L / n_side
    nx, ny, dx, dy = api.sweep_grid(L, L, RCUT, n_gpus)
    col0, ncol = api.slab_columns(nx, n_gpus, rank)
    x0, x1 = col0 * dx, (col0 + ncol) * dx
    i0 = max(0, int(math.floor(x0 / a)) - 1)
    i1 = min(n_side, int(math.ceil(x1 / a)) + 1)
    xy = lattice_columns(i0, i1, n_side, a, a, 12345)
    xy[:, 0] %= L
    xy[:, 1] %= L
    return dict(L=L, n_total=n_side * n_side, xy=xy, name="gcmc_lj_%dk_per_gpu" % (per_gpu // 1000))


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


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref = unmodified serial_src
    compiled in place; the oracle port if that .so is absent) on one host core — the reference
    is single-threaded and its MPI path cannot run here (no MPI)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    wl = workload(1, 0, 1_000_000)
    L, xy = wl["L"], wl["xy"]
    if ol.have_ref():
        ref = ol.Ref()
        mc = ref.mc(L, L, xy, T=T_BENCH, ptype=ol.LJ, rcut=RCUT, mu=Z_LJ_T1_RHO05, delta=DELTA, ensemble=ol.GCMC, seed=42)
        kind = "reference"

        def step():
            return ref.time_moves(mc, 1, 10, True)
    else:
        orc = ol.Oracle()
        mc = orc.mc(L, L, xy, orc.calc(T_BENCH, 0.0, ol.LJ, RCUT, Z_LJ_T1_RHO05), RCUT, DELTA, ol.GCMC, 42)
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
    sample = "1 serial sweep per step over the 1M-particle config #4 (N displacement trials + 1 GC trial, cell rebuild)"
    line = {
        "impl": "reference", "metric": "attempted_mc_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "gcmc_lj_1000k_per_gpu", "particles": int(len(xy)), "potential": "LennardJones",
                   "rcut": RCUT, "T": T_BENCH, "rho": RHO, "activity": Z_LJ_T1_RHO05, "delta": DELTA},
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sweeps-per-step", type=int, default=5)
    ap.add_argument("--particles-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--equil-sweeps", type=int, default=300)
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

    wl = workload(world, rank, args.particles_per_gpu or None)
    L = wl["L"]
    ctx = m.Context(L, L, RCUT, potential="LennardJones", T=T_BENCH, activity=Z_LJ_T1_RHO05, delta=DELTA, seed=42,
                    device=local, rank=rank, nranks=world)
    if world > 1:
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0])
    ctx.upload(wl["xy"])
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    S = args.sweeps_per_step

    def step():
        st = ctx.run_sweeps(S, disp_per_particle=1, gc_per_cell=1)
        U, W, N = ctx.total_energy_virial(resync=True, allreduce=world > 1)
        return st, U, W, N

    # untimed equilibration: the jittered lattice start relaxes (N dips by ~10 % and recovers) before
    # anything is measured, so the timed steps run on a fluid at <rho> ~ 0.5
    if args.equil_sweeps > 0:
        ctx.run_sweeps(args.equil_sweeps, disp_per_particle=1, gc_per_cell=1)
    for _ in range(args.warmup):
        step()
    ctx.reset_counters()
    launches0 = ctx.info().launches
    sampler = ClockSampler(local)
    sampler.start()
    # L2 (126 MB) is flushed between timed steps by writing a 256 MiB buffer on the same stream;
    # every step is bracketed by its own event pair so the flush is outside the timed sum.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    pair_evals = pair_tests = 0
    energy_ms = 0.0
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        ev[k][0].record(stream)
        st, U, W, N = step()
        ev[k][1].record(stream)
        info = ctx.info()
        pair_evals += info.last_pair_evals
        pair_tests += info.last_pair_tests
        energy_ms += info.last_energy_ms
    barrier()
    ms = allmax(sum(a.elapsed_time(b) for a, b in ev))
    clocks = sampler.summary()
    launches = ctx.info().launches - launches0
    moves_local = sum(st["attempted"])
    moves = allsum(moves_local)
    n_now = allsum(st["n_particles"])
    value = moves / (ms * 1e-3)
    # half-shell reduction: every undirected in-cutoff pair is evaluated once
    pair_rate = allsum(pair_evals) / (allmax(energy_ms) * 1e-3) if energy_ms > 0 else None

    # ---- the same step with two displacement trials per particle and cell visit (HOOMD's nselect = 2) -----------
    # A colour pass has a fixed cost (launch, index chain, staging of the 9 cells) that 1 trial per particle pays
    # in full; more trials per visit are a legal schedule of the same ensemble (detailed balance holds per trial).
    # Reported next to `value`, which stays at the reference's sweep: N displacement trials + the GC draws.
    ctx.reset_counters()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(3, args.steps // 2))]
    for _ in range(2):
        ctx.run_sweeps(S, disp_per_particle=2, gc_per_cell=1)
    ctx.reset_counters()
    barrier()
    for a2, b2 in ev2:
        with torch.cuda.stream(stream):
            flush.zero_()
        a2.record(stream)
        st2 = ctx.run_sweeps(S, disp_per_particle=2, gc_per_cell=1)
        ctx.total_energy_virial(resync=True, allreduce=world > 1)
        b2.record(stream)
    barrier()
    ms2 = allmax(sum(a2.elapsed_time(b2) for a2, b2 in ev2))
    nselect2 = {"value": allsum(sum(st2["attempted"])) / (ms2 * 1e-3), "unit": "moves/s", "steps": len(ev2),
                "ms_per_step": ms2 / len(ev2),
                "what": "same step with disp_per_particle = 2 (two displacement trials per particle per cell visit)"}
    ctx.run_sweeps(S, disp_per_particle=1, gc_per_cell=1)

    # ---- roofline of the dominant kernel (k_sweep), CUDA events around every launch ----------
    ctx.set_kernel_timing(True)
    ctx.reset_counters()
    st_r = ctx.run_sweeps(S, disp_per_particle=1, gc_per_cell=1)
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
    roofline = {"bound": "hbm", "kernel": "k_sweep_p<LJ, G=%d> (one colour pass)" % info.sweep_kernel_lanes, "achieved": achieved, "peak": peak, "unit": "GB/s",
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
    trs = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")))
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
        ctx.upload(host[:n_host])  # pinned host -> device, key sort + cell build, initial energy
        h2d += 16 * n_host
        st_e = ctx.run_sweeps(S, disp_per_particle=1, gc_per_cell=1)
        U, W, N = ctx.total_energy_virial(resync=True, allreduce=world > 1)
        n_host = ctx.download_into(host)  # device -> pinned host (the caller's particle vector)
        d2h += 16 * n_host + 24 + 80  # positions + {U, W, N} + the stats block
        moves_e2e += sum(st_e["attempted"])
    barrier()
    dt = allmax(time.perf_counter() - t0)
    e2e = {"value": allsum(moves_e2e) / dt, "unit": "moves/s", "h2d_bytes_per_step": h2d // args.e2e_steps,
           "d2h_bytes_per_step": d2h // args.e2e_steps, "steps": args.e2e_steps,
           "ms_per_step": 1e3 * dt / args.e2e_steps, "slot_reallocations": ctx.info().slot_allocs - allocs0,
           "what": "mc2d_upload(host xy) + %d sweeps + energy/virial + mc2d_download(host xy) per step" % S}

    # ---- CPU baseline: the reference itself on this box's host cores (rank 0, N=1 only) -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol

        sample_n = 200_000
        wl_s = workload(1, 0, sample_n)
        if ol.have_ref():
            ref = ol.Ref()
            mcr = ref.mc(wl_s["L"], wl_s["L"], wl_s["xy"], T=T_BENCH, ptype=ol.LJ, rcut=RCUT, mu=Z_LJ_T1_RHO05,
                         delta=DELTA, ensemble=ol.GCMC, seed=42)
            secs, mv = ref.time_moves(mcr, 25, 10, True)
            es, _, _ = ref.time_energy_virial(wl_s["L"], wl_s["L"], wl_s["xy"], T=T_BENCH, ptype=ol.LJ, rcut=RCUT)
            kind = "reference"
        else:
            orc = ol.Oracle()
            mco = orc.mc(wl_s["L"], wl_s["L"], wl_s["xy"], orc.calc(T_BENCH, 0.0, ol.LJ, RCUT, Z_LJ_T1_RHO05), RCUT,
                         DELTA, ol.GCMC, 42)
            t0 = time.perf_counter()
            mco.run(25, 1 << 40, 1 << 40)
            secs, mv = time.perf_counter() - t0, 25 * (sample_n + 1)
            es, kind = None, "port"
        cpu = {"value": mv / secs, "unit": "moves/s", "cores": 1, "kind": kind,
               "sample": "25 serial sweeps (N displacement trials + 1 GC trial each) of a %d-particle cut of the same "
                         "LJ workload, move functions only (no logging / energy re-sync)" % sample_n,
               "energy_virial_particles_per_s": (sample_n / es) if es else None, "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": "attempted_mc_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "particles_start": wl["n_total"], "particles_now": int(n_now),
                       "potential": "LennardJones", "rcut": RCUT, "T": T_BENCH, "rho": RHO,
                       "activity": Z_LJ_T1_RHO05, "delta": DELTA, "ensemble": "GCMC",
                       "sweeps_per_step": S, "step": "%d GCMC sweeps + 1 energy/virial reduction" % S,
                       "sweep": "4 colour passes; per cell visit 1 displacement trial per particle + 1 GC draw "
                                "(the reference's step: N displacement trials + GC, MC.cpp:36-52)",
                       "grid": [info.nx, info.ny], "cell_capacity": info.cell_capacity,
                       "decomposition": "1-D slabs in x, %d rank(s)" % world,
                       "halo": {0: "none (single rank)", 1: "ncclSend/ncclRecv of whole slot columns",
                                2: "k_halo_sync: live slots stored into the neighbours' ghost columns over NVLink "
                                   "(CUDA IPC peer memory), flag-word sync"}[ctx.info().halo_mode],
                       "l2": "flushed between timed steps (256 MiB write on the launch stream, outside the per-step "
                             "event pairs); slot arrays are %.0f MB per buffer"
                             % (info.ncol * info.ny * info.cell_capacity * 16 / 1e6)},
            "pair_evals_per_s": pair_rate, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "nselect2": nselect2, "roofline_fp64": fp64_roof,
            "acceptance": {"displacement": st["accepted"][0] / max(st["attempted"][0], 1),
                           "insertion": st["accepted"][1] / max(st["attempted"][1], 1),
                           "deletion": st["accepted"][2] / max(st["attempted"][2], 1)},
        }
        print(json.dumps(line))
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
