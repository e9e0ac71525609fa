#!/usr/bin/env python
"""Statistical parity of the checkerboard GPU sampler with the reference's serial sampler on the BASELINE configurations,
at the north-star gate of TWO combined standard errors.  Two phases, so that the long CPU chains do not hold a GPU box:

    python tools/stat_parity.py gpu  [--quick]     # on a GPU box  -> gpurun_out/stat_gpu.npz
    python tools/stat_parity.py ref  [--quick]     # anywhere (CPU) -> profiles/r02_stat_parity.json + .md

Reference side: the UNMODIFIED reference (oracle/_ref), K independent chains on the host cores (different mt19937
seeds); mean = mean of the chain means, SE = their standard deviation / sqrt(K).  GPU side: one long chain, SE by block
averaging (32 blocks).  z = (gpu - ref) / sqrt(SE_gpu^2 + SE_ref^2).

How the reference is driven (ref_mode):
  run    MonteCarlo::run with a cell rebuild EVERY sweep (fUpdateCell = 1): the stock loop at its most exact.  Between
         rebuilds the reference evaluates dE on a stale list: a particle that crossed a cell wall, or one
         a GC move has just inserted, can be missed.  Harmless for soft tails, not for hard cores.
  exact  the same moves with the cell list rebuilt before EVERY trial (oracle/ref_shim.cpp: ref_mc_run_exact; O(N) per
         trial).  Used where `run` lets overlaps through (WCA) and for the small cases.
Dense states start the reference chains from decorrelated snapshots of the equilibrated GPU chain: the serial sampler
makes ONE insertion/deletion trial per sweep of N displacements (MC.cpp:45-52), so condensing a liquid from the gas
takes it millions of sweeps; where it starts does not change what it samples.

Cases:
  cfg1   NVT LJ, N = 1024, L = 45.254834 (rho 0.5), rcut 2.5, T = 1, delta 0.1 (serial_src/main.cpp:102-144): <E>/N, P
  cfg2   GCMC ideal gas, V = 100, z in {0.1 .. 10} (GCMC_ideal_particles_1.cpp): <N> = zV, Var N = zV analytically
  cfg3   GCMC LJ at the reference's sys_test_serial states: the shipped input_z_0.024 (20 x 20, T = 0.47, z = 0.0305, delta
         0.1 added; a liquid, N ~ 280) and T = 0.45 on L = 40 (generate_file.py:3-19) at z = 0.015, 0.020 (gas) and 0.024
         (above coexistence: the stable phase is the liquid, N ~ 1150): <N>, <E>/N, P
  t10    GCMC LJ, T = 10, L = 20, z = exp(mu / T) for mu in {-20, 0, 20}: the state points of
         docs/test_results/average_simulation_data_LJ.csv (an older code version with an undocumented box; its <N> is
         listed for reference only)
  wca    GCMC WCA, T = 1, z = 2, L = 20
The pressure is rho T + W / (2 V) (thermodynamic_calculator.cpp:235-242) on both sides."""
import json
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GPU_FILE = os.path.join(ROOT, "gpurun_out", "stat_gpu.npz")


def lattice(n, L, jitter, seed):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2).astype(float)
    return (g + 0.5) * (L / n) + rng.uniform(-jitter, jitter, g.shape)


def block_se(x, nblocks=32):
    x = np.asarray(x, float)
    nb = min(nblocks, len(x))
    m = np.array([b.mean() for b in np.array_split(x, nb)])
    return float(m.std(ddof=1) / math.sqrt(nb))


def make_cases(quick):
    def case(name, potential, L, T, z, rcut, delta, gcmc, n0, **kw):
        side = max(1, int(round(math.sqrt(n0))))
        c = dict(name=name, potential=potential, L=L, T=T, z=z, rcut=rcut, delta=delta, gcmc=gcmc, n0=side * side,
                 ref_mode="exact", ref_equil=20000, ref_stride=40, ref_samples=4000, from_gpu=False,
                 gpu_equil=3000, gpu_stride=4, gpu_samples=8000)
        c.update(kw)
        if quick:
            for k in ("ref_equil", "ref_samples", "gpu_equil", "gpu_samples"):
                c[k] = max(50, c[k] // 10)
        return c

    cases = [
        case("cfg1_nvt_lj_n1024", "LennardJones", 45.254834, 1.0, 0.0, 2.5, 0.1, False, 1024, ref_mode="run",
             ref_equil=3000, ref_stride=10, ref_samples=4000, gpu_equil=3000, gpu_stride=5, gpu_samples=30000),
        case("cfg3_shipped_input_T0.47_z0.0305_L20", "LennardJones", 20.0, 0.47, 0.0305, 2.5, 0.1, True, 25, from_gpu=True,
             ref_mode="run", ref_equil=60000, ref_stride=100, ref_samples=3000, gpu_equil=30000, gpu_stride=10,
             gpu_samples=20000),
        case("cfg3_T0.45_z0.015_L40_gas", "LennardJones", 40.0, 0.45, 0.015, 2.5, 0.1, True, 25,
             ref_equil=200000, ref_stride=100, ref_samples=6000, gpu_equil=20000, gpu_stride=10, gpu_samples=12000),
        case("cfg3_T0.45_z0.020_L40_gas", "LennardJones", 40.0, 0.45, 0.020, 2.5, 0.1, True, 36,
             ref_equil=200000, ref_stride=100, ref_samples=6000, gpu_equil=20000, gpu_stride=10, gpu_samples=12000),
        case("cfg3_T0.45_z0.024_L40_liquid", "LennardJones", 40.0, 0.45, 0.024, 2.5, 0.1, True, 36, from_gpu=True,
             ref_mode="run", ref_equil=40000, ref_stride=200, ref_samples=1000, gpu_equil=60000, gpu_stride=10,
             gpu_samples=20000),
    ]
    for mu in (-20.0, 0.0, 20.0):
        cases.append(case("t10_gcmc_lj_mu%+g" % mu, "LennardJones", 20.0, 10.0, math.exp(mu / 10.0), 2.5, 0.3, True, 100,
                          ref_equil=60000, ref_stride=40, ref_samples=4000))
    cases.append(case("gcmc_wca_T1_z2_L20", "WCA", 20.0, 1.0, 2.0, 2.0 ** (1.0 / 6.0), 0.2, True, 100,
                      ref_equil=60000, ref_stride=40, ref_samples=4000))
    return cases


# ---- phase 1: the GPU chains -----------------------------------------------------------------------------------------------
def phase_gpu(quick):
    import montecarlo_simulation_b200 as m

    out = {}
    for c in make_cases(quick):
        t0 = time.time()
        V = c["L"] ** 2
        kw = dict(gc_per_cell=1) if c["gcmc"] else {}
        side = int(round(math.sqrt(c["n0"])))
        start = lattice(side, c["L"], 0.125 * c["L"] / side, 4)
        with m.Context(c["L"], c["L"], c["rcut"], potential=c["potential"], T=c["T"], activity=c["z"], delta=c["delta"],
                       seed=3) as ctx:
            ctx.upload(start)
            ctx.run_sweeps(c["gpu_equil"], **kw)
            N, E, P, snaps = [], [], [], []
            every = max(1, c["gpu_samples"] // 16)
            for k in range(c["gpu_samples"]):
                st = ctx.run_sweeps(c["gpu_stride"], **kw)
                U, W, n = ctx.total_energy_virial()
                N.append(n)
                E.append(U)
                P.append(n / V * c["T"] + W / (2 * V))
                if c["from_gpu"] and k % every == 0 and len(snaps) < 16:
                    snaps.append(ctx.download().copy())
            assert abs(st["energy"] - U) <= 1e-8 * max(1.0, abs(U)), (st["energy"], U)
        out[c["name"] + "/N"], out[c["name"] + "/E"], out[c["name"] + "/P"] = np.array(N, float), np.array(E), np.array(P)
        for k, sn in enumerate(snaps):
            out["%s/snap%d" % (c["name"], k)] = sn
        sys.stderr.write("%s: gpu chain %.0f s, <N> = %.2f\n" % (c["name"], time.time() - t0, np.mean(N)))
    # cfg2: ideal gas, analytic
    Lx, Ly = 10 * math.sqrt(math.sqrt(3) / 2), 10 / math.sqrt(math.sqrt(3) / 2)
    for z in (0.1, 0.5, 1.0, 5.0, 10.0):
        with m.Context(Lx, Ly, 2.5, potential="Ideal", T=1.0, activity=z, delta=0.1, seed=5) as ctx:
            ctx.upload(np.zeros((0, 2)))
            ctx.run_sweeps(30000 if not quick else 3000, gc_per_cell=1)
            out["ideal/%g" % z] = np.array([ctx.run_sweeps(8, gc_per_cell=1)["n_particles"]
                                            for _ in range(4000 if not quick else 400)], float)
    os.makedirs(os.path.dirname(GPU_FILE), exist_ok=True)
    np.savez_compressed(GPU_FILE, **out)
    print("wrote", GPU_FILE)


# ---- phase 2: the reference chains (CPU) and the verdict ------------------------------------------------------------------------
def ref_chain(job):
    """one serial chain of the reference in a worker process -> (mean N, mean E/N, mean P)"""
    import oracle_lib as ol

    c, seed, start = job
    ref = ol.Ref()
    ptype = {"LennardJones": ol.LJ, "WCA": ol.WCA}[c["potential"]]
    ens = ol.GCMC if c["gcmc"] else ol.NVT
    mc = ref.mc(c["L"], c["L"], start, T=c["T"], ptype=ptype, rcut=c["rcut"], mu=c["z"], delta=c["delta"], ensemble=ens,
                seed=seed)
    press = lambda xy: ref.observable(5, c["L"], c["L"], xy, T=c["T"], ptype=ptype, rcut=c["rcut"]) if len(xy) else 0.0
    total_u = lambda xy: ref.observable(3, c["L"], c["L"], xy, T=c["T"], ptype=ptype, rcut=c["rcut"]) if len(xy) else 0.0
    stride = c["ref_stride"]

    def advance(nsweeps):
        if c["ref_mode"] == "exact":
            # (the class's running energy is only re-synchronised inside MonteCarlo::run, MC.cpp:60-70, and drifts
            # without it: the sample is the energy recomputed from the positions, computeTotalEnergyCellList)
            n, _ = ref.run_exact(mc, nsweeps, c["gcmc"])
            return n, total_u(mc.particles())
        _, n1, e1 = mc.run(nsweeps, nsweeps, 1)
        return int(n1[-1]), float(e1[-1])

    advance(max(stride, c["ref_equil"] // stride * stride))
    N, E, P = [], [], []
    for _ in range(c["ref_samples"]):
        n, e = advance(stride)
        N.append(n)
        E.append(e)
        P.append(press(mc.particles()))
    N, E, P = np.array(N, float), np.array(E), np.array(P)
    return float(N.mean()), float((E / np.maximum(N, 1)).mean()), float(P.mean())


def phase_ref(quick):
    g = np.load(GPU_FILE)
    ncpu = os.cpu_count() or 1
    K = max(4, min(16, ncpu))
    published = {}
    try:
        import csv

        rows = list(csv.DictReader(open(os.path.join(ROOT, "tests", "golden", "lj_gcmc_T10_published.csv"))))
        published = {float(r["mu"]): float(r["sim_avgN"]) for r in rows}
    except Exception:
        pass
    out = {"chains_reference": K, "host_cores": ncpu, "gate_se": 2.0, "cases": []}
    keep = {}  # `ref --redo-exact`: the records of the stock-loop cases are taken from the last report
    if "--redo-exact" in sys.argv:
        old = json.load(open(os.path.join(ROOT, "profiles", "r02_stat_parity" + ("_quick" if quick else "") + ".json")))
        keep = {r["case"]: r for r in old["cases"] if r["reference_mode"] != "exact"}
    with mp.Pool(min(K, ncpu)) as pool:
        for c in make_cases(quick):
            if c["name"] in keep:
                out["cases"].append(keep[c["name"]])
                continue
            t0 = time.time()
            side = int(round(math.sqrt(c["n0"])))
            jobs = []
            for k in range(K):
                start = g["%s/snap%d" % (c["name"], k % 16)] if c["from_gpu"] else lattice(side, c["L"], 0.125 * c["L"] / side, 4)
                jobs.append((c, 1000 + 17 * k, start))
            chains = np.array(pool.map(ref_chain, jobs))
            Ng, Eg, Pg = g[c["name"] + "/N"], g[c["name"] + "/E"], g[c["name"] + "/P"]
            rec = {"case": c["name"], "potential": c["potential"], "L": c["L"], "T": c["T"], "activity": c["z"],
                   "ensemble": "GCMC" if c["gcmc"] else "NVT", "reference_mode": c["ref_mode"],
                   "reference_start": "snapshots of the equilibrated GPU chain" if c["from_gpu"] else "jittered lattice",
                   "obs": {}}
            for j, (nm, gv) in enumerate((("N", Ng), ("E_per_N", Eg / np.maximum(Ng, 1)), ("P", Pg))):
                if nm == "N" and not c["gcmc"]:
                    continue
                rm, rse = chains[:, j].mean(), chains[:, j].std(ddof=1) / math.sqrt(K)
                gm, gse = float(gv.mean()), block_se(gv)
                se = math.hypot(rse, gse)
                rec["obs"][nm] = {"gpu": gm, "gpu_se": gse, "reference": float(rm), "reference_se": float(rse),
                                  "z": float((gm - rm) / se) if se > 0 else 0.0}
            if c["name"].startswith("t10"):
                mu = 10.0 * math.log(c["z"])
                rec["published_N_docs_csv_older_code"] = published.get(round(mu, 6))
            rec["seconds_reference"] = time.time() - t0
            rec["within_2se"] = all(abs(o["z"]) <= 2.0 for o in rec["obs"].values())
            out["cases"].append(rec)
            sys.stderr.write("%s: %s\n" % (c["name"], json.dumps(rec["obs"])))
    ideal = []
    V = 100.0
    for z in (0.1, 0.5, 1.0, 5.0, 10.0):
        n = g["ideal/%g" % z]
        se = block_se(n)
        ideal.append({"z": z, "zV": z * V, "mean_N": float(n.mean()), "se": se, "z_score": float((n.mean() - z * V) / se),
                      "var_N": float(n.var()), "var_over_zV": float(n.var() / (z * V))})
    out["cfg2_ideal_gas"] = ideal
    zs = [o["z"] for c in out["cases"] for o in c["obs"].values()] + [i["z_score"] for i in ideal]
    out["z_scores"] = {"count": len(zs), "max_abs": max(abs(v) for v in zs), "beyond_2": sum(abs(v) > 2 for v in zs),
                       "expected_beyond_2_if_unbiased": 0.0455 * len(zs), "rms": float(np.sqrt(np.mean(np.square(zs))))}
    out["all_within_2se"] = all(abs(v) <= 2 for v in zs)
    tag = "r02_stat_parity" + ("_quick" if quick else "")
    json.dump(out, open(os.path.join(ROOT, "profiles", tag + ".json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", tag + ".md"), "w") as fh:
        fh.write("# Statistical parity with the reference's serial sampler (tools/stat_parity.py)\n\n"
                 "GPU: one chain per case on a B200 (block-averaged SE).  Reference: the unmodified `serial_src` "
                 "(oracle/_ref), %d independent chains on %d host cores.  Gate: |z| <= 2 (two combined standard errors).\n\n"
                 "| case | observable | GPU | reference | z |\n|---|---|---:|---:|---:|\n" % (K, ncpu))
        for c in out["cases"]:
            for nm, o in c["obs"].items():
                fh.write("| %s (%s) | %s | %.6g ± %.2g | %.6g ± %.2g | %+.2f |\n" %
                         (c["case"], c["reference_mode"], nm, o["gpu"], o["gpu_se"], o["reference"], o["reference_se"], o["z"]))
        fh.write("\nIdeal gas (analytic, <N> = zV, Var N = zV):\n\n| z | zV | <N> | z-score | Var N / zV |\n|---|---:|---:|---:|---:|\n")
        for i in ideal:
            fh.write("| %g | %.0f | %.2f ± %.2f | %+.2f | %.3f |\n" % (i["z"], i["zV"], i["mean_N"], i["se"], i["z_score"], i["var_over_zV"]))
        fh.write("\n%d z-scores: max |z| = %.2f, %d beyond 2 (%.1f expected for an unbiased sampler), rms %.2f.\n" %
                 (out["z_scores"]["count"], out["z_scores"]["max_abs"], out["z_scores"]["beyond_2"],
                  out["z_scores"]["expected_beyond_2_if_unbiased"], out["z_scores"]["rms"]))
    print(json.dumps(out["z_scores"]), "all_within_2se", out["all_within_2se"])


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    if len(sys.argv) > 1 and sys.argv[1] == "gpu":
        phase_gpu(quick)
    elif len(sys.argv) > 1 and sys.argv[1] == "ref":
        phase_ref(quick)
    else:
        sys.exit(__doc__)
