#!/usr/bin/env python
"""Statistical parity of the checkerboard GPU sampler with the reference's serial sampler on the BASELINE configurations,
at the north-star gate of TWO combined standard errors.  Run on a GPU box:

    python tools/stat_parity.py [--quick] > profiles/r02_stat_parity.json

Reference side: the UNMODIFIED reference (oracle/_ref, MonteCarlo::run with its own mt19937 streams), K independent
chains on K host cores (different seeds); mean = mean of the chain means, SE = their standard deviation / sqrt(K).
GPU side: one long chain, SE by block averaging (32 blocks).  z = (gpu - ref) / sqrt(SE_gpu^2 + SE_ref^2).

Cases:
  cfg1   NVT LJ, N = 1024, L = 45.254834 (rho 0.5), rcut 2.5, T = 1, delta 0.1 (serial_src/main.cpp:102-144): <E>/N, P
  cfg2   GCMC ideal gas, V = 100, z in {0.1 .. 10} (GCMC_ideal_particles_1.cpp): <N> = zV, Var N = zV analytically
  cfg3   GCMC LJ at the reference's own sys_test_serial states: the shipped input_z_0.024 (20 x 20, T = 0.47, z = 0.0305,
         delta 0.1 added) and T = 0.45 on L = 40 (generate_file.py:3-19) at z = 0.015 and 0.020 (one-phase gas): <N>, <E>/N, P.
         T = 0.45, z = 0.024, L = 40 itself is listed as `cfg3_bistable`: that activity lies ABOVE coexistence, the stable
         phase is the liquid (N ~ 890), which the checkerboard sampler (one GC draw per cell and sweep) reaches within
         ~1e4 sweeps while the serial sampler (ONE GC trial per sweep of N displacements, MC.cpp:45-52) stays in the
         metastable gas for any affordable run, so a 2-SE comparison of means is not defined there; both branches are
         reported instead (gas branch: the serial chain and the first GPU samples before nucleation).
  t10    GCMC LJ, T = 10, L = 20, z = exp(mu / T) for mu in {-20, 0, 20}: the states of
         docs/test_results/average_simulation_data_LJ.csv (whose published <N> is listed beside ours)
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


def lattice(n, L, jitter, seed):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2).astype(float)
    return (g + 0.5) * (L / n) + rng.uniform(-jitter, jitter, g.shape)


def block_se(x, nblocks=32):
    x = np.asarray(x, float)
    nb = min(nblocks, len(x))
    m = np.array([b.mean() for b in np.array_split(x, nb)])
    return m.std(ddof=1) / math.sqrt(nb)


def ref_chain(job):
    """one serial chain of the reference in a worker process -> (mean N, mean E/N, mean P, var N)"""
    import oracle_lib as ol

    c, seed = job
    ref = ol.Ref() if ol.have_ref() else None
    ptype = getattr(ol, c["ptype"])
    ens = ol.GCMC if c["gcmc"] else ol.NVT
    if ref is not None:
        mc = ref.mc(c["L"], c["L"], c["start"], T=c["T"], ptype=ptype, rcut=c["rcut"], mu=c["z"], delta=c["delta"],
                    ensemble=ens, seed=seed)
        press = lambda xy: ref.observable(5, c["L"], c["L"], xy, T=c["T"], ptype=ptype, rcut=c["rcut"]) if len(xy) else 0.0
    else:
        orc = ol.Oracle()
        calc = orc.calc(c["T"], 0.0, ptype, c["rcut"], c["z"])
        mc = orc.mc(c["L"], c["L"], c["start"], calc, c["rcut"], c["delta"], ens, seed)
        press = lambda xy: orc.pressure_celllist(calc, c["L"], c["L"], xy) if len(xy) else 0.0
    # The reference rebuilds its cell list every `fUpdateCell` sweeps and evaluates dE on the stale list in between
    # (a particle inserted by a GC move is invisible to later trials until the next rebuild, so with hard
    # cores it accepts overlaps whose energy then explodes).  fUpdateCell = 1 is the reference at its most exact.
    stride = c["ref_stride"]
    mc.run(c["ref_equil"] // stride * stride, stride, 1)
    N, E, P = [], [], []
    for _ in range(c["ref_samples"]):
        _, n1, e1 = mc.run(stride, stride, 1)
        N.append(n1[-1])
        E.append(e1[-1])
        P.append(press(mc.particles()))
    N, E, P = np.array(N, float), np.array(E), np.array(P)
    return float(N.mean()), float((E / np.maximum(N, 1)).mean()), float(P.mean()), float(N.var())


def gpu_chain(m, c):
    V = c["L"] ** 2
    kw = dict(gc_per_cell=1) if c["gcmc"] else {}
    with m.Context(c["L"], c["L"], c["rcut"], potential=c["potential"], T=c["T"], activity=c["z"], delta=c["delta"],
                   seed=c.get("gpu_seed", 3)) as ctx:
        ctx.upload(c["start"])
        ctx.run_sweeps(c["gpu_equil"], **kw)
        N, E, P = [], [], []
        for _ in range(c["gpu_samples"]):
            st = ctx.run_sweeps(c["gpu_stride"], **kw)
            U, W, n = ctx.total_energy_virial()
            N.append(n)
            E.append(U)
            P.append(n / V * c["T"] + W / (2 * V))
        assert abs(st["energy"] - U) <= 1e-8 * max(1.0, abs(U)), (st["energy"], U)
    N, E, P = np.array(N, float), np.array(E), np.array(P)
    return N, E / np.maximum(N, 1), P


def case(name, potential, L, T, z, rcut, delta, gcmc, n0, quick, **kw):
    side = max(1, int(round(math.sqrt(n0))))
    c = dict(name=name, potential=potential, ptype={"LennardJones": "LJ", "WCA": "WCA", "Ideal": "IDEAL"}[potential], L=L,
             T=T, z=z, rcut=rcut, delta=delta, gcmc=gcmc, start=lattice(side, L, 0.25 * L / side * 0.5, 4),
             ref_equil=20000, ref_stride=40, ref_samples=4000, gpu_equil=3000, gpu_stride=4, gpu_samples=6000)
    c.update(kw)
    if quick:
        for k in ("ref_equil", "ref_samples", "gpu_equil", "gpu_samples"):
            c[k] = max(50, c[k] // 8)
    return c


def main():
    quick = "--quick" in sys.argv
    import montecarlo_simulation_b200 as m

    ncpu = os.cpu_count() or 1
    K = max(4, min(16, ncpu))
    cases = [
        case("cfg1_nvt_lj_n1024", "LennardJones", 45.254834, 1.0, 0.0, 2.5, 0.1, False, 1024, quick, ref_equil=2000,
             ref_stride=10, ref_samples=1500, gpu_equil=2000, gpu_stride=5, gpu_samples=8000),
        case("cfg3_shipped_input_T0.47_z0.0305_L20", "LennardJones", 20.0, 0.47, 0.0305, 2.5, 0.1, True, 25, quick,
             ref_equil=100000, ref_stride=100, ref_samples=6000, gpu_equil=20000, gpu_stride=10, gpu_samples=12000),
        case("cfg3_T0.45_z0.015_L40_gas", "LennardJones", 40.0, 0.45, 0.015, 2.5, 0.1, True, 25, quick,
             ref_equil=200000, ref_stride=100, ref_samples=6000, gpu_equil=20000, gpu_stride=10, gpu_samples=12000),
        case("cfg3_T0.45_z0.020_L40_gas", "LennardJones", 40.0, 0.45, 0.020, 2.5, 0.1, True, 36, quick,
             ref_equil=200000, ref_stride=100, ref_samples=6000, gpu_equil=20000, gpu_stride=10, gpu_samples=12000),
        case("cfg3_bistable_T0.45_z0.024_L40", "LennardJones", 40.0, 0.45, 0.024, 2.5, 0.1, True, 36, quick,
             ref_equil=200000, ref_stride=100, ref_samples=4000, gpu_equil=40000, gpu_stride=10, gpu_samples=12000,
             informational=True),
    ]
    published = {}
    try:
        import csv

        rows = list(csv.DictReader(open(os.path.join(ROOT, "tests", "golden", "lj_gcmc_T10_published.csv"))))
        published = {float(r["mu"]): float(r["sim_avgN"]) for r in rows}
    except Exception:
        pass
    for mu in (-20.0, 0.0, 20.0):
        cases.append(case("t10_gcmc_lj_mu%+g" % mu, "LennardJones", 20.0, 10.0, math.exp(mu / 10.0), 2.5, 0.3, True, 100, quick,
                          ref_equil=60000, ref_stride=40, ref_samples=4000, published_N=published.get(mu)))
    cases.append(case("gcmc_wca_T1_z2_L20", "WCA", 20.0, 1.0, 2.0, 2.0 ** (1.0 / 6.0), 0.2, True, 100, quick,
                      ref_equil=60000, ref_stride=40, ref_samples=4000))
    out = {"chains_reference": K, "host_cores": ncpu, "gate_se": 2.0, "cases": []}
    with mp.Pool(min(K, ncpu)) as pool:
        for c in cases:
            t0 = time.time()
            res = pool.map_async(ref_chain, [(c, 1000 + 17 * k) for k in range(K)])
            Ng, Eg, Pg = gpu_chain(m, c)
            chains = np.array(res.get())
            rec = {"case": c["name"], "potential": c["potential"], "L": c["L"], "T": c["T"], "activity": c["z"],
                   "ensemble": "GCMC" if c["gcmc"] else "NVT", "obs": {}, "seconds": None}
            for j, (nm, g) in enumerate((("N", Ng), ("E_per_N", Eg), ("P", Pg))):
                if nm == "N" and not c["gcmc"]:
                    continue
                rm, rse = chains[:, j].mean(), chains[:, j].std(ddof=1) / math.sqrt(K)
                gm, gse = float(g.mean()), block_se(g)
                se = math.hypot(rse, gse)
                rec["obs"][nm] = {"gpu": gm, "gpu_se": gse, "reference": float(rm), "reference_se": float(rse),
                                  "z": float((gm - rm) / se) if se > 0 else 0.0}
            if c.get("published_N"):
                rec["published_N_docs_csv"] = c["published_N"]
            rec["seconds"] = time.time() - t0
            rec["within_2se"] = all(abs(o["z"]) <= 2.0 for o in rec["obs"].values())
            if c.get("informational"):
                rec["informational"] = "bistable state (see the module docstring): not part of the 2-SE verdict"
            out["cases"].append(rec)
            sys.stderr.write("%s: %s\n" % (c["name"], json.dumps(rec["obs"])))
    # cfg2: ideal gas, analytic
    ideal = []
    Lx, Ly = 10 * math.sqrt(math.sqrt(3) / 2), 10 / math.sqrt(math.sqrt(3) / 2)
    for z in (0.1, 0.5, 1.0, 5.0, 10.0):
        with m.Context(Lx, Ly, 2.5, potential="Ideal", T=1.0, activity=z, delta=0.1, seed=5) as ctx:
            ctx.upload(np.zeros((0, 2)))
            ctx.run_sweeps(30000, gc_per_cell=1)
            n = np.array([ctx.run_sweeps(8, gc_per_cell=1)["n_particles"] for _ in range(4000 if not quick else 500)], float)
        V = Lx * Ly
        se = block_se(n)
        ideal.append({"z": z, "zV": z * V, "mean_N": float(n.mean()), "se": se, "z_score": float((n.mean() - z * V) / se),
                      "var_N": float(n.var()), "var_over_zV": float(n.var() / (z * V))})
    out["cfg2_ideal_gas"] = ideal
    out["all_within_2se"] = (all(c["within_2se"] for c in out["cases"] if "informational" not in c)
                             and all(abs(i["z_score"]) <= 2 for i in ideal))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
