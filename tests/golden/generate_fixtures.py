#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the reference tree.  Run in the BUILD container only
(`python tests/golden/generate_fixtures.py`); /root/reference does not exist on the GPU box, so
tests read the committed JSON, never the reference.

Fixtures written:
  sanity_rows.json      rows of sanity_check/sanity_parallel_data.csv (U, W, P of 200 random
                        N=24000 AthermalStar frames, np=10 MPI run shipped with the reference) and
                        the first/last particles of frames 0 and 1 as produced by the reference's
                        own mpi_src/initial_mpi.cpp + rng_parallel.h compiled here (MPI-free files).
  ref_vectors.json      outputs of the UNMODIFIED serial reference (oracle/_ref) on seeded inputs:
                        pair functions, U/W/P, cell ids, neighbour counts, per-trial dE probes,
                        short NVT / GCMC trajectories.  Lets the GPU box (no /root/reference, but
                        also usable without oracle/_ref) pin the C oracle.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MC2D_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))

SANITY_CPP = r"""
#include <cstdio>
#include <vector>
#include "particle.h"
#include "simulation_box.h"
#include "initial_mpi.h"
int main() {
    SimulationBox box(1000.0, 800.0);
    auto d = box.bestDecomposition(10);
    std::printf("{\"Px\": %d, \"Py\": %d, \"frames\": {", d.Px, d.Py);
    for (int frame = 0; frame < 2; ++frame) {
        std::vector<Particle> all;
        for (int r = 0; r < 10; ++r) {
            std::vector<Particle> o;
            initializeParticles_globalN(o, box, d, r, 24000, true, 12345u + frame, 1000000LL);
            all.insert(all.end(), o.begin(), o.end());
        }
        std::printf("%s\"%d\": {\"n\": %zu, \"head\": [", frame ? ", " : "", frame, all.size());
        for (int i = 0; i < 4; ++i) std::printf("%s[%.17g, %.17g]", i ? ", " : "", all[i].x, all[i].y);
        std::printf("], \"tail\": [");
        for (int i = 0; i < 4; ++i) {
            auto &p = all[all.size() - 4 + i];
            std::printf("%s[%.17g, %.17g]", i ? ", " : "", p.x, p.y);
        }
        double sx = 0, sy = 0;
        for (auto &p : all) { sx += p.x; sy += p.y; }
        std::printf("], \"sum_x\": %.17g, \"sum_y\": %.17g}", sx, sy);
    }
    std::printf("}}\n");
    return 0;
}
"""


def sanity_rows():
    rows = {}
    with open(os.path.join(REF, "sanity_check", "sanity_parallel_data.csv")) as fh:
        header = fh.readline().strip().split(",")
        for line in fh:
            f = line.strip().split(",")
            frame = int(f[0])
            if frame in (0, 1, 2, 3, 100, 199):
                rows[str(frame)] = {"N": int(f[1]), "rho": float(f[2]), "U": float(f[3]), "W": float(f[4]),
                                    "P": float(f[5])}
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "gen.cpp")
        open(src, "w").write(SANITY_CPP)
        exe = os.path.join(td, "gen")
        m = os.path.join(REF, "mpi_src")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + m, src, os.path.join(m, "initial_mpi.cpp"),
                               os.path.join(m, "simulatin_box.cpp"), os.path.join(m, "particle.cpp"), "-o", exe])
        gen = json.loads(subprocess.check_output([exe]).decode())
    out = {
        "source": "sanity_check/sanity_parallel_data.csv (header: %s); inputs: sanity_check/main_parallel.cpp:31-80"
                  % ",".join(header),
        "config": {"Lx": 1000.0, "Ly": 800.0, "N": 24000, "seed0": 12345, "np": 10, "potential": "AthermalStar",
                   "T": 1.0, "rcut": 5.0, "f": 8.0, "alpha": 0.3, "A0": 0.0, "kappa": 0.0},
        "rows": rows,
        "generator_probe": gen,
    }
    json.dump(out, open(os.path.join(HERE, "sanity_rows.json"), "w"), indent=1)


def ref_vectors():
    import oracle_lib as ol

    ref = ol.Ref()
    out = {"note": "outputs of the unmodified reference serial_src built by oracle/Makefile (g++ -std=c++17 -O3)"}
    # pair functions over a grid of r2 for the six potentials (float-narrowed star parameters)
    r2s = [0.255, 0.5, 0.8, 0.95, 1.0, 1.2, 1.2598, 1.2599, 1.26, 1.5, 2.0, 2.5, 4.0, 6.25]
    fp, fdp, kappa, alpha = (2 + 9 * 64) / 24.0, 2.0 * 64 / 1.2, 1.2, 0.3
    pf = {}
    for t in range(6):
        pf[str(t)] = [[ref.lib.ref_pair_potential(r, t, fp, fdp, kappa, alpha),
                       ref.lib.ref_pair_force(r, t, fp, fdp, kappa, alpha)] for r in r2s]
    out["pair"] = {"r2": r2s, "params": [fp, fdp, kappa, alpha], "values": pf}

    # seeded jittered-lattice configuration (min separation keeps LJ finite)
    rng = np.random.default_rng(20251019)
    Lx, Ly, rcut = 30.0, 27.5, 2.5
    g = np.stack(np.meshgrid(np.arange(20), np.arange(18), indexing="ij"), -1).reshape(-1, 2).astype(float)
    xy = (g + 0.5) * np.array([Lx / 20, Ly / 18]) + rng.uniform(-0.3, 0.3, size=g.shape)
    xy[:, 0] %= Lx
    xy[:, 1] %= Ly
    out["config"] = {"Lx": Lx, "Ly": Ly, "rcut": rcut, "xy": xy.tolist()}
    obs = {}
    for name, t, kw in (("LJ", ol.LJ, {}), ("WCA", ol.WCA, {"rcut": 2 ** (1 / 6)}), ("Yukawa", ol.YUKAWA, {}),
                        ("AthermalStar", ol.ATHERMAL_STAR, {"f": 8.0, "alpha": 0.3}),
                        ("ThermalStar", ol.THERMAL_STAR, {"f": 8.0, "alpha": 0.3, "A0": 2.0, "kappa": 1.2}),
                        ("Ideal", ol.IDEAL, {})):
        k = dict(T=1.3, ptype=t, rcut=rcut)
        k.update(kw)
        obs[name] = {"kw": {a: b for a, b in k.items() if a != "ptype"}, "ptype": t,
                     "U_bf": ref.observable(0, Lx, Ly, xy, **k), "W_bf": ref.observable(1, Lx, Ly, xy, **k),
                     "P_bf": ref.observable(2, Lx, Ly, xy, **k), "U_cl": ref.observable(3, Lx, Ly, xy, **k),
                     "W_cl": ref.observable(4, Lx, Ly, xy, **k), "P_cl": ref.observable(5, Lx, Ly, xy, **k)}
    out["observables"] = obs
    out["cell_ids"] = ref.cell_ids(Lx, Ly, rcut, xy).tolist()
    out["neighbor_counts"] = ref.neighbor_counts(Lx, Ly, rcut, xy).tolist()
    probes = []
    for i in (0, 7, 19, 180, 359):
        nx, ny = xy[i, 0] + 0.07, xy[i, 1] - 0.05
        probes.append({"idx": i, "new": [nx, ny], "U_old": ref.local_energy(Lx, Ly, xy, i),
                       "U_new": ref.point_energy(Lx, Ly, xy, nx, ny, exclude=i)})
    ins = []
    for k in range(6):
        px, py = rng.uniform(0, Lx), rng.uniform(0, Ly)
        ins.append({"pt": [px, py], "U": ref.point_energy(Lx, Ly, xy, px, py)})
    out["probes"] = {"displace": probes, "insert": ins}

    # short trajectories through MonteCarlo::run
    start = ref.initialize_particles_random(12.0, 12.0, 60, 77)
    out["init_random"] = {"Lx": 12.0, "Ly": 12.0, "N": 60, "seed": 77, "head": start[:3].tolist()}
    m = ref.mc(12.0, 12.0, start, T=2.0, ptype=ol.WCA, rcut=2 ** (1 / 6), delta=0.15, ensemble=ol.NVT, seed=5)
    t, n, e = m.run(40, 10, 5)
    out["traj_nvt"] = {"t": t.tolist(), "N": n.tolist(), "E": e.tolist(), "xy_head": m.particles()[:3].tolist()}
    m = ref.mc(12.0, 12.0, start, T=2.0, ptype=ol.WCA, rcut=2 ** (1 / 6), mu=0.4, delta=0.15, ensemble=ol.GCMC,
               seed=9)
    t, n, e = m.run(400, 50, 10)
    out["traj_gcmc"] = {"t": t.tolist(), "N": n.tolist(), "E": e.tolist(), "xy_head": m.particles()[:3].tolist()}
    json.dump(out, open(os.path.join(HERE, "ref_vectors.json"), "w"))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s" % REF)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)
    sanity_rows()
    ref_vectors()
    print("fixtures written to", HERE)
