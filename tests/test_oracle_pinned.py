"""Pins the CPU oracle (oracle/mc2d_oracle.c) before anything is allowed to trust it:
  1. the reference's own known-answer tests (unit_test_serial/test_potential.cpp, test_calc.cpp,
     test_cell_list.cpp, test_thermodynamic_celllist.cpp, test_mc_displacement.cpp), restated;
  2. committed outputs of the unmodified reference (tests/golden/ref_vectors.json);
  3. the sanity_check golden CSV shipped with the reference (tests/golden/sanity_rows.json);
  4. bit-for-bit against the reference compiled in place (oracle/_ref), when present.
No GPU needed."""
import math

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import ATHERMAL_STAR, GCMC, IDEAL, LJ, NVT, THERMAL_STAR, WCA, YUKAWA


def approx(v, eps=1.2e-5):  # Catch2's Approx default epsilon is 1.19e-5 (relative)
    return pytest.approx(v, rel=eps, abs=1e-12)


# ---- 1. reference known-answer tests -----------------------------------------------------

def test_pair_known_answers(oracle):
    """unit_test_serial/test_potential.cpp:6-43"""
    o = oracle
    assert o.pair_potential(1.0, LJ) == approx(0.0)
    assert o.pair_force(1.0, LJ) == approx(24.0)
    assert o.pair_potential(4.0, LJ) == approx(-0.061523, 1e-4)
    assert o.pair_force(4.0, LJ) == approx(-0.36328, 1e-4)
    assert o.pair_potential(1.0, WCA) == approx(1.0)
    assert o.pair_force(1.0, WCA) == approx(24.0)
    assert o.pair_potential(2.0, WCA) == 0.0 and o.pair_force(2.0, WCA) == 0.0
    assert o.pair_potential(1.0, YUKAWA) == approx(0.3678794412)
    assert o.pair_force(1.0, YUKAWA) == approx(2.0 * math.exp(-1.0))
    assert o.pair_potential(1.0, IDEAL) == 0.0 and o.pair_force(1.0, IDEAL) == 0.0


@pytest.mark.parametrize("r2,f,alpha", [(1.0, 10, 0.9), (0.255, 20, 0.85)])
def test_athermal_star_known(oracle, r2, f, alpha):
    """test_potential.cpp:45-79"""
    r = math.sqrt(r2)
    f_dep = (2 + 9 * f * f) / 24
    if r < 1:
        eu, ef = -f_dep * math.log(r) + alpha * f_dep, f_dep
    else:
        eu, ef = f_dep * alpha * math.exp((1 - r2) / (2 * alpha)), f_dep * r2 * math.exp((1 - r2) / (2 * alpha))
    assert oracle.pair_potential(r2, ATHERMAL_STAR, f_dep, 0, 0, alpha) == approx(eu)
    assert oracle.pair_force(r2, ATHERMAL_STAR, f_dep, 0, 0, alpha) == approx(ef)


@pytest.mark.parametrize("r2,f,alpha,l", [(1.2, 10, 0.9, 0.4), (0.8, 20, 0.85, 0.6)])
def test_thermal_star_known(oracle, r2, f, alpha, l):
    """test_potential.cpp:81-127"""
    r = math.sqrt(r2)
    f_dep = (2 + 9 * f * f) / 24
    A0 = f * f * l * (2.27 - 1.49 * l)
    kappa = 6.56 - 7.61 * l
    att = A0 * math.exp(-kappa * r)
    if r < 1:
        eu, ef = -f_dep * math.log(r) + alpha * f_dep - att, f_dep - att * kappa * r
    else:
        eu = f_dep * alpha * math.exp((1 - r2) / (2 * alpha)) - att
        ef = f_dep * r2 * math.exp((1 - r2) / (2 * alpha)) - att * kappa * r
    assert oracle.pair_potential(r2, THERMAL_STAR, f_dep, A0, kappa, alpha) == approx(eu)
    assert oracle.pair_force(r2, THERMAL_STAR, f_dep, A0, kappa, alpha) == approx(ef)


def test_select_potential(oracle):
    """test_potential.cpp:129-137"""
    names = ["LennardJones", "WCA", "Yukawa", "AthermalStar", "ThermalStar", "Ideal"]
    assert [oracle.select_potential(n) for n in names] == [0, 1, 2, 3, 4, 5]
    assert oracle.select_potential("Unknown") == -1


def test_calc_known_answers(oracle):
    """unit_test_serial/test_calc.cpp:16-120"""
    o = oracle
    P = np.full((5, 2), 1.0)
    c = o.calc(1.0, 0.0, IDEAL, 1.0)
    assert o.total_energy(c, 10, 10, P) == 0.0 and o.total_virial(c, 10, 10, P) == 0.0
    assert o.pressure(c, 10, 10, P) == approx(5 / 100.0)
    r0 = 2.0 ** (1 / 6) - 1e-14
    c = o.calc(1.0, 0.0, LJ, 3 * r0)
    P = np.array([[0, 0], [r0, 0]])
    assert o.total_energy(c, 6 * r0, 6 * r0, P) == approx(-1.0)
    assert o.total_virial(c, 6 * r0, 6 * r0, P) + 1 == approx(1.0)
    c = o.calc(1.0, 0.0, LJ, 5 * r0)
    P = np.array([[0, 0], [r0, 0], [r0 / 2, r0 * math.sqrt(3) / 2]])
    assert o.total_energy(c, 10 * r0, 10 * r0, P) == approx(-3.0)
    assert o.total_virial(c, 10 * r0, 10 * r0, P) + 1 == approx(1.0)
    c = o.calc(1.0, 0.0, YUKAWA, 5.0, 1.0, 1.0)
    P = np.array([[0, 0], [1.0, 0]])
    assert o.total_energy(c, 3, 3, P) == approx(math.exp(-1))
    assert o.total_virial(c, 3, 3, P) == approx(2 * math.exp(-1))
    P = np.array([[0, 0], [1.0, 0], [0.5, math.sqrt(3) / 2]])
    assert o.total_energy(c, 4, 4, P) == approx(3 * math.exp(-1))
    assert o.total_virial(c, 4, 4, P) == approx(6 * math.exp(-1))
    rc = 2.0 ** (1 / 6)
    c = o.calc(1.0, 0.0, WCA, rc)
    assert o.total_energy(c, 3 * rc, 3 * rc, np.array([[0, 0], [1.0, 0]])) == approx(1.0)
    assert o.total_energy(c, 3 * rc, 3 * rc, np.array([[0, 0], [1.5, 0]])) == 0.0


def test_calc_star_lattices(oracle):
    """test_calc.cpp:84-120 — AthermalStar 4x4 and ThermalStar 4x8 square lattices"""
    o = oracle
    P = np.array([[i, j] for i in range(4) for j in range(4)], dtype=float)
    fp, alpha = 10.0, 0.5
    c = o.calc(1.0, 0.0, ATHERMAL_STAR, 1.0001, 0.0, fp, alpha)
    Upair, Wpair, pairs = (fp * fp * 9 + 2) / 24 * alpha, (fp * fp * 9 + 2) / 24, 16 * 4 / 2
    assert o.total_energy(c, 4, 4, P) == approx(pairs * Upair)
    assert o.total_virial(c, 4, 4, P) == approx(pairs * Wpair)
    assert o.pressure(c, 4, 4, P) == approx(1.0 + pairs * Wpair / 16 / 2)
    P = np.array([[i, j] for i in range(4) for j in range(8)], dtype=float)
    f, A0, kappa, alpha = 8.0, 2.0, 1.2, 0.3
    c = o.calc(1.0, 0.0, THERMAL_STAR, 1.4, 0.0, f, alpha, A0, kappa)
    e2 = A0 * f * f * math.exp(-kappa) / kappa
    Upair = (f * f * 9 + 2) / 24 * alpha - e2
    Wpair = (f * f * 9 + 2) / 24 - A0 * f * f * math.exp(-kappa)
    pairs = 32 * 4 / 2
    assert o.total_energy(c, 4, 8, P) == approx(pairs * Upair)
    assert o.total_virial(c, 4, 8, P) == approx(pairs * Wpair)
    assert o.pressure(c, 4, 8, P) == approx(1.0 + pairs * Wpair / 32 / 2)


def test_cell_list_known(oracle):
    """unit_test_serial/test_cell_list.cpp:10-81 (periodic neighbour, inclusive cutoff, mixed cells)"""
    o = oracle
    P = np.array([[0.1, 0.1], [2.9, 2.9], [1.6, 1.5]])
    j, r2 = o.neighbors(3, 3, 0.5, P, 0)
    assert list(j) == [1] and r2[0] == pytest.approx(0.08, rel=1e-6)
    assert len(o.neighbors(3, 3, 0.5, P, 2)[0]) == 0
    P = np.array([[1.0, 1.0], [2.0, 1.0], [3.1, 1.0]])
    j, r2 = o.neighbors(5, 5, 1.0, P, 0)
    assert list(j) == [1] and r2[0] == 1.0  # exactly at r = rcut: included
    assert len(o.neighbors(5, 5, 1.0, P, 2)[0]) == 0
    P = np.array([[0.5, 0.5], [1.0, 1.0], [2.0, 0.5], [3.5, 3.5]])
    assert sorted(o.neighbors(6, 6, 1.5, P, 0)[0]) == [1, 2]
    assert len(o.neighbors(6, 6, 1.5, P, 3)[0]) == 0


@pytest.mark.parametrize("N,Lx,Ly,seed", [(100, 10.0, 10.0 * math.sqrt(3), 123456), (5000, 100.0, 10.0, 123)])
def test_celllist_equals_bruteforce(oracle, N, Lx, Ly, seed):
    """test_cell_list.cpp:85-118 and test_thermodynamic_celllist.cpp:6-28 (eps 1e-12).  Uniform
    random points; the reference draws them from std::mt19937_64, here from numpy with the same
    seed value (the property, not the sample, is what the reference asserts)."""
    xy = np.random.default_rng(seed).uniform(0, 1, (N, 2)) * [Lx, Ly]
    c = oracle.calc(1.0, 0.0, LJ, 2.5)
    assert oracle.total_energy_celllist(c, Lx, Ly, xy) == pytest.approx(oracle.total_energy(c, Lx, Ly, xy), rel=1e-12)
    assert oracle.total_virial_celllist(c, Lx, Ly, xy) == pytest.approx(oracle.total_virial(c, Lx, Ly, xy), rel=1e-12)


def test_mc_displacement_celllist_vs_bruteforce(oracle):
    """unit_test_serial/test_mc_displacement.cpp:8-70, shortened (warm-up 100 sweeps, 2000 steps)."""
    SEED, L, N, RCUT = 4922, 10.0, 100, 2.5
    c = oracle.calc(1.0, 0.0, LJ, RCUT)
    start = oracle.initialize_particles_random(L, L, N, SEED)
    warm = oracle.mc(L, L, start, c, RCUT, 0.1, NVT, SEED)
    for t in range(N * 100):
        warm.step_celllist()
        if t % 10 == 0:
            warm.update_celllist()
    p = warm.particles()
    a = oracle.mc(L, L, p, c, RCUT, 0.1, NVT, SEED)
    b = oracle.mc(L, L, p, c, RCUT, 0.1, NVT, SEED)
    assert a.energy == pytest.approx(b.energy)
    for t in range(2000):
        acc_a, acc_b = a.step_celllist(), b.step_bruteforce()
        if t % 100 == 0:
            a.update_celllist()
        assert abs(a.energy - b.energy) / max(1.0, abs(a.energy)) < 5e-4
        assert acc_a == acc_b
    np.testing.assert_allclose(a.particles(), b.particles(), rtol=1e-5)


# ---- 2. committed outputs of the unmodified reference ---------------------------------------

def test_golden_pair_vectors(oracle, golden):
    g = golden["ref"]["pair"]
    fp, fdp, kappa, alpha = g["params"]
    for t in range(6):
        for r2, (u, w) in zip(g["r2"], g["values"][str(t)]):
            assert oracle.pair_potential(r2, t, fp, fdp, kappa, alpha) == u  # bit-exact
            assert oracle.pair_force(r2, t, fp, fdp, kappa, alpha) == w


def test_golden_observables(oracle, golden):
    g = golden["ref"]
    cfg = g["config"]
    xy = np.array(cfg["xy"])
    Lx, Ly = cfg["Lx"], cfg["Ly"]
    for name, o in g["observables"].items():
        kw = o["kw"]
        c = oracle.calc(kw["T"], 0.0, o["ptype"], kw["rcut"], 0.0, kw.get("f", 0.0), kw.get("alpha", 0.0),
                        kw.get("A0", 0.0), kw.get("kappa", 0.0))
        assert oracle.total_energy(c, Lx, Ly, xy) == o["U_bf"], name
        assert oracle.total_virial(c, Lx, Ly, xy) == o["W_bf"], name
        assert oracle.pressure(c, Lx, Ly, xy) == o["P_bf"], name
        assert oracle.total_energy_celllist(c, Lx, Ly, xy) == o["U_cl"], name
        assert oracle.total_virial_celllist(c, Lx, Ly, xy) == o["W_cl"], name
        assert oracle.pressure_celllist(c, Lx, Ly, xy) == o["P_cl"], name
    assert oracle.cell_ids(Lx, Ly, cfg["rcut"], xy).tolist() == g["cell_ids"]
    assert oracle.neighbor_counts(Lx, Ly, cfg["rcut"], xy).tolist() == g["neighbor_counts"]
    c = oracle.calc(1.0, 0.0, LJ, cfg["rcut"])
    for p in g["probes"]["displace"]:
        assert oracle.local_energy(c, Lx, Ly, xy, p["idx"]) == p["U_old"]
        assert oracle.point_energy(c, Lx, Ly, xy, p["new"][0], p["new"][1], p["idx"]) == p["U_new"]
    for p in g["probes"]["insert"]:
        assert oracle.point_energy(c, Lx, Ly, xy, p["pt"][0], p["pt"][1]) == p["U"]


def test_golden_trajectories(oracle, golden):
    """MonteCarlo::run through the restated mt19937 / libstdc++ distributions: same N(t), E(t),
    positions as the reference, bit for bit."""
    g = golden["ref"]
    ini = g["init_random"]
    start = oracle.initialize_particles_random(ini["Lx"], ini["Ly"], ini["N"], ini["seed"])
    assert start[:3].tolist() == ini["head"]
    rc = 2 ** (1 / 6)
    m = oracle.mc(12.0, 12.0, start, oracle.calc(2.0, 0.0, WCA, rc), rc, 0.15, NVT, 5)
    t, n, e = m.run(40, 10, 5)
    assert t.tolist() == g["traj_nvt"]["t"] and n.tolist() == g["traj_nvt"]["N"]
    assert e.tolist() == g["traj_nvt"]["E"]
    assert m.particles()[:3].tolist() == g["traj_nvt"]["xy_head"]
    m = oracle.mc(12.0, 12.0, start, oracle.calc(2.0, 0.0, WCA, rc, 0.4), rc, 0.15, GCMC, 9)
    t, n, e = m.run(400, 50, 10)
    assert t.tolist() == g["traj_gcmc"]["t"] and n.tolist() == g["traj_gcmc"]["N"]
    assert e.tolist() == g["traj_gcmc"]["E"]
    assert m.particles()[:3].tolist() == g["traj_gcmc"]["xy_head"]


# ---- 3. the sanity_check golden CSV ------------------------------------------------------------

def test_sanity_golden_csv(oracle, golden):
    """sanity_check/sanity_parallel_data.csv: 200 frames x (U, W, P), N=24000 AthermalStar f=8
    alpha=0.3 rcut=5, box 1000x800, produced by the reference's MPI build with np=10.  The serial
    calculator must reproduce it (the reference's own check_sanity.py tolerance is 1e-6; the
    summation order differs, so 1e-12 relative here)."""
    s = golden["sanity"]
    cfg = s["config"]
    px, py = oracle.best_decomposition(cfg["Lx"], cfg["Ly"], cfg["np"])
    assert (px, py) == (s["generator_probe"]["Px"], s["generator_probe"]["Py"]) == (5, 2)
    c = oracle.calc(cfg["T"], 0.0, ATHERMAL_STAR, cfg["rcut"], 0.0, cfg["f"], cfg["alpha"], cfg["A0"], cfg["kappa"])
    for frame in (0, 1, 199):
        xy = oracle.sanity_frame(cfg["Lx"], cfg["Ly"], px, py, cfg["N"], cfg["seed0"] + frame)
        probe = s["generator_probe"]["frames"].get(str(frame))
        if probe:  # the generator itself is pinned against the reference's initial_mpi.cpp output
            assert xy[:4].tolist() == probe["head"] and xy[-4:].tolist() == probe["tail"]
        row = s["rows"][str(frame)]
        assert oracle.total_energy_celllist(c, cfg["Lx"], cfg["Ly"], xy) == pytest.approx(row["U"], rel=1e-12)
        assert oracle.total_virial_celllist(c, cfg["Lx"], cfg["Ly"], xy) == pytest.approx(row["W"], rel=1e-12)
        assert oracle.pressure_celllist(c, cfg["Lx"], cfg["Ly"], xy) == pytest.approx(row["P"], rel=1e-12)


# ---- 4. bit-for-bit against the reference compiled in place ---------------------------------------

def _random_cfg(seed, n_side, Lx, Ly, jitter=0.3):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij"), -1).reshape(-1, 2).astype(float)
    xy = (g + 0.5) * np.array([Lx / n_side, Ly / n_side]) + rng.uniform(-jitter, jitter, size=g.shape)
    xy[:, 0] %= Lx
    xy[:, 1] %= Ly
    return xy


POTS = [(LJ, {}), (WCA, {"rcut": 2 ** (1 / 6)}), (YUKAWA, {}), (ATHERMAL_STAR, {"f": 8.0, "alpha": 0.3}),
        (THERMAL_STAR, {"f": 6.0, "alpha": 0.4, "A0": 1.5, "kappa": 1.1}), (IDEAL, {})]


@pytest.mark.parametrize("ptype,kw", POTS)
def test_ref_bitwise_observables(oracle, ref, ptype, kw):
    Lx, Ly = 31.0, 26.0
    xy = _random_cfg(7 + ptype, 24, Lx, Ly)
    k = dict(T=0.9, ptype=ptype, rcut=2.5)
    k.update(kw)
    c = oracle.calc(k["T"], 0.0, ptype, k["rcut"], 0.0, k.get("f", 0.0), k.get("alpha", 0.0), k.get("A0", 0.0),
                    k.get("kappa", 0.0))
    fns = [oracle.total_energy, oracle.total_virial, oracle.pressure, oracle.total_energy_celllist,
           oracle.total_virial_celllist, oracle.pressure_celllist]
    for what, fn in enumerate(fns):
        assert fn(c, Lx, Ly, xy) == ref.observable(what, Lx, Ly, xy, **k)
    for i in (0, 100, 575):
        assert oracle.local_energy(c, Lx, Ly, xy, i) == ref.local_energy(
            Lx, Ly, xy, i, **{a: b for a, b in k.items()})
        x, y = xy[i] + [0.05, -0.08]
        assert oracle.point_energy(c, Lx, Ly, xy, x, y, i) == ref.point_energy(Lx, Ly, xy, x, y, i, **k)


@pytest.mark.parametrize("Lx,Ly,rcut", [(31.0, 26.0, 2.5), (10.0, 10.0, 2.5), (9.306, 10.746, 2.5), (7.3, 4.9, 2.5),
                                         (3.0, 3.0, 5.0), (45.254834, 45.254834, 2.5)])
def test_ref_bitwise_cells_and_neighbors(oracle, ref, Lx, Ly, rcut):
    """cell geometry, cell ids, neighbour lists (order and r^2) incl. degenerate grids nx<3 where
    the reference's 3x3 scan revisits cells."""
    assert oracle.cell_dims(Lx, Ly, rcut) == ref.cell_dims(Lx, Ly, rcut)
    rng = np.random.default_rng(11)
    xy = rng.uniform(0, 1, (300, 2)) * [Lx, Ly]
    xy[0] = [0.0, 0.0]
    xy[1] = [np.nextafter(Lx, 0), np.nextafter(Ly, 0)]
    assert (oracle.cell_ids(Lx, Ly, rcut, xy) == ref.cell_ids(Lx, Ly, rcut, xy)).all()
    assert (oracle.neighbor_counts(Lx, Ly, rcut, xy) == ref.neighbor_counts(Lx, Ly, rcut, xy)).all()
    for i in (0, 1, 17, 299):
        jo, ro = oracle.neighbors(Lx, Ly, rcut, xy, i)
        jr, rr = ref.neighbors(Lx, Ly, rcut, xy, i)
        assert jo.tolist() == jr.tolist() and ro.tolist() == rr.tolist()
    jo, ro = oracle.neighbors_point(Lx, Ly, rcut, xy, 0.3 * Lx, 0.9 * Ly)
    jr, rr = ref.neighbors_point(Lx, Ly, rcut, xy, 0.3 * Lx, 0.9 * Ly)
    assert jo.tolist() == jr.tolist() and ro.tolist() == rr.tolist()


def test_ref_bitwise_rng(oracle, ref):
    o = oracle.rng(12345)
    r = ref.lib.ref_rng_create(12345)
    import ctypes as C

    r = C.c_void_p(r)
    for k in range(3000):
        if k % 3 == 0:
            assert oracle.lib.orc_rng_uniform01(o) == ref.lib.ref_rng_uniform01(r)
        elif k % 3 == 1:
            assert oracle.lib.orc_rng_uniform(o, -0.1, 0.1) == ref.lib.ref_rng_uniform(r, -0.1, 0.1)
        else:
            hi = 1 + (k * 7919) % 100000
            assert oracle.lib.orc_rng_uniform_int(o, 0, hi) == ref.lib.ref_rng_uniform_int(r, 0, hi)
    ref.lib.ref_rng_destroy(r)


@pytest.mark.parametrize("ens,ptype,mu", [(NVT, LJ, 0.0), (GCMC, LJ, 0.05), (GCMC, IDEAL, 0.5)])
def test_ref_bitwise_trajectory(oracle, ref, ens, ptype, mu):
    L, N, rcut = 14.0, 70, 2.5
    start = ref.initialize_particles_random(L, L, N, 31)
    assert (start == oracle.initialize_particles_random(L, L, N, 31)).all()
    # relax overlaps first with a soft potential so that LJ energies are finite and meaningful
    rc = 2 ** (1 / 6)
    pre = oracle.mc(L, L, start, oracle.calc(1.0, 0.0, WCA, rc), rc, 0.2, NVT, 3)
    pre.run(200, 200, 10)
    start = pre.particles()
    a = oracle.mc(L, L, start, oracle.calc(1.1, 0.0, ptype, rcut, mu), rcut, 0.12, ens, 17)
    b = ref.mc(L, L, start, T=1.1, ptype=ptype, rcut=rcut, mu=mu, delta=0.12, ensemble=ens, seed=17)
    ta, na, ea = a.run(300, 20, 10)
    tb, nb, eb = b.run(300, 20, 10)
    assert ta.tolist() == tb.tolist() and na.tolist() == nb.tolist()
    assert ea.tolist() == eb.tolist()
    assert (a.particles() == b.particles()).all()


# ---- 5. ensemble fact the reference documents: ideal-gas GCMC <N> = zV ---------------------------

def test_ideal_gas_gcmc_mean(oracle):
    """docs/test_scenarios.md:13-56 / integration_test_serial/GCMC_ideal_particles_1.cpp: V=100,
    z=0.5 -> <N> = 50, Var N = 50 (short run: 4 sigma of the block-averaged error)."""
    Lx, Ly = 10 * math.sqrt(math.sqrt(3) / 2), 10 / math.sqrt(math.sqrt(3) / 2)
    start = np.random.default_rng(5).uniform(0, 1, (50, 2)) * [Lx, Ly]
    m = oracle.mc(Lx, Ly, start, oracle.calc(1.0, 0.0, IDEAL, 2.5, 0.5), 2.5, 0.1, GCMC, 99)
    m.run(2000, 2000, 10)
    _, n, _ = m.run(40000, 10, 10)
    blocks = n.reshape(20, -1).mean(axis=1)
    se = blocks.std(ddof=1) / math.sqrt(len(blocks))
    assert abs(n.mean() - 50.0) < 4 * se + 0.5
