"""GPU tests of the reference-interface layer: the C++ shims (CellList, ThermodynamicCalculator,
MonteCarlo with the reference's signatures) and the input.txt driver, plus the neighbour-list and
brute-force hooks of the C ABI against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "montecarlo_simulation_b200", "host")


@pytest.fixture(scope="module")
def mc():
    import montecarlo_simulation_b200 as m
    from montecarlo_simulation_b200 import api

    assert api.device_count() > 0
    return m


@pytest.fixture(scope="module")
def host_built():
    subprocess.check_call(["make", "-s", "-C", HOST, "all"])
    return HOST


def test_neighbor_lists_match_reference_order(mc, oracle):
    """CellList::getNeighbors / getNeighbors2: same (j, r^2) pairs in the same order, bit-exact."""
    for Lx, Ly, rcut in [(31.0, 26.0, 2.5), (7.3, 4.9, 2.5), (10.0, 10.0, 2.5)]:
        xy = np.random.default_rng(2).uniform(0, 1, (400, 2)) * [Lx, Ly]
        with mc.Context(Lx, Ly, rcut) as ctx:
            for i in (0, 17, 399):
                j, r2 = ctx.calc_neighbor_list(xy, index=i)
                jo, ro = oracle.neighbors(Lx, Ly, rcut, xy, i)
                assert j.tolist() == jo.tolist() and r2.tolist() == ro.tolist()
            j, r2 = ctx.calc_neighbor_list(xy, point=(0.3 * Lx, 0.9 * Ly))
            jo, ro = oracle.neighbors_point(Lx, Ly, rcut, xy, 0.3 * Lx, 0.9 * Ly)
            assert j.tolist() == jo.tolist() and r2.tolist() == ro.tolist()


@pytest.mark.parametrize("name,ptype,kw", [("LennardJones", ol.LJ, {}), ("Yukawa", ol.YUKAWA, {}),
                                            ("AthermalStar", ol.ATHERMAL_STAR, dict(f=8.0, alpha=0.3))])
def test_bruteforce_matches_oracle(mc, oracle, name, ptype, kw):
    """computeTotalEnergy / computeTotalVirial (thermodynamic_calculator.cpp:56-101), incl. a box < 3 cells."""
    for Lx, Ly in [(30.0, 27.5), (6.0, 4.0)]:
        rng = np.random.default_rng(4)
        g = np.stack(np.meshgrid(np.arange(10), np.arange(8), indexing="ij"), -1).reshape(-1, 2).astype(float)
        xy = (g + 0.5) * [Lx / 10, Ly / 8] + rng.uniform(-0.1, 0.1, g.shape)
        calc = oracle.calc(1.0, 0.0, ptype, 2.5, 0.0, kw.get("f", 0.0), kw.get("alpha", 0.0))
        with mc.Context(Lx, Ly, 2.5, potential=name, **kw) as ctx:
            U, W = ctx.calc_energy_virial_bruteforce(xy)
        Uo, Wo = oracle.total_energy(calc, Lx, Ly, xy), oracle.total_virial(calc, Lx, Ly, xy)
        assert abs(U - Uo) <= 1e-10 * max(1, abs(Uo)) and abs(W - Wo) <= 1e-10 * max(1, abs(Wo))


def test_reference_unit_tests_against_the_shims(host_built):
    """host_selftest.cpp restates unit_test_serial/{test_cell_list,test_calc,test_thermodynamic_celllist}.cpp
    and drives MonteCarlo through the reference's constructor."""
    out = subprocess.run([os.path.join(host_built, "host_selftest")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert re.search(r"\d+ checks, 0 failures", out.stdout), out.stdout


def test_input_txt_driver_and_log_formats(host_built, tmp_path):
    """mc2d_serial reads the reference's input format (sys_test_serial/input_z_0.024 keys) and writes the
    data line / LAMMPS dump the reference's PostProcessing parses (logging.cpp:44-90)."""
    inp = tmp_path / "input.txt"
    inp.write_text("\n".join([
        "Lx 40.0", "Ly 40.0", "N 400", "rcut 2.5", "T 1.0", "nSteps 60", "eSteps 20", "outputFreq 20",
        "cellUpdateFreq 10", "delta 0.1", "mu 0.1", "seed 42", "potential LennardJones", "ensemble GCMC",
        "out_xyz %s" % (tmp_path / "traj.dump"), "out_data %s" % (tmp_path / "data.log"), ""]))
    out = subprocess.run([os.path.join(host_built, "mc2d_serial"), str(inp)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Simulation completed." in out.stdout
    lines = (tmp_path / "data.log").read_text().strip().splitlines()
    assert len(lines) == 1 + 3  # one sample at the end of equilibration, three in production
    pat = re.compile(r"^Timestep:\d+ ,\tEnergy:-?\d+\.\d{5},\tTempreture:1\.00000,\tPressure:-?\d+\.\d{5},"
                     r"\tVolume:1600\.00000,\tdensity of particles:\d+\.\d{5},\tNumberofParticles:\d+$")
    for l in lines:
        assert pat.match(l), repr(l)
    # PostProcessing/gcmc_post_processing/data_loader.py:132-138 takes comma fields 0, 1, 3, 6
    f = lines[-1].split(",")
    n = int(f[6].split(":")[1])
    dump = (tmp_path / "traj.dump").read_text().splitlines()
    assert dump[0] == "ITEM: TIMESTEP" and dump[2] == "ITEM: NUMBER OF ATOMS"
    assert dump[4] == "ITEM: BOX BOUNDS pp pp ff" and dump[5] == "0 40" and dump[7] == "0 0"
    assert dump[8] == "ITEM: ATOMS id x y z"
    last = max(i for i, l in enumerate(dump) if l == "ITEM: NUMBER OF ATOMS")
    assert int(dump[last + 1]) == n and len(dump) - (last + 7) == n
    assert dump[last + 7].split()[0] == "1" and dump[last + 7].split()[3] == "0"
    # an unknown ensemble is refused loudly
    inp.write_text(inp.read_text().replace("ensemble GCMC", "ensemble muPT"))
    bad = subprocess.run([os.path.join(host_built, "mc2d_serial"), str(inp)], capture_output=True, text=True)
    assert bad.returncode == 1 and "Unrecognised ensemble" in bad.stderr


def test_slab_driver_and_log_formats(host_built, tmp_path):
    """mc2d_slabs = mpi_src/main.cpp on GPU slabs (it forks one process per rank itself: no mpirun).  Same input
    keys; thermo CSV as logging_data_mpi.cpp:56-67, LAMMPS dump with rank-as-type as logging_traj_mpi.cpp:268-303.
    One rank here (the test box has one GPU); tools/check_multi_gpu.py and the 2-rank run in profiles/ cover more."""
    base = tmp_path / "run"
    inp = tmp_path / "input.txt"
    inp.write_text("\n".join([
        "Lx 60.0", "Ly 50.0", "N 1500", "rcut 2.5", "T 1.0", "nSteps 40", "eSteps 20", "outputFreq 20",
        "cellUpdateFreq 10", "delta 0.1", "seed 7", "potential LennardJones", "ensemble NVT", "init lattice",
        "out_xyz %s.xyz" % base, "out_data %s.csv" % base, ""]))
    out = subprocess.run([os.path.join(host_built, "mc2d_slabs"), str(inp), "1"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Simulation completed." in out.stdout
    rows = (tmp_path / "run.thermo.csv").read_text().strip().splitlines()
    assert rows[0] == "timestep,N_global,rho,T,U,W,P,V"
    assert len(rows) == 1 + 3
    for r in rows[1:]:
        f = r.split(",")
        assert len(f) == 8 and int(f[1]) == 1500 and float(f[7]) == 3000.0 and float(f[3]) == 1.0
        assert abs(float(f[2]) - 0.5) < 1e-12
        assert abs(float(f[6]) - (0.5 * 1.0 + float(f[5]) / (2 * 3000.0))) < 1e-9  # P = rho T + W / 2V
    dump = (tmp_path / "run.dump").read_text().splitlines()
    assert dump[0] == "ITEM: TIMESTEP" and dump[3] == "1500" and dump[8] == "ITEM: ATOMS id type x y z"
    assert len(dump) == 3 * (9 + 1500)
    first = dump[9].split()
    assert first[0] == "1" and first[1] == "1" and first[4] == "0"
    # a bad ensemble is refused loudly
    inp.write_text(inp.read_text().replace("ensemble NVT", "ensemble NPT"))
    bad = subprocess.run([os.path.join(host_built, "mc2d_slabs"), str(inp), "1"], capture_output=True, text=True)
    assert bad.returncode != 0 and "supports NVT and GCMC" in bad.stderr


def test_input_txt_driver_npt(host_built, tmp_path):
    """`ensemble NPT` through the reference's MonteCarlo interface (MC.cpp:53-60,153-188): N stays fixed, the
    logged volume moves, density = N / V in every row, the final dump has the final box."""
    inp = tmp_path / "input.txt"
    inp.write_text("\n".join([
        "Lx 30.0", "Ly 30.0", "N 400", "rcut 2.5", "T 2.0", "P 1.0", "delta_V 0.05", "nSteps 300", "eSteps 100",
        "outputFreq 50", "cellUpdateFreq 10", "delta 0.2", "seed 5", "potential LennardJones", "ensemble NPT",
        "out_xyz %s" % (tmp_path / "traj.dump"), "out_data %s" % (tmp_path / "data.log"), ""]))
    out = subprocess.run([os.path.join(host_built, "mc2d_serial"), str(inp)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = (tmp_path / "data.log").read_text().strip().splitlines()
    vols = []
    for l in rows:
        f = dict(kv.strip().split(":") for kv in l.replace("\t", "").split(","))
        assert int(f["NumberofParticles"]) == 400
        V = float(f["Volume"])
        assert abs(float(f["density of particles"]) - 400 / V) < 1e-4
        vols.append(V)
    assert len(set(vols)) > 3 and all(300.0 < v < 3000.0 for v in vols)  # started at 900 (rho = 0.44), P = 1, T = 2
    dump = (tmp_path / "traj.dump").read_text().splitlines()
    last = max(i for i, l in enumerate(dump) if l == "ITEM: BOX BOUNDS pp pp ff")
    Lx = float(dump[last + 1].split()[1])
    assert abs(Lx * Lx - vols[-1]) < 1e-3 * vols[-1]


def test_input_txt_driver_gibbs(host_built, tmp_path):
    """`ensemble Gibbs` (serial_src/main.cpp:146-236): two boxes from Lx2 Ly2 N2 out_xyz2 out_data2; particles and
    total volume are conserved, both logs are written in the reference's format."""
    inp = tmp_path / "input.txt"
    inp.write_text("\n".join([
        "Lx 30.0", "Ly 30.0", "N 300", "Lx2 30.0", "Ly2 30.0", "N2 120", "rcut 2.5", "T 1.5", "delta_V 0.05",
        "nSteps 400", "eSteps 50", "outputFreq 100", "cellUpdateFreq 10", "delta 0.2", "seed 5",
        "potential WCA", "ensemble Gibbs",
        "out_xyz %s" % (tmp_path / "t1.dump"), "out_data %s" % (tmp_path / "d1.log"),
        "out_xyz2 %s" % (tmp_path / "t2.dump"), "out_data2 %s" % (tmp_path / "d2.log"), ""]))
    out = subprocess.run([os.path.join(host_built, "mc2d_serial"), str(inp)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    last = []
    for name in ("d1.log", "d2.log"):
        rows = (tmp_path / name).read_text().strip().splitlines()
        assert len(rows) >= 4
        f = dict(kv.strip().split(":") for kv in rows[-1].replace("\t", "").split(","))
        last.append((int(f["NumberofParticles"]), float(f["Volume"])))
    assert last[0][0] + last[1][0] == 420
    assert abs(last[0][1] + last[1][1] - 1800.0) < 1e-3
