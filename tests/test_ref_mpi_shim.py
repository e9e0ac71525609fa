"""CPU test of the threads-as-ranks MPI shim (oracle/mpi_shim): the reference's MPI driver (mpi_src/main.cpp and friends,
compiled unmodified into oracle/_ref/mc2d_ref_mpi) must run to completion at 1, 2 and 4 "ranks", conserve N and write its
thermo line and dump — i.e. the ~25 MPI calls it makes (point-to-point halo exchange, Alltoallv migration, Allreduce,
Bcast, Exscan, MPI-IO) behave.  The binary only feeds bench.py's `cpu_baseline_mpi`; it is never ground truth."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "mc2d_ref_mpi")


@pytest.mark.parametrize("ranks", [1, 2, 4])
def test_reference_mpi_driver_runs_on_the_shim(tmp_path, ranks):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/mc2d_ref_mpi not built (needs /root/reference at build time)")
    # eSteps = 0: a second MonteCarloNVT_MPI::run() call sizes its attempt loop from owned.size() AFTER migration
    # (MC_parallel.cpp:67), which differs between ranks, and the ranks then disagree on the number of collectives — a
    # defect of the reference (it would hang under a real MPI too), so the baseline uses one run() call.
    n, L = 900, 42.426407
    (tmp_path / "input.txt").write_text(
        "Lx %g\nLy %g\nN %d\nrcut 2.5\nT 1.0\nnSteps 3\neSteps 0\noutputFreq 1\ncellUpdateFreq 10\n"
        "potential LennardJones\nout_xyz run.xyz\nout_data run.dat\ndelta 0.1\nseed 7\nensemble NVT\n" % (L, L, n))
    res = subprocess.run([EXE, str(ranks), "input.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    timing = [l for l in res.stdout.splitlines() if l.startswith("MC2D_MPI_TIMING")]
    assert len(timing) == 1 and "attempted=%d" % (4 * n * 3) in timing[0]
    rows = (tmp_path / "run.thermo.csv").read_text().strip().splitlines()
    assert rows[0].startswith("timestep,N_global")
    assert all(int(r.split(",")[1]) == n for r in rows[1:]) and len(rows) >= 2  # N conserved across ranks
    dump = (tmp_path / "run.dump").read_text().splitlines()
    assert dump[0] == "ITEM: TIMESTEP" and int(dump[3]) == n
