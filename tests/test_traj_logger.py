"""LoggingTrajSlabs (host/slab_traj_logger.hpp) against the reference's LoggingTrajMPI (mpi_src/logging_traj_mpi.cpp:98-337):
all three file modes x {xyz, dump} x {plain, rank_as_type}, an empty rank, several frames, append — byte for byte.
No GPU involved: the logger only formats what mc2d_download hands it, the ranks are forked as mc2d_slabs forks them."""
import filecmp
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "montecarlo_simulation_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden", "traj")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "traj_ref")


@pytest.fixture(scope="module")
def selftest():
    subprocess.check_call(["make", "-s", "-C", HOST, os.path.join(HOST, "slab_traj_selftest")])
    return os.path.join(HOST, "slab_traj_selftest")


def golden_files():
    return sorted(f for f in os.listdir(GOLD) if f.endswith((".xyz", ".dump")))


def run(binary, nranks, outdir):
    os.makedirs(outdir, exist_ok=True)
    subprocess.run([binary, str(nranks), str(outdir)], check=True, timeout=120, capture_output=True)


def test_every_mode_matches_the_reference_files(selftest, tmp_path):
    names = [f for f in golden_files() if not f.startswith("serial.")]
    assert len(names) == 20  # 2 modes x 2 x 2 shared files + 3 ranks x 2 x 2 per-rank files
    run(selftest, 3, tmp_path)
    for f in names:
        assert filecmp.cmp(os.path.join(GOLD, f), tmp_path / f, shallow=False), f
    # append = true continues the file of an earlier logger: same bytes as three frames in one go
    assert filecmp.cmp(tmp_path / "append_gather.dump", tmp_path / "gather_typed.dump", shallow=False)
    assert filecmp.cmp(tmp_path / "append_mpiio.dump", tmp_path / "mpiio_typed.dump", shallow=False)


def test_format_details(selftest, tmp_path):
    """What PostProcessing/ovito and the reference's own readers rely on (logging_traj_mpi.cpp:268-303)."""
    run(selftest, 3, tmp_path)
    lines = open(tmp_path / "mpiio_typed.dump").read().split("\n")
    assert lines[0] == "ITEM: TIMESTEP" and lines[1] == "0" and lines[2] == "ITEM: NUMBER OF ATOMS"
    n = int(lines[3])
    assert n == 3 + 0 + 7 and lines[8] == "ITEM: ATOMS id type x y z"
    body = [l.split() for l in lines[9:9 + n]]
    assert [int(r[0]) for r in body] == list(range(1, n + 1))  # global ids across the ranks
    assert [int(r[1]) for r in body] == [1] * 3 + [3] * 7  # type = rank + 1, rank 1 is empty
    assert all(r[4] == "0" for r in body)
    assert len(body[0][2]) >= 17  # 17 significant digits in the shared-file mode
    first = open(tmp_path / "gather.xyz").read().split("\n")
    assert first[0] == "10" and first[1] == "Lx=20 Ly=12.3456789"
    assert open(tmp_path / "perrank.rank0001.xyz").read().startswith("0\nrank=1\n")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/traj_ref not built (needs /root/reference)")
@pytest.mark.parametrize("nranks", [1, 3, 4])
def test_against_the_reference_class_itself(selftest, tmp_path, nranks):
    ref_dir, mine_dir = tmp_path / "ref", tmp_path / "mine"
    run(REF_BIN, nranks, ref_dir)
    run(selftest, nranks, mine_dir)
    names = sorted(os.listdir(ref_dir))
    assert len(names) == 8 + 4 * nranks
    for f in names:
        assert filecmp.cmp(ref_dir / f, mine_dir / f, shallow=False), f
        if nranks == 3:  # the committed copies are current
            assert filecmp.cmp(ref_dir / f, os.path.join(GOLD, f), shallow=False), f


def _serial_frames():
    """the frames host/slab_traj_selftest.cpp hands to the serial Logging class (traj_cases.hpp: ranks 0, 2, 0)"""
    import numpy as np

    def frame_xy(rank, frame):
        n = (0 if rank == 1 else 2 * rank + 3) + frame
        xy = []
        for i in range(n):
            x = 0.123456789012345678 + 1.7320508075688772 * i + 3.0 * rank + 0.001 * frame
            y = 12.345678901234567 / (1.0 + i + rank) + 1e-7 * frame
            if i == 1:
                x = 1.0e-9 * (rank + 1)
            if i == 2:
                y = 12.0
            xy.append((x, y))
        return np.array(xy, dtype=np.float64).reshape(-1, 2)

    return [frame_xy(0, 0), frame_xy(2, 1), frame_xy(0, 2)], [0, 100, 200]


def test_serial_position_logs(selftest, tmp_path):
    """Logging::logPositions_xyz / logPositions_dump of the serial shim (host/mc2d_host.cpp) against the reference's
    (serial_src/logging.cpp:30-79): byte for byte with the committed files and, when oracle/_ref is built, with the
    unmodified reference class fed the same doubles."""
    run(selftest, 3, tmp_path)
    for f in ("serial.xyz", "serial.dump"):
        assert filecmp.cmp(os.path.join(GOLD, f), tmp_path / f, shallow=False), f
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    if not ol.have_ref() or not hasattr(ol.Ref().lib, "ref_log_positions"):
        pytest.skip("oracle/_ref/libmc2d_ref.so not built with ref_log_positions")
    frames, ts = _serial_frames()
    ref = ol.Ref()
    ref.log_positions(tmp_path / "ref.xyz", tmp_path / "ref_xyz.dat", False, 20.0, 12.345678901234567, frames, ts)
    ref.log_positions(tmp_path / "ref.dump", tmp_path / "ref_dump.dat", True, 20.0, 12.345678901234567, frames, ts)
    assert filecmp.cmp(tmp_path / "ref.xyz", tmp_path / "serial.xyz", shallow=False)
    assert filecmp.cmp(tmp_path / "ref.dump", tmp_path / "serial.dump", shallow=False)
