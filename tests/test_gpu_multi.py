"""Multi-rank GPU parity (-m gpu; needs at least two devices on the box, skipped otherwise): tools/check_multi_gpu.py
under torchrun with 2 ranks (and 4 when the box has them).  What it pins is the reference contract of
mpi_src/MC_parallel.cpp:57-114 (4-colour sweep with ghost refresh after every colour),
particle_exchange.cpp:247-258,305-337 (ghost updates, ownership migration) and
thermodynamic_calculator_parallel.cpp:56-134 (U, W all-reduced, cross-rank pairs counted once) — as the reference's
own tests do (unit_test_mpi/test_mc_parallel.cpp:136-187,222-278; test_thermodynamic_calculator_parallel.cpp:378-514:
N conserved, particles stay in the box and on their owner, energies equal the serial calculator's) — but tightened to
BIT equality with the one-rank trajectory, which the counter-based Philox streams make possible."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ndev():
    from montecarlo_simulation_b200 import api

    return api.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_reproduce_the_one_gpu_trajectory_bit_for_bit(world):
    n = _ndev()
    if n < world:
        pytest.skip("needs %d CUDA devices, this box has %d" % (world, n))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=400)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MISMATCH" not in res.stdout and res.stdout.count("-> OK") >= 8


def test_peer_timeout_returns_a_status_instead_of_hanging():
    """A rank whose neighbour never shows up must get MC2D_ERR_PEER_TIMEOUT back (bounded in-kernel waits), not hang
    inside the cooperative sweep kernel: rank 1 uploads and then simply stops calling; rank 0 runs sweeps with a 300 ms
    wait budget."""
    if _ndev() < 2:
        pytest.skip("needs 2 CUDA devices")
    code = r"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import montecarlo_simulation_b200 as m
from montecarlo_simulation_b200 import api
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = 321.0
g = np.stack(np.meshgrid(np.arange(227), np.arange(227), indexing="ij"), -1).reshape(-1, 2).astype(float)
xy = (g + 0.5) * (L / 227)
ctx = m.Context(L, L, 2.5, T=1.0, delta=0.1, seed=11, device=local, rank=rank, nranks=2, tuning=dict(spin_timeout_ms=300))
ids = [api.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_init(ids[0])
ctx.upload(xy)
ctx.run_sweeps(2)
status = 0
if rank == 0:
    t0 = time.time()
    try:
        ctx.run_sweeps(3)          # the neighbour never launches its side
    except m.MC2DError as e:
        status = e.status
    print("rank 0: status %%d after %%.2f s" %% (status, time.time() - t0), flush=True)
    assert status == api.ERR_PEER_TIMEOUT, status
else:
    time.sleep(3.0)
dist.barrier()
os._exit(0)
""" % ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", "-c", code]
    # torchrun has no -c: write the script next to the test outputs
    path = os.path.join(ROOT, "gpurun_out", "_timeout_case.py")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(code)
    cmd = cmd[:-2] + [path]
    res = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True,
                         timeout=300)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "status 9" in res.stdout


def _dump_frames(path):
    """LAMMPS dump -> list of (timestep, natoms, sorted [(x, y)] as written, [types])."""
    lines = open(path).read().splitlines()
    frames, i = [], 0
    while i < len(lines):
        assert lines[i] == "ITEM: TIMESTEP"
        ts, n = int(lines[i + 1]), int(lines[i + 3])
        assert lines[i + 8] == "ITEM: ATOMS id type x y z"
        rows = [l.split() for l in lines[i + 9:i + 9 + n]]
        assert [int(r[0]) for r in rows] == list(range(1, n + 1))
        frames.append((ts, n, sorted((r[2], r[3]) for r in rows), [int(r[1]) for r in rows]))
        i += 9 + n
    return frames


@pytest.mark.parametrize("ensemble", ["NVT", "GCMC"])
def test_slab_driver_two_ranks_equals_one_rank(tmp_path, ensemble):
    """mc2d_slabs (= mpi_src/main.cpp:28-204 + MC_parallel.cpp:57-114 on GPU slabs) with 2 forked ranks writes the same
    thermo rows (logging_data_mpi.cpp:56-67) and — as a set, at 17 digits — the same particles into the dump
    (logging_traj_mpi.cpp:268-303) as with 1 rank; types are rank + 1; the gather-to-root file mode of the trajectory
    logger gives the same frames at 6 digits."""
    if _ndev() < 2:
        pytest.skip("needs 2 CUDA devices")
    host = os.path.join(ROOT, "montecarlo_simulation_b200", "host")
    subprocess.check_call(["make", "-s", "-C", host, "all"])

    def drive(tag, nranks, extra=()):
        base = tmp_path / tag
        inp = tmp_path / (tag + ".txt")
        inp.write_text("\n".join([
            "Lx 120.0", "Ly 50.0", "N 3000", "rcut 2.5", "T 1.0", "nSteps 30", "eSteps 10", "outputFreq 10",
            "cellUpdateFreq 10", "delta 0.1", "seed 7", "potential LennardJones", "ensemble " + ensemble, "mu 0.3",
            "init lattice", "out_xyz %s.xyz" % base, "out_data %s.csv" % base, *extra, ""]))
        out = subprocess.run([os.path.join(host, "mc2d_slabs"), str(inp), str(nranks)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return base

    one, two = drive("one", 1), drive("two", 2)
    rows1 = open(str(one) + ".thermo.csv").read().splitlines()
    rows2 = open(str(two) + ".thermo.csv").read().splitlines()
    assert rows1[0] == rows2[0] == "timestep,N_global,rho,T,U,W,P,V" and len(rows1) == len(rows2) == 5
    for a, b in zip(rows1[1:], rows2[1:]):
        fa, fb = [float(v) for v in a.split(",")], [float(v) for v in b.split(",")]
        assert fa[:4] == fb[:4] and fa[7] == fb[7]  # timestep, N, rho, T, V
        for k in (4, 5, 6):  # U, W, P: summation order only
            assert abs(fa[k] - fb[k]) <= 1e-11 * max(1.0, abs(fa[k]))
    f1, f2 = _dump_frames(str(one) + ".dump"), _dump_frames(str(two) + ".dump")
    assert len(f1) == len(f2) == 4
    for (ts1, n1, xy1, t1), (ts2, n2, xy2, t2) in zip(f1, f2):
        assert ts1 == ts2 and n1 == n2 and xy1 == xy2
        assert set(t1) == {1} and set(t2) == {1, 2} and t2 == sorted(t2)
    if ensemble == "GCMC":
        assert len({f[1] for f in f1}) > 1  # the particle number moved
    gat = drive("gat", 2, ["traj_mode gather"])
    fg = _dump_frames(str(gat) + ".dump")
    assert [(f[0], f[1], f[3]) for f in fg] == [(f[0], f[1], f[3]) for f in f2]
    per = drive("per", 2, ["traj_mode perrank"])
    assert os.path.exists(str(per) + ".rank0000.dump") and os.path.exists(str(per) + ".rank0001.dump")
