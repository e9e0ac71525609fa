"""CPU tests (no GPU): the C-ABI library loads and exports exactly what include/mc2d.h declares,
the pure-host geometry / schedule entry points behave, and compute calls fail LOUDLY without a
device (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from montecarlo_simulation_b200 import api as a
    from montecarlo_simulation_b200 import build

    build.build_library()  # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    a.load_library()
    return a


def test_header_and_library_export_the_same_symbols(api):
    hdr = open(os.path.join(ROOT, "include", "mc2d.h")).read()
    declared = set(re.findall(r"\b(mc2d_[a-z0-9_]+)\s*\(", hdr))
    out = subprocess.check_output(["nm", "-D", "--defined-only", api.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T mc2d_" in l}
    assert declared == exported == set(api.EXPORTS)
    L = api.load_library()
    for name in declared:
        assert hasattr(L, name)
    assert L.mc2d_abi_version() == 2


def test_library_is_sm100a_only(api):
    out = subprocess.check_output(["cuobjdump", "-lelf", api.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_reference_grid_matches_cell_list_ctor(api, oracle):
    """cell_list.cpp:5-15"""
    for Lx, Ly, rc in [(45.254834, 45.254834, 2.5), (10.0, 17.32, 2.5), (3.0, 3.0, 5.0), (1000.0, 800.0, 5.0)]:
        assert api.reference_grid(Lx, Ly, rc) == oracle.cell_dims(Lx, Ly, rc)


def test_sweep_grid_even_per_slab(api):
    for nranks in (1, 2, 4, 8):
        nx, ny, dx, dy = api.sweep_grid(5656.854, 5656.854, 2.5, nranks)
        assert nx % (2 * nranks) == 0 and ny % 2 == 0 and dx >= 2.5 and dy >= 2.5
        assert nx + 2 * nranks > int(5656.854 / 2.5)  # largest admissible
        cols = [api.slab_columns(nx, nranks, r) for r in range(nranks)]
        assert cols[0][0] == 0 and sum(c[1] for c in cols) == nx
        for (a0, an), (b0, _) in zip(cols, cols[1:]):
            assert a0 + an == b0 and an % 2 == 0 and a0 % 2 == 0
    with pytest.raises(api.MC2DError) as e:
        api.sweep_grid(4.0, 40.0, 2.5, 1)  # one cell across: no checkerboard
    assert e.value.status == api.ERR_GEOMETRY
    with pytest.raises(api.MC2DError):
        api.sweep_grid(20.0, 20.0, 2.5, 8)  # 8 slabs need 16 columns


def test_halo_routes_pair_up(api):
    """rank r sends to s exactly when s receives from r; left/right follows the column parity."""
    for nranks in (2, 4, 8):
        for colour in range(4):
            routes = [api.halo_route(colour, r, nranks) for r in range(nranks)]
            for r, (send_left, to, frm) in enumerate(routes):
                assert send_left == (colour % 2 == 0)
                assert to == (r - 1) % nranks if send_left else to == (r + 1) % nranks
                assert routes[to][2] == r and routes[frm][1] == r


def test_schedule_is_a_pure_function_of_seed_and_sweep(api):
    seen = set()
    for sweep in range(200):
        order, sx, sy = api.sweep_schedule(42, sweep, 2.5, 2.6)
        assert sorted(order) == [0, 1, 2, 3] and 0 <= sx < 2.5 and 0 <= sy < 2.6
        assert (order, sx, sy) == api.sweep_schedule(42, sweep, 2.5, 2.6)
        seen.add(tuple(order))
    assert len(seen) == 24  # all colour permutations occur
    assert api.sweep_schedule(42, 3, 2.5, 2.6) != api.sweep_schedule(43, 3, 2.5, 2.6)
    assert api.sweep_schedule(42, 3, 2.5, 2.6, shift_grid=False)[1:] == (0.0, 0.0)


def test_no_device_fails_loudly(api):
    """There is no CPU fallback: without a GPU the context cannot be created."""
    import montecarlo_simulation_b200 as m

    if api.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(m.MC2DError) as e:
        m.Context(20.0, 20.0, 2.5)
    assert e.value.status == api.ERR_NO_DEVICE
    assert "no CPU" in str(e.value)


def test_invalid_parameters_rejected(api):
    import montecarlo_simulation_b200 as m

    for kw in (dict(Lx=-1.0), dict(rcut=0.0), dict(T=0.0), dict(potential=9), dict(rank=2, nranks=2)):
        args = dict(Lx=20.0, Ly=20.0, rcut=2.5)
        args.update(kw)
        with pytest.raises(m.MC2DError) as e:
            m.Context(**args)
        assert e.value.status == api.ERR_INVALID


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "montecarlo_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "libmc2d_oracle" not in text and "libmc2d_ref" not in text, f
                assert "mc2d_oracle" not in text, f
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libmc2d.so")]).decode()
    assert "oracle" not in out
