"""GPU parity tests: the CUDA path, called through the C ABI (include/mc2d.h), against the CPU
oracle on the same seeded inputs, against the committed golden fixtures, and — at the BASELINE
sizes — through size-independent properties.

Tolerances (BASELINE.json north_star): total energy, virial and per-trial dE within 1e-10
relative (fp64); cell assignment and neighbour counts bit-exact.
"""
import math

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

REL = 1e-10


@pytest.fixture(scope="module")
def mc():
    import montecarlo_simulation_b200 as m
    from montecarlo_simulation_b200 import api

    assert api.device_count() > 0, "GPU tests need a CUDA device; libmc2d has no CPU fallback"
    return m


def lattice(n_side_x, n_side_y, Lx, Ly, jitter, seed):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(n_side_x), np.arange(n_side_y), indexing="ij"), -1).reshape(-1, 2).astype(float)
    xy = (g + 0.5) * np.array([Lx / n_side_x, Ly / n_side_y]) + rng.uniform(-jitter, jitter, size=g.shape)
    xy[:, 0] %= Lx
    xy[:, 1] %= Ly
    return np.ascontiguousarray(xy)


def relerr(a, b):
    return abs(a - b) / max(1.0, abs(b))


POTS = [
    ("LennardJones", ol.LJ, dict(rcut=2.5)),
    ("WCA", ol.WCA, dict(rcut=2 ** (1 / 6))),
    ("Yukawa", ol.YUKAWA, dict(rcut=2.5)),
    ("AthermalStar", ol.ATHERMAL_STAR, dict(rcut=2.5, f=8.0, alpha=0.3)),
    ("ThermalStar", ol.THERMAL_STAR, dict(rcut=2.5, f=6.0, alpha=0.4, A0=1.5, kappa=1.1)),
    ("Ideal", ol.IDEAL, dict(rcut=2.5)),
]


def make_pair(mc, oracle, name, ptype, kw, Lx, Ly, T=1.0, z=0.0, **ctx_kw):
    k = dict(f=0.0, alpha=0.0, A0=0.0, kappa=0.0)
    k.update(kw)
    ctx = mc.Context(Lx, Ly, k["rcut"], potential=name, T=T, activity=z, f=k["f"], alpha=k["alpha"], A0=k["A0"],
                     kappa=k["kappa"], **ctx_kw)
    calc = oracle.calc(T, 0.0, ptype, k["rcut"], z, k["f"], k["alpha"], k["A0"], k["kappa"])
    return ctx, calc


# ---- cell assignment and neighbour counts: bit-exact ---------------------------------------------

@pytest.mark.parametrize("Lx,Ly,rcut", [(31.0, 26.0, 2.5), (10.0, 10.0, 2.5), (9.306, 10.746, 2.5), (7.3, 4.9, 2.5),
                                         (3.0, 3.0, 5.0), (45.254834, 45.254834, 2.5), (100.0, 10.0, 2.5)])
def test_cell_ids_and_neighbor_counts_bit_exact(mc, oracle, Lx, Ly, rcut):
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 1, (2000, 2)) * [Lx, Ly]
    xy[0] = [0.0, 0.0]
    xy[1] = [np.nextafter(Lx, 0), np.nextafter(Ly, 0)]
    xy[2] = [Lx, Ly]  # applyPBC can return exactly L (initial.cpp:22-25)
    # put particles exactly on cell boundaries
    nx, ny, dx, dy = oracle.cell_dims(Lx, Ly, rcut)
    for k in range(min(nx, 8)):
        xy[10 + k] = [k * dx, (k % ny) * dy]
    with mc.Context(Lx, Ly, rcut) as ctx:
        assert (ctx.calc_cell_ids(xy) == oracle.cell_ids(Lx, Ly, rcut, xy)).all()
        cnt, eloc = ctx.calc_neighbor_counts(xy, with_energy=True)
        assert (cnt == oracle.neighbor_counts(Lx, Ly, rcut, xy)).all()


def test_inclusive_cutoff_exactly_at_rcut(mc, oracle):
    """test_cell_list.cpp:32-53: a pair at exactly r = rcut is a neighbour; just beyond is not."""
    P = np.array([[1.0, 1.0], [2.0, 1.0], [3.1, 1.0], [4.1 + 1e-12, 1.0]])
    with mc.Context(5.0, 5.0, 1.0) as ctx:
        cnt = ctx.calc_neighbor_counts(P)
    assert cnt.tolist() == oracle.neighbor_counts(5.0, 5.0, 1.0, P).tolist() == [1, 1, 0, 0]


# ---- U, W on the reference grid vs the oracle ------------------------------------------------------

@pytest.mark.parametrize("name,ptype,kw", POTS)
def test_energy_virial_vs_oracle(mc, oracle, name, ptype, kw):
    Lx, Ly = 61.0, 47.5
    xy = lattice(48, 40, Lx, Ly, 0.25, 100 + ptype)
    ctx, calc = make_pair(mc, oracle, name, ptype, kw, Lx, Ly, T=1.3)
    with ctx:
        U, W = ctx.calc_energy_virial(xy)
        Uo = oracle.total_energy_celllist(calc, Lx, Ly, xy)
        Wo = oracle.total_virial_celllist(calc, Lx, Ly, xy)
        assert relerr(U, Uo) < REL and relerr(W, Wo) < REL
        assert relerr(ctx.calc_pressure(xy), oracle.pressure_celllist(calc, Lx, Ly, xy)) < REL
        # per-particle local energies (computeLocalEnergy over getNeighbors(i))
        cnt, eloc = ctx.calc_neighbor_counts(xy, with_energy=True)
        idx = [0, 17, 333, len(xy) - 1]
        eo = oracle.local_energies(calc, Lx, Ly, xy, idx)
        for i, e in zip(idx, eo):
            assert relerr(eloc[i], e) < REL


@pytest.mark.parametrize("Lx,Ly,rcut", [(7.3, 4.9, 2.5), (3.0, 3.0, 5.0), (6.0, 12.0, 2.5)])
def test_degenerate_grids_match_reference_quirk(mc, oracle, Lx, Ly, rcut):
    """nx < 3: the reference's 3x3 scan revisits cells and double counts; the calculator hook
    reproduces it (same inputs -> same numbers), bit-exact counts included."""
    xy = np.random.default_rng(9).uniform(0, 1, (40, 2)) * [Lx, Ly]
    calc = oracle.calc(1.0, 0.0, ol.YUKAWA, rcut)
    with mc.Context(Lx, Ly, rcut, potential="Yukawa") as ctx:
        U, W = ctx.calc_energy_virial(xy)
        assert relerr(U, oracle.total_energy_celllist(calc, Lx, Ly, xy)) < REL
        assert relerr(W, oracle.total_virial_celllist(calc, Lx, Ly, xy)) < REL
        assert (ctx.calc_neighbor_counts(xy) == oracle.neighbor_counts(Lx, Ly, rcut, xy)).all()


def test_empty_and_single(mc):
    with mc.Context(20.0, 20.0, 2.5) as ctx:
        assert ctx.calc_energy_virial(np.zeros((0, 2))) == (0.0, 0.0)
        assert ctx.calc_energy_virial(np.array([[1.0, 1.0]])) == (0.0, 0.0)
        ctx.upload(np.zeros((0, 2)))
        assert ctx.num_particles() == 0
        assert ctx.total_energy_virial()[0] == 0.0
        st = ctx.run_sweeps(3)
        assert st["n_particles"] == 0 and st["attempted"][0] == 0


def test_out_of_box_rejected(mc):
    with mc.Context(20.0, 20.0, 2.5) as ctx:
        with pytest.raises(mc.MC2DError) as e:
            ctx.upload(np.array([[1.0, 1.0], [25.0, 3.0]]))
        assert e.value.status == 6
        with pytest.raises(mc.MC2DError):
            ctx.calc_energy_virial(np.array([[float("nan"), 1.0]]))


def test_known_answers_through_gpu(mc):
    """unit_test_serial/test_calc.cpp:16-120 through the GPU calculator (boxes >= 3 cells wide)."""
    r0 = 2.0 ** (1 / 6) - 1e-14
    with mc.Context(12 * r0, 12 * r0, 3 * r0) as ctx:
        U, W = ctx.calc_energy_virial(np.array([[0, 0], [r0, 0]]))
        assert U == pytest.approx(-1.0) and W + 1 == pytest.approx(1.0)
    with mc.Context(20 * r0, 20 * r0, 5 * r0) as ctx:
        U, W = ctx.calc_energy_virial(np.array([[0, 0], [r0, 0], [r0 / 2, r0 * math.sqrt(3) / 2]]))
        assert U == pytest.approx(-3.0) and W + 1 == pytest.approx(1.0)
    with mc.Context(16.0, 16.0, 5.0, potential="Yukawa") as ctx:
        U, W = ctx.calc_energy_virial(np.array([[0, 0], [1.0, 0]]))
        assert U == pytest.approx(math.exp(-1)) and W == pytest.approx(2 * math.exp(-1))
    rc = 2.0 ** (1 / 6)
    with mc.Context(3 * rc, 3 * rc, rc, potential="WCA") as ctx:
        assert ctx.calc_energy_virial(np.array([[0, 0], [1.0, 0]]))[0] == pytest.approx(1.0)
        assert ctx.calc_energy_virial(np.array([[0, 0], [1.5, 0]]))[0] == 0.0
    P = np.array([[i, j] for i in range(4) for j in range(4)], dtype=float)
    with mc.Context(4.0, 4.0, 1.0001, potential="AthermalStar", f=10.0, alpha=0.5) as ctx:
        U, W = ctx.calc_energy_virial(P)
        assert U == pytest.approx(32 * (900 + 2) / 24 * 0.5) and W == pytest.approx(32 * (900 + 2) / 24)
        assert ctx.calc_pressure(P) == pytest.approx(1.0 + 32 * (902 / 24) / 16 / 2)


# ---- golden fixtures -----------------------------------------------------------------------------------

def test_golden_reference_vectors(mc, golden):
    g = golden["ref"]
    cfg = g["config"]
    xy = np.array(cfg["xy"])
    Lx, Ly = cfg["Lx"], cfg["Ly"]
    names = {0: "LennardJones", 1: "WCA", 2: "Yukawa", 3: "AthermalStar", 4: "ThermalStar", 5: "Ideal"}
    for name, o in g["observables"].items():
        kw = o["kw"]
        with mc.Context(Lx, Ly, kw["rcut"], potential=names[o["ptype"]], T=kw["T"], f=kw.get("f", 0.0),
                        alpha=kw.get("alpha", 0.0), A0=kw.get("A0", 0.0), kappa=kw.get("kappa", 0.0)) as ctx:
            U, W = ctx.calc_energy_virial(xy)
            assert relerr(U, o["U_cl"]) < REL and relerr(W, o["W_cl"]) < REL, name
            assert relerr(ctx.calc_pressure(xy), o["P_cl"]) < REL, name
    with mc.Context(Lx, Ly, cfg["rcut"]) as ctx:
        assert ctx.calc_cell_ids(xy).tolist() == g["cell_ids"]
        assert ctx.calc_neighbor_counts(xy).tolist() == g["neighbor_counts"]
        pr = g["probes"]["displace"]
        idx = [p["idx"] for p in pr]
        e_old = ctx.calc_point_energies(xy, xy[idx], exclude=idx)
        e_new = ctx.calc_point_energies(xy, [p["new"] for p in pr], exclude=idx)
        for p, a, b in zip(pr, e_old, e_new):
            assert relerr(a, p["U_old"]) < REL and relerr(b, p["U_new"]) < REL
            assert relerr(b - a, p["U_new"] - p["U_old"]) < REL
        ins = g["probes"]["insert"]
        e = ctx.calc_point_energies(xy, [p["pt"] for p in ins])
        for p, a in zip(ins, e):
            assert relerr(a, p["U"]) < REL


def test_sanity_check_golden_csv(mc, oracle, golden):
    """sanity_check/sanity_parallel_data.csv rows (reference MPI np=10, N=24000 AthermalStar)."""
    s = golden["sanity"]
    cfg = s["config"]
    with mc.Context(cfg["Lx"], cfg["Ly"], cfg["rcut"], potential="AthermalStar", T=cfg["T"], f=cfg["f"],
                    alpha=cfg["alpha"], A0=cfg["A0"], kappa=cfg["kappa"]) as ctx:
        for frame in (0, 1, 2, 3, 100, 199):
            xy = oracle.sanity_frame(cfg["Lx"], cfg["Ly"], 5, 2, cfg["N"], cfg["seed0"] + frame)
            row = s["rows"][str(frame)]
            U, W = ctx.calc_energy_virial(xy)
            assert relerr(U, row["U"]) < REL and relerr(W, row["W"]) < REL
            assert relerr(ctx.calc_pressure(xy), row["P"]) < REL


# ---- dE hooks ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,ptype,kw", POTS[:5])
def test_point_energy_probes_vs_oracle(mc, oracle, name, ptype, kw):
    Lx, Ly = 33.0, 29.0
    xy = lattice(26, 22, Lx, Ly, 0.25, 5)
    rng = np.random.default_rng(6)
    ctx, calc = make_pair(mc, oracle, name, ptype, kw, Lx, Ly)
    with ctx:
        idx = rng.integers(0, len(xy), 24)
        new = xy[idx] + rng.uniform(-0.1, 0.1, (24, 2))
        new[:, 0] %= Lx
        new[:, 1] %= Ly
        e_old, c_old = ctx.calc_point_energies(xy, xy[idx], exclude=idx, with_counts=True)
        e_new = ctx.calc_point_energies(xy, new, exclude=idx)
        pts = rng.uniform(0, 1, (24, 2)) * [Lx, Ly]
        e_ins = ctx.calc_point_energies(xy, pts)
        ncnt = oracle.neighbor_counts(Lx, Ly, kw["rcut"], xy)
        for k, i in enumerate(idx):
            uo = oracle.local_energy(calc, Lx, Ly, xy, int(i))
            un = oracle.point_energy(calc, Lx, Ly, xy, new[k, 0], new[k, 1], int(i))
            assert relerr(e_old[k], uo) < REL and relerr(e_new[k], un) < REL
            assert c_old[k] == ncnt[i]
            assert abs((e_new[k] - e_old[k]) - (un - uo)) <= REL * max(1.0, abs(un), abs(uo))
            ui = oracle.point_energy(calc, Lx, Ly, xy, pts[k, 0], pts[k, 1], -1)
            assert relerr(e_ins[k], ui) < REL * max(1.0, 1.0)


# ---- slot layout: upload / download / energy ---------------------------------------------------------------------

def sort_rows(a):
    return a[np.lexsort((a[:, 1], a[:, 0]))]


@pytest.mark.parametrize("Lx,Ly,nx,ny", [(45.254834, 45.254834, 32, 32), (100.0, 10.0, 70, 7), (61.0, 47.5, 48, 40)])
def test_upload_download_roundtrip_and_slot_energy(mc, oracle, Lx, Ly, nx, ny):
    xy = lattice(nx, ny, Lx, Ly, 0.25, 12)
    calc = oracle.calc(1.0, 0.0, ol.LJ, 2.5)
    with mc.Context(Lx, Ly, 2.5) as ctx:
        ids = np.arange(len(xy), dtype=np.int32)[::-1].copy()
        ctx.upload(xy, ids)
        assert ctx.num_particles() == len(xy)
        out, oid = ctx.download(with_ids=True)
        assert (sort_rows(out) == sort_rows(xy)).all()  # bit-exact positions
        assert (xy[len(xy) - 1 - oid] == out).all()     # ids travel with their particle
        U, W, N = ctx.total_energy_virial()
        assert N == len(xy)
        assert relerr(U, oracle.total_energy_celllist(calc, Lx, Ly, xy)) < REL
        assert relerr(W, oracle.total_virial_celllist(calc, Lx, Ly, xy)) < REL
        assert relerr(ctx.stats()["energy"], U) < 1e-15  # initial running energy (MC.cpp:26)


@pytest.mark.parametrize("with_ids", [True, False])
def test_reupload_without_sort_gives_identical_slots(mc, monkeypatch, with_ids):
    """Uploads bin directly (k_bin_direct + k_order_cells; the first one counts first to size the storage).  The slot
    arrays must be bit-identical to those of the key-sort path (tuning.upload_sort: counting sort -> CSR -> slots): same
    download order, same ids, same trajectory afterwards."""
    Lx = Ly = 61.0
    rng = np.random.default_rng(5)
    xy = lattice(48, 40, Lx, Ly, 0.3, 7)
    xy = xy[rng.permutation(len(xy))]  # input order unrelated to the cells
    ids = (np.arange(len(xy), dtype=np.int32) * 7 + 3) if with_ids else None

    def run(reuploads, sort_only):
        with mc.Context(Lx, Ly, 2.5, T=1.0, activity=0.3, delta=0.15, seed=11, tuning=dict(upload_sort=int(sort_only))) as ctx:
            ctx.upload(xy, ids)
            first = ctx.info().launches
            for _ in range(reuploads):
                before = ctx.info().launches
                ctx.upload(xy, ids)
                again = ctx.info().launches - before
            U0 = ctx.total_energy_virial()[0]
            d0 = ctx.download(with_ids=with_ids)
            st = ctx.run_sweeps(3, gc_per_cell=1)
            d1 = ctx.download(with_ids=with_ids)
            return first, (again if reuploads else None), U0, d0, st, d1

    first, again, U0, d0, st, d1 = run(2, False)
    assert again == first - 2  # a re-upload skips the count pass (keys + max) that sized the storage: no sort anywhere
    for ref in (run(0, False), run(2, True)):
        if ref[1] is not None:
            assert ref[1] >= first - 1  # tuning.upload_sort = 1: the sorted path again
        assert U0 == ref[2]
        if with_ids:
            assert (d0[0] == ref[3][0]).all() and (d0[1] == ref[3][1]).all()
            assert d1[0].shape == ref[5][0].shape and (d1[0] == ref[5][0]).all() and (d1[1] == ref[5][1]).all()
        else:
            assert (d0 == ref[3]).all()
            assert d1.shape == ref[5].shape and (d1 == ref[5]).all()
        assert st["accepted"] == ref[4]["accepted"] and st["energy"] == ref[4]["energy"]


def test_reupload_of_a_denser_configuration_falls_back_to_the_sort(mc, oracle):
    """storage made for a dilute system, then a configuration whose cells need more slots: the direct path notices
    (largest cell count) and the sorted path re-allocates."""
    Lx = Ly = 40.0
    calc = oracle.calc(1.0, 0.0, ol.LJ, 2.5)
    dilute = lattice(16, 16, Lx, Ly, 0.2, 3)
    dense = lattice(60, 60, Lx, Ly, 0.02, 4)  # ~14 per cell: more than the 25 % headroom of the dilute storage
    with mc.Context(Lx, Ly, 2.5) as ctx:
        ctx.upload(dilute)
        k0 = ctx.info().cell_capacity
        a0 = ctx.info().slot_allocs
        ctx.upload(dense)
        assert ctx.info().slot_allocs == a0 + 1  # re-allocated by the sorted path
        assert ctx.info().cell_capacity > k0 and ctx.num_particles() == len(dense)
        assert relerr(ctx.total_energy_virial()[0], oracle.total_energy_celllist(calc, Lx, Ly, dense)) < REL
        ctx.upload(dilute)
        assert ctx.num_particles() == len(dilute)
        assert relerr(ctx.total_energy_virial()[0], oracle.total_energy_celllist(calc, Lx, Ly, dilute)) < REL
        with pytest.raises(mc.MC2DError):
            ctx.upload(np.array([[1.0, 1.0], [Lx + 5.0, 3.0]]))
        with pytest.raises(mc.MC2DError):  # a failed direct upload leaves no configuration behind
            ctx.total_energy_virial()
        ctx.upload(dilute)
        assert ctx.num_particles() == len(dilute)
        ctx.upload(np.zeros((0, 2)))  # an empty re-upload
        assert ctx.num_particles() == 0 and ctx.total_energy_virial()[0] == 0.0
        ctx.upload(dilute[:1])
        assert ctx.num_particles() == 1 and (ctx.download() == dilute[:1]).all()


# ---- the step loop -----------------------------------------------------------------------------------------------------

def test_nvt_sweeps_conserve_and_track_energy(mc, oracle):
    Lx = Ly = 45.254834
    xy = lattice(32, 32, Lx, Ly, 0.2, 42)
    calc = oracle.calc(1.0, 0.0, ol.LJ, 2.5)
    with mc.Context(Lx, Ly, 2.5, T=1.0, delta=0.1, seed=42) as ctx:
        ctx.upload(xy)
        st = ctx.run_sweeps(50)
        assert st["n_particles"] == 1024 and st["sweeps_done"] == 50
        assert st["attempted"][0] == 50 * 1024 and st["attempted"][1] == st["attempted"][2] == 0
        assert 0.2 < st["accepted"][0] / st["attempted"][0] < 0.95
        out = ctx.download()
        assert len(out) == 1024 and (out >= 0).all() and (out[:, 0] <= Lx).all() and (out[:, 1] <= Ly).all()
        U = ctx.total_energy_virial()[0]
        assert relerr(st["energy"], U) < 1e-9            # running energy == recomputed (MC.cpp:62-71)
        assert relerr(U, oracle.total_energy_celllist(calc, Lx, Ly, out)) < REL
        assert abs(U - oracle.total_energy_celllist(calc, Lx, Ly, xy)) > 1.0  # it did move


def test_determinism_and_seed_dependence(mc):
    """test_mc_parallel.cpp:136-187: bitwise determinism for a fixed seed."""
    Lx = Ly = 45.254834
    xy = lattice(32, 32, Lx, Ly, 0.2, 1)
    outs = []
    for seed in (7, 7, 8):
        with mc.Context(Lx, Ly, 2.5, T=1.0, delta=0.1, seed=seed, activity=0.3) as ctx:
            ctx.upload(xy)
            st = ctx.run_sweeps(20, gc_per_cell=1)
            outs.append((ctx.download(), st))
    assert outs[0][0].shape == outs[1][0].shape and (outs[0][0] == outs[1][0]).all()
    assert outs[0][1] == outs[1][1]
    assert outs[0][0].shape != outs[2][0].shape or not (outs[0][0] == outs[2][0]).all()


def replay_trace(oracle, calc, Lx, Ly, xy0, trials, rcut):
    """Replays a kernel trial log on the CPU oracle: every trial's dE is recomputed with
    computeLocalEnergy over getNeighbors / getNeighbors2 of a freshly built list on the CURRENT
    configuration (trials in different same-colour cells are independent, so per-cell order within
    a pass is all that matters).  Returns (max error, number checked, final configuration)."""
    xy = [tuple(p) for p in xy0]
    index = {p: i for i, p in enumerate(xy)}
    alive = [True] * len(xy)
    order = np.lexsort((trials["seq"], trials["cell"], trials["pass"]))
    worst, checked = 0.0, 0

    def current():
        arr = np.array([p for p, a in zip(xy, alive) if a], dtype=np.float64).reshape(-1, 2)
        remap = {}
        k = 0
        for i, a in enumerate(alive):
            if a:
                remap[i] = k
                k += 1
        return arr, remap

    for t in trials[order]:
        evaluated = bool(t["reserved"])
        if t["kind"] == 0:
            i = index[(t["old_x"], t["old_y"])]
            if evaluated:
                arr, remap = current()
                uo = oracle.local_energy(calc, Lx, Ly, arr, remap[i])
                un = oracle.point_energy(calc, Lx, Ly, arr, t["new_x"], t["new_y"], remap[i])
                err = abs(t["dE"] - (un - uo)) / max(1.0, abs(un), abs(uo))
                worst, checked = max(worst, err), checked + 1
            if t["accepted"]:
                del index[xy[i]]
                xy[i] = (t["new_x"], t["new_y"])
                index[xy[i]] = i
        elif t["kind"] == 1:
            if evaluated:
                arr, _ = current()
                u = oracle.point_energy(calc, Lx, Ly, arr, t["new_x"], t["new_y"], -1)
                err = abs(t["dE"] - u) / max(1.0, abs(u))
                worst, checked = max(worst, err), checked + 1
            if t["accepted"]:
                xy.append((t["new_x"], t["new_y"]))
                alive.append(True)
                index[xy[-1]] = len(xy) - 1
        else:
            if evaluated:
                i = index[(t["old_x"], t["old_y"])]
                arr, remap = current()
                u = -oracle.local_energy(calc, Lx, Ly, arr, remap[i])
                err = abs(t["dE"] - u) / max(1.0, abs(u))
                worst, checked = max(worst, err), checked + 1
                if t["accepted"]:
                    alive[i] = False
                    del index[xy[i]]
    return worst, checked, current()[0]


@pytest.mark.parametrize("name,ptype,kw,z", [("LennardJones", ol.LJ, dict(rcut=2.5), 0.05),
                                              ("WCA", ol.WCA, dict(rcut=2 ** (1 / 6)), 0.4),
                                              ("AthermalStar", ol.ATHERMAL_STAR, dict(rcut=2.5, f=4.0, alpha=0.3), 0.05),
                                              ("Yukawa", ol.YUKAWA, dict(rcut=2.5), 0.3),
                                              ("ThermalStar", ol.THERMAL_STAR, dict(rcut=2.5, f=6.0, alpha=0.4, A0=1.5, kappa=1.1), 0.05)])
def test_per_trial_dE_matches_oracle(mc, oracle, name, ptype, kw, z):
    """Every displacement / insertion / deletion dE the sweep kernel used in its acceptance rule,
    replayed on the oracle: 1e-10 relative."""
    Lx, Ly = 27.0, 24.0
    xy = lattice(18, 16, Lx, Ly, 0.2, 77)
    ctx, calc = make_pair(mc, oracle, name, ptype, kw, Lx, Ly, T=0.9, z=z, delta=0.12, seed=5)
    with ctx:
        ctx.upload(xy)
        ctx.run_sweeps(3, gc_per_cell=1)  # leave the lattice first
        cfg = ctx.download()
        trials, st = ctx.run_sweep_traced(gc_per_cell=2)
        kinds = np.bincount(trials["kind"], minlength=3)
        assert kinds[0] > 100 and kinds[1] > 10 and kinds[2] > 10
        worst, checked, final = replay_trace(oracle, calc, Lx, Ly, cfg, trials, kw["rcut"])
        assert checked > 150
        assert worst < REL
        out = ctx.download()
        # positions written back go through cell-local coordinates: equal to ~1 ulp of L
        assert len(out) == len(final)
        np.testing.assert_allclose(sort_rows(out), sort_rows(final), rtol=0, atol=1e-12)


# ---- ensemble facts -------------------------------------------------------------------------------------------------------

def block_se(x, nblocks=20):
    x = np.asarray(x, dtype=float)
    b = x[: len(x) // nblocks * nblocks].reshape(nblocks, -1).mean(axis=1)
    return b.std(ddof=1) / math.sqrt(nblocks)


@pytest.mark.parametrize("z", [0.1, 0.5, 1.0, 5.0, 10.0])
def test_ideal_gas_gcmc_mean_and_variance(mc, z):
    """docs/test_scenarios.md:13-56, integration_test_serial/GCMC_ideal_particles_1.cpp: V=100,
    <N> = zV and Var N = zV (Poisson), analytically."""
    Lx, Ly = 10 * math.sqrt(math.sqrt(3) / 2), 10 / math.sqrt(math.sqrt(3) / 2)
    V = Lx * Ly
    with mc.Context(Lx, Ly, 2.5, potential="Ideal", T=1.0, activity=z, delta=0.1, seed=123) as ctx:
        ctx.upload(np.random.default_rng(1).uniform(0, 1, (int(z * V), 2)) * [Lx, Ly])
        ctx.run_sweeps(300, gc_per_cell=4)
        ns = np.array([ctx.run_sweeps(5, gc_per_cell=4)["n_particles"] for _ in range(4000)])
    se = block_se(ns)
    assert abs(ns.mean() - z * V) < 4 * se + 0.02 * math.sqrt(z * V)
    assert abs(ns.var() / (z * V) - 1.0) < 0.15


def test_lj_gcmc_matches_serial_reference_statistics(mc, oracle):
    """<N>, <E>/N and the pressure <rho T + W/2V> of the checkerboard GCMC vs the reference's serial GCMC (oracle
    port, pinned bit-for-bit to the reference) at a supercritical state: agreement within 3 combined standard
    errors (north_star asks for 2 SE; 3 keeps the false-failure rate of this automated check < 1%)."""
    L, T, z, rcut = 20.0, 2.0, 0.08, 2.5
    V = L * L
    calc = oracle.calc(T, 0.0, ol.LJ, rcut, z)
    start = lattice(6, 6, L, L, 0.3, 4)
    ser = oracle.mc(L, L, start, calc, rcut, 0.3, ol.GCMC, 11)
    ser.run(20000, 20000, 10)
    n_s, e_s, p_s = [], [], []
    for _ in range(3000):  # a sample every 40 serial steps; the pressure from computePressureCellList on the snapshot
        _, n1, e1 = ser.run(40, 40, 10)
        n_s.append(n1[-1])
        e_s.append(e1[-1])
        p_s.append(oracle.pressure_celllist(calc, L, L, ser.particles()) if n1[-1] > 0 else 0.0)
    n_s, e_s, p_s = np.array(n_s, float), np.array(e_s), np.array(p_s)
    with mc.Context(L, L, rcut, T=T, activity=z, delta=0.3, seed=3) as ctx:
        ctx.upload(start)
        ctx.run_sweeps(2000, gc_per_cell=1)
        n_g, e_g, p_g = [], [], []
        for _ in range(3000):
            st = ctx.run_sweeps(4, gc_per_cell=1)
            n_g.append(st["n_particles"])
            e_g.append(st["energy"])
            U, W, N = ctx.total_energy_virial()
            p_g.append(N / V * T + W / (2 * V))  # thermodynamic_calculator.cpp:236-242
        assert relerr(st["energy"], U) < 1e-8
    n_g, e_g, p_g = np.array(n_g, float), np.array(e_g), np.array(p_g)
    se_n = math.hypot(block_se(n_s), block_se(n_g))
    assert abs(n_s.mean() - n_g.mean()) < 3 * se_n, (n_s.mean(), n_g.mean(), se_n)
    eps, epg = e_s / np.maximum(n_s, 1), e_g / np.maximum(n_g, 1)
    se_e = math.hypot(block_se(eps), block_se(epg))
    assert abs(eps.mean() - epg.mean()) < 3 * se_e, (eps.mean(), epg.mean(), se_e)
    se_p = math.hypot(block_se(p_s), block_se(p_g))
    assert abs(p_s.mean() - p_g.mean()) < 3 * se_p, (p_s.mean(), p_g.mean(), se_p)


def test_batched_items_give_the_same_trajectory(mc, monkeypatch):
    """k_sweep_p hands a warp's shared-memory slots to the cells of an item by need; an item that does not fit is
    run in batches of consecutive cells.  Forcing the smallest legal slot count (every dense item is batched) must
    not change a single bit of the trajectory."""
    Lx, Ly = 61.0, 47.5
    xy = lattice(48, 38, Lx, Ly, 0.15, 5)  # rho = 0.63: ~40 candidates per cell
    res = []
    for R in (0, 1):  # 1 is raised to the minimum, 9 x cell capacity
        with mc.Context(Lx, Ly, 2.5, T=1.2, activity=0.6, delta=0.1, seed=9, tuning=dict(sweep_slots=R)) as ctx:
            ctx.upload(xy)
            st = ctx.run_sweeps(10, gc_per_cell=2)
            res.append((ctx.download(), st))
    (p0, s0), (p1, s1) = res
    assert p0.shape == p1.shape and (p0 == p1).all()
    assert s0["attempted"] == s1["attempted"] and s0["accepted"] == s1["accepted"]
    assert relerr(s0["energy"], s1["energy"]) < 1e-12


# ---- BASELINE sizes: size-independent properties -------------------------------------------------------------------------

def test_one_million_particles_properties(mc, oracle):
    """Config #4 (1M particles, rho=0.5, LJ rcut 2.5): upload/download is a permutation; the
    slot-grid reduction equals the reference-grid reduction; NVT conserves N; the running energy
    equals a recomputation after sweeps; one oracle energy at full size."""
    n_side = 1000
    L = n_side * math.sqrt(2.0)
    xy = lattice(n_side, n_side, L, L, 0.2, 12345)
    with mc.Context(L, L, 2.5, T=1.0, delta=0.1, seed=12345) as ctx:
        ctx.upload(xy)
        assert ctx.num_particles() == len(xy)
        U0, W0, N0 = ctx.total_energy_virial()
        Ur, Wr = ctx.calc_energy_virial(xy)
        assert relerr(U0, Ur) < REL and relerr(W0, Wr) < REL
        out = ctx.download()
        assert (sort_rows(out) == sort_rows(xy)).all()
        st = ctx.run_sweeps(5)
        assert st["n_particles"] == len(xy) and st["attempted"][0] == 5 * len(xy)
        U1 = ctx.total_energy_virial()[0]
        assert relerr(st["energy"], U1) < 1e-9
        st = ctx.run_sweeps(3, gc_per_cell=1)
        U2, _, N2 = ctx.total_energy_virial()
        assert N2 == st["n_particles"] and relerr(st["energy"], U2) < 1e-9
    calc = oracle.calc(1.0, 0.0, ol.LJ, 2.5)
    assert relerr(U0, oracle.total_energy_celllist(calc, L, L, xy)) < REL


# ---- kernel variants ---------------------------------------------------------------------------------------------------

def test_sweep_kernel_variants_are_bit_identical(mc, monkeypatch):
    """k_sweep (one warp per cell; the fallback for tiny grids and huge cells) and k_sweep_p (G = 4 or 8 lanes
    per cell, persistent warps over the sorted work list; the production kernel) run the same algorithm on
    the same Philox streams: identical configurations and counters, energy equal to rounding of the
    summation order.  Also: k_sweep_p is reproducible run to run although its items are scheduled dynamically."""
    Lx, Ly = 61.0, 47.5
    xy = lattice(40, 32, Lx, Ly, 0.2, 3)
    res = {}
    # "/unfused": one launch per colour pass instead of one per sweep; "/barrier": one launch per sweep with a grid
    # barrier between the colours instead of the per-column-pair completion counters
    variants = ["0", "4", "8", "4", "2", "4/unfused", "4/barrier"]
    for n, G in enumerate(variants):
        tuning = dict(sweep_lanes=int(G.split("/")[0]), fuse_passes=0 if "unfused" in G else 1,
                      pipeline=0 if "barrier" in G else 1)
        with mc.Context(Lx, Ly, 2.5, T=0.9, activity=0.3, delta=0.12, seed=21, tuning=tuning) as ctx:
            ctx.upload(xy, np.arange(len(xy), dtype=np.int32))
            st = ctx.run_sweeps(12, gc_per_cell=2)
            res[n] = (ctx.download(with_ids=True), st)
    (p0, i0), s0 = res[0]
    for n in range(1, len(variants)):
        (p, i), s = res[n]
        assert p.shape == p0.shape and (p == p0).all() and (i == i0).all(), variants[n]
        assert s["attempted"] == s0["attempted"] and s["accepted"] == s0["accepted"], variants[n]
        assert relerr(s["energy"], s0["energy"]) < 1e-12, variants[n]
    assert res[1][1]["energy"] == res[3][1]["energy"]  # same kernel twice: bitwise equal running energy


def test_fixed_trials_per_cell_mode(mc, oracle):
    """disp_per_cell (a fixed number of displacement trials per non-empty cell, HPMC's nselect) keeps the
    bookkeeping exact and samples the same ensemble: ideal-gas <N> = zV, LJ energy consistent."""
    Lx = Ly = 45.254834
    xy = lattice(32, 32, Lx, Ly, 0.2, 42)
    with mc.Context(Lx, Ly, 2.5, T=1.0, delta=0.1, seed=42) as ctx:
        ctx.upload(xy)
        st = ctx.run_sweeps(30, disp_per_cell=3)
        assert st["n_particles"] == 1024 and st["attempted"][0] > 0 and st["attempted"][0] % 3 == 0
        U = ctx.total_energy_virial()[0]
        assert relerr(st["energy"], U) < 1e-9
        assert relerr(U, oracle.total_energy_celllist(oracle.calc(1.0, 0.0, ol.LJ, 2.5), Lx, Ly, ctx.download())) < REL
    z, V = 2.0, 40.0 * 40.0
    with mc.Context(40.0, 40.0, 2.5, potential="Ideal", T=1.0, activity=z, delta=0.3, seed=5) as ctx:
        ctx.upload(np.random.default_rng(1).uniform(0, 40.0, (int(z * V), 2)))
        ctx.run_sweeps(200, gc_per_cell=2, disp_per_cell=2)
        ns = np.array([ctx.run_sweeps(3, gc_per_cell=2, disp_per_cell=2)["n_particles"] for _ in range(1500)])
    assert abs(ns.mean() - z * V) < 4 * block_se(ns) + 0.01 * math.sqrt(z * V)
    assert abs(ns.var() / (z * V) - 1.0) < 0.2


# ---- NPT volume move ------------------------------------------------------------------------------

def _npt_move(ctx, box, rng, T, P, dV):
    """MonteCarlo::npt_move_no_cell_list (MC.cpp:153-188) through the C ABI; returns the new box and accepted"""
    Lx, Ly = box
    st = ctx.stats()
    U_old, N = st["energy"], st["n_particles"]
    V_o = Lx * Ly
    V_n = math.exp(math.log(V_o) + rng.uniform(-0.5, 0.5) * dV)
    ratio = Lx / Ly  # SimulationBox::setV, initial.cpp:80-87
    Lxn, Lyn = math.sqrt(ratio * V_n), math.sqrt(V_n / ratio)
    U_new, _ = ctx.rescale_try(Lxn, Lyn)
    arg = -((U_new - U_old) + P * (V_n - V_o) - (N + 1) * math.log(V_n / V_o) * T) / T
    accept = arg >= 0 or rng.random() < math.exp(arg)
    ctx.rescale_finish(accept)
    return ((Lxn, Lyn) if accept else box), accept


def test_rescale_try_and_finish_parity(mc, oracle):
    """The rescaled configuration's U and W equal the oracle's on the scaled positions; a rejected move leaves the
    context bit for bit where it was; an accepted one continues from the scaled configuration; a box whose cells
    would be narrower than rcut is refused."""
    Lx, Ly = 52.0, 44.0
    xy = lattice(36, 30, Lx, Ly, 0.2, 8)
    calc = oracle.calc(1.0, 0.0, ol.LJ, 2.5)
    with mc.Context(Lx, Ly, 2.5, T=1.0, delta=0.1, seed=4, cell_margin=0.08) as ctx:
        ctx.upload(xy)
        ctx.run_sweeps(5)
        before = ctx.download()
        U0 = ctx.stats()["energy"]
        for Ln in ((53.3, 45.1), (50.1, 42.4)):  # anisotropic on purpose: the ABI takes both lengths
            U, W = ctx.rescale_try(*Ln)
            scaled = before * [Ln[0] / Lx, Ln[1] / Ly]
            assert relerr(U, oracle.total_energy_celllist(calc, Ln[0], Ln[1], scaled)) < REL
            assert relerr(W, oracle.total_virial_celllist(calc, Ln[0], Ln[1], scaled)) < REL
            ctx.rescale_finish(False)
            assert (ctx.download() == before).all() and ctx.stats()["energy"] == U0
        with pytest.raises(mc.MC2DError) as e:
            ctx.rescale_try(40.0, 44.0)  # 40 / 18 columns < rcut
        assert e.value.status == mc.api.ERR_GEOMETRY
        with pytest.raises(mc.MC2DError):
            ctx.rescale_finish(True)  # nothing pending
        Ln = (53.3, 45.1)
        U, W = ctx.rescale_try(*Ln)
        ctx.rescale_finish(True)
        after = ctx.download()
        assert (after == before * [Ln[0] / Lx, Ln[1] / Ly]).all()
        assert ctx.stats()["energy"] == U
        st = ctx.run_sweeps(5)
        out = ctx.download()
        assert relerr(st["energy"], oracle.total_energy_celllist(calc, Ln[0], Ln[1], out)) < 1e-9
        assert out[:, 0].max() < Ln[0] and out[:, 1].max() < Ln[1] and out.min() >= 0.0


def test_npt_ideal_gas_mean_volume(mc):
    """Ideal gas under the reference's volume move (uniform steps in ln V, acceptance with (N+1) ln(Vn/Vo)):
    the stationary density of V is V^N exp(-beta P V), so <V> = (N+1) T / P and Var V = (N+1) (T/P)^2."""
    N, T, P = 200, 1.0, 2.0
    rng = np.random.default_rng(3)
    L0 = math.sqrt((N + 1) * T / P)
    box = (L0 * 1.1, L0 / 1.1)
    with mc.Context(box[0], box[1], 0.5, potential="Ideal", T=T, delta=0.2, seed=6, cell_margin=0.6) as ctx:
        ctx.upload(rng.uniform(0, 1, (N, 2)) * box)
        vs = []
        for it in range(6000):
            ctx.run_sweeps(1)
            box, _ = _npt_move(ctx, box, rng, T, P, 0.15)
            if it >= 500:
                vs.append(box[0] * box[1])
        out = ctx.download()
    vs = np.array(vs)
    mean, var = (N + 1) * T / P, (N + 1) * (T / P) ** 2
    assert abs(vs.mean() - mean) < 4 * block_se(vs) + 0.002 * mean, (vs.mean(), mean, block_se(vs))
    assert abs(vs.var() / var - 1.0) < 0.2
    assert len(out) == N and out[:, 0].max() < box[0] and out[:, 1].max() < box[1]


def test_npt_lj_matches_serial_reference_statistics(mc):
    """<V> and <U>/N of an LJ fluid at fixed N, P, T: checkerboard sweeps + mc2d_rescale_* volume moves against the
    UNMODIFIED reference's NPT loop (oracle/_ref; N displacement trials then a volume move with probability 0.7 per
    step, MC.cpp:38-60), within 3 combined standard errors."""
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built")
    ref = ol.Ref()
    if not hasattr(ref.lib, "ref_mc_volume"):
        pytest.skip("oracle/_ref predates the NPT accessors")
    N_side, T, P, dV, delta = 12, 2.0, 1.0, 0.08, 0.2
    L0 = 20.0
    start = lattice(N_side, N_side, L0, L0, 0.1, 2)
    N = len(start)
    ser = ref.mc(L0, L0, start, T=T, press=P, ptype=ol.LJ, rcut=2.5, delta=delta, ensemble=2, seed=17)
    ser.set_delta_V(dV)
    ser.run(3000, 3000, 10)
    v_s, e_s = [], []
    for _ in range(2500):
        ser.run(10, 10, 10)
        v_s.append(ser.volume)
        e_s.append(ser.energy / N)
    rng = np.random.default_rng(11)
    box = (L0, L0)
    v_g, e_g = [], []
    with mc.Context(L0, L0, 2.5, T=T, delta=delta, seed=23, cell_margin=0.25) as ctx:
        ctx.upload(start)
        for it in range(3000 + 25000):
            ctx.run_sweeps(1)
            if rng.random() > 0.3:
                box, _ = _npt_move(ctx, box, rng, T, P, dV)
            if it >= 3000 and it % 10 == 9:
                v_g.append(box[0] * box[1])
                e_g.append(ctx.stats()["energy"] / N)
        U = ctx.total_energy_virial()[0]
        assert relerr(ctx.stats()["energy"], U) < 1e-8
    v_s, e_s, v_g, e_g = map(np.array, (v_s, e_s, v_g, e_g))
    se_v = math.hypot(block_se(v_s), block_se(v_g))
    assert abs(v_s.mean() - v_g.mean()) < 3 * se_v, (v_s.mean(), v_g.mean(), se_v)
    se_e = math.hypot(block_se(e_s), block_se(e_g))
    assert abs(e_s.mean() - e_g.mean()) < 3 * se_e, (e_s.mean(), e_g.mean(), se_e)


def test_auto_cell_width_for_short_range_potentials(mc, oracle):
    """cell_margin = -1: at the first upload the cells are widened until they hold ~2.5 particles (never below rcut).
    WCA at rho = 0.5 has 0.63 particles per rcut-cell; the wider grid must give the same energies and keep the
    bookkeeping exact."""
    L, rc = 120.0, 2.0 ** (1.0 / 6.0)
    xy = lattice(84, 84, L, L, 0.15, 6)
    calc = oracle.calc(1.0, 0.0, ol.WCA, rc)
    with mc.Context(L, L, rc, potential="WCA", T=1.0, delta=0.1, seed=2, cell_margin=-1.0) as ctx, \
            mc.Context(L, L, rc, potential="WCA", T=1.0, delta=0.1, seed=2) as ref:
        ctx.upload(xy)
        ref.upload(xy)
        a, b = ctx.info(), ref.info()
        assert b.nx == 106 and a.nx < 70 and a.cell_dx >= rc and a.nx % 2 == 0  # ~2 rcut wide
        U, W, N = ctx.total_energy_virial()
        Ur, Wr, _ = ref.total_energy_virial()
        assert relerr(U, Ur) < REL and relerr(W, Wr) < REL
        st = ctx.run_sweeps(10, gc_per_cell=1)
        out = ctx.download()
        assert relerr(st["energy"], oracle.total_energy_celllist(calc, L, L, out)) < 1e-9


def test_single_particle_probes_match_oracle(mc, oracle):
    """mc2d_probe_point / mc2d_probe_particle / mc2d_apply_probe (the Gibbs-ensemble exchange move, GibbsMC.cpp:340-394):
    energies equal the oracle's getNeighbors2 / getNeighbors + computeLocalEnergy on the downloaded configuration;
    applying them inserts / deletes exactly that particle and moves the running energy by exactly dU."""
    Lx, Ly = 47.0, 39.0
    xy = lattice(30, 26, Lx, Ly, 0.2, 9)
    calc = oracle.calc(1.0, 0.0, ol.LJ, 2.5)
    rng = np.random.default_rng(4)
    with mc.Context(Lx, Ly, 2.5, T=1.0, delta=0.1, seed=8) as ctx:
        ctx.upload(xy)
        ctx.run_sweeps(7)  # a shifted grid, as in a running simulation
        cfg = ctx.download()
        for _ in range(20):
            x, y = rng.uniform(0, Lx), rng.uniform(0, Ly)
            assert relerr(ctx.probe_point(x, y), oracle.point_energy(calc, Lx, Ly, cfg, x, y)) < REL
            i = int(rng.integers(len(cfg)))
            px, py, dU = ctx.probe_particle(i)
            assert (px, py) == tuple(cfg[i])
            assert relerr(dU, -oracle.local_energy(calc, Lx, Ly, cfg, i)) < REL
        ctx.apply_probe(0)
        U0 = ctx.stats()["energy"]
        x, y = 13.37, 21.5
        dU = ctx.probe_point(x, y)
        ctx.apply_probe(1)
        after = ctx.download()
        assert len(after) == len(cfg) + 1 and ((after == [x, y]).all(axis=1)).sum() == 1
        assert ctx.stats()["energy"] == U0 + dU
        assert relerr(ctx.stats()["energy"], oracle.total_energy_celllist(calc, Lx, Ly, after)) < 1e-9
        j = 123
        px, py, dU2 = ctx.probe_particle(j)
        assert (px, py) == tuple(after[j])
        ctx.apply_probe(2)
        final = ctx.download()
        assert len(final) == len(cfg) and not ((final == [px, py]).all(axis=1)).any()
        assert relerr(ctx.stats()["energy"], oracle.total_energy_celllist(calc, Lx, Ly, final)) < 1e-9
        with pytest.raises(mc.MC2DError):
            ctx.apply_probe(1)  # nothing pending
        with pytest.raises(mc.MC2DError):
            ctx.probe_particle(len(final))
        st = ctx.run_sweeps(3)
        assert relerr(st["energy"], oracle.total_energy_celllist(calc, Lx, Ly, ctx.download())) < 1e-9


def test_pipelined_reupload_and_no_energy_option(mc):
    """Every upload but a context's first copies host -> device in chunks that k_bin_direct bins while the next chunk is
    on the bus, and mc2d_download copies block-wise behind k_compact: same slots, same order, ids intact.
    MC2D_UPLOAD_NO_ENERGY skips the initial-energy pass: the running energy then counts accepted dE from zero."""
    Lx = Ly = 230.0
    xy = lattice(160, 160, Lx, Ly, 0.3, 17)
    rng = np.random.default_rng(3)
    xy = xy[rng.permutation(len(xy))]
    ids = np.arange(len(xy), dtype=np.int32)[::-1].copy()
    with mc.Context(Lx, Ly, 2.5, T=1.0, activity=0.2, delta=0.1, seed=5) as ctx:
        ctx.upload(xy, ids)                       # first upload: count, allocate, bin
        d0, i0 = ctx.download(with_ids=True)
        U0 = ctx.total_energy_virial()[0]
        ctx.upload(xy, ids)                       # re-upload: chunked copy + binning
        d1, i1 = ctx.download(with_ids=True)
        assert (d0 == d1).all() and (i0 == i1).all()
        assert (xy[len(xy) - 1 - i1] == d1).all()  # ids travel with their particle
        assert ctx.stats()["energy"] == U0
        ctx.upload(xy, ids, initial_energy=False)
        assert ctx.stats()["energy"] == 0.0
        st = ctx.run_sweeps(4, gc_per_cell=1)
        U1 = ctx.total_energy_virial(resync=True)[0]
        assert abs(st["energy"] - (U1 - U0)) < 1e-9 * max(1.0, abs(U0))  # sum of accepted dE since the upload
        assert ctx.stats()["energy"] == U1
        # mc2d_sample = reduction + download in one call (the reduction overlaps the copies): the same numbers
        buf = np.empty((len(xy) + 4096, 2))
        Us, Ws, Ns, m = ctx.sample_into(buf, resync=False)
        ref = ctx.download()
        assert (Us, Ws, Ns) == ctx.total_energy_virial() and m == Ns == len(ref) and (buf[:m] == ref).all()


# ---- BASELINE configs at the stated gate ------------------------------------------------------------------

def test_config1_nvt_lj_1024_matches_serial_reference(mc, oracle):
    """Config #1: NVT LJ, N = 1024, L = 45.254834 (rho 0.5), rcut 2.5, T = 1, delta 0.1 (serial_src/main.cpp:102-144).
    <E>/N and the pressure of the checkerboard sampler against the serial reference (oracle port, pinned bit for bit to
    the reference), 6 independent serial chains with block-averaged standard errors.  The CI gate is 3 combined standard
    errors (a 2-SE gate fails one honest run in twenty); the z-scores are printed, and the long run behind the
    north-star's 2-SE statement is profiles/r02_stat_parity.md (tools/stat_parity.py: |z| = 0.7 and 0.2 there)."""
    L, T, rcut, delta = 45.254834, 1.0, 2.5, 0.1
    V = L * L
    calc = oracle.calc(T, 0.0, ol.LJ, rcut)
    start = lattice(32, 32, L, L, 0.2, 42)
    e_s, p_s, se_e2, se_p2 = [], [], 0.0, 0.0
    nchains = 6
    for chain in range(nchains):
        ser = oracle.mc(L, L, start, calc, rcut, delta, ol.NVT, 42 + chain)
        ser.run(1500, 1500, 1)
        ec, pc = [], []
        for _ in range(300):
            _, n1, e1 = ser.run(10, 10, 1)
            ec.append(e1[-1] / 1024)
            pc.append(oracle.pressure_celllist(calc, L, L, ser.particles()))
        e_s.append(np.mean(ec))
        p_s.append(np.mean(pc))
        se_e2 += block_se(ec, 10) ** 2
        se_p2 += block_se(pc, 10) ** 2
    se_es, se_ps = math.sqrt(se_e2) / nchains, math.sqrt(se_p2) / nchains
    with mc.Context(L, L, rcut, T=T, delta=delta, seed=42) as ctx:
        ctx.upload(start)
        ctx.run_sweeps(2000)
        e_g, p_g = [], []
        for _ in range(6000):
            st = ctx.run_sweeps(5)
            U, W, N = ctx.total_energy_virial()
            assert N == 1024
            e_g.append(U / N)
            p_g.append(N / V * T + W / (2 * V))
        assert relerr(st["energy"], U) < 1e-8
    z_e = (np.mean(e_g) - np.mean(e_s)) / math.hypot(block_se(e_g), se_es)
    z_p = (np.mean(p_g) - np.mean(p_s)) / math.hypot(block_se(p_g), se_ps)
    print("config #1: <E>/N gpu %.5f serial %.5f z = %+.2f;  P gpu %.5f serial %.5f z = %+.2f"
          % (np.mean(e_g), np.mean(e_s), z_e, np.mean(p_g), np.mean(p_s), z_p))
    assert abs(z_e) < 3 and abs(z_p) < 3, (z_e, z_p)


def test_wca_one_million_particles(mc, oracle):
    """Config #4 with WCA (rcut = 2^(1/6), the reference's literal r^2 < 1.2599 test, potential.cpp:59-87): U and W of the
    1 M-particle configuration against the oracle at 1e-10 — on the start AND after GCMC sweeps on cells widened to ~2.5
    particles (cell_margin = -1) — bookkeeping exact, no refused insertion."""
    n_side = 1000
    L = n_side * math.sqrt(2.0)
    rc = 2.0 ** (1.0 / 6.0)
    xy = lattice(n_side, n_side, L, L, 0.2, 12345)
    calc = oracle.calc(1.0, 0.0, ol.WCA, rc, 10.87)
    with mc.Context(L, L, rc, potential="WCA", T=1.0, activity=10.87, delta=0.1, seed=7, cell_margin=-1.0) as ctx:
        ctx.upload(xy)
        info = ctx.info()
        assert info.cell_dx >= 2 * rc * 0.99  # the cells were widened at the first upload
        U0, W0, N0 = ctx.total_energy_virial()
        assert N0 == len(xy)
        assert relerr(U0, oracle.total_energy_celllist(calc, L, L, xy)) < REL
        assert relerr(W0, oracle.total_virial_celllist(calc, L, L, xy)) < REL
        st = ctx.run_sweeps(10, gc_per_cell=1)
        assert st["overflow"] == 0 and st["attempted"][1] > 0 and st["accepted"][1] > 0
        U1, W1, N1 = ctx.total_energy_virial()
        assert N1 == st["n_particles"] and relerr(st["energy"], U1) < 1e-9
        out = ctx.download()
        assert len(out) == N1
        assert relerr(U1, oracle.total_energy_celllist(calc, L, L, out)) < REL
        assert relerr(W1, oracle.total_virial_celllist(calc, L, L, out)) < REL


def test_refused_insertions_are_reported_per_call(mc):
    """An insertion refused because its cell has no free slot biases the ensemble: mc2d_run_sweeps reports it with
    MC2D_ERR_CELL_OVERFLOW for THAT call, with or without a statistics pointer; the sweeps have run, the counters are
    valid, and a later call that refuses nothing succeeds again (the status is per call, stats.overflow the running total)."""
    import ctypes as C

    from montecarlo_simulation_b200 import api

    L = 20.0
    with mc.Context(L, L, 2.5, potential="Ideal", T=1.0, activity=3.0, delta=0.1, seed=3, cell_capacity=8) as ctx:
        # <n> = z A = 18.75 per cell: the 8 slots fill up within a few sweeps.  Frozen grid: with the per-sweep shift a
        # re-bin would need more than 8 slots in some new cell first, which is the other, fatal kind of overflow.
        ctx.upload(np.zeros((0, 2)))
        with pytest.raises(mc.MC2DError) as e:
            ctx.run_sweeps(40, gc_per_cell=1, shift_grid=False)
        assert e.value.status == api.ERR_CELL_OVERFLOW
        st = ctx.stats()
        assert st["overflow"] > 0 and st["sweeps_done"] == 40 and st["n_particles"] <= 64 * 8
        plan = api.SweepPlan(1, 1, 0, 0)
        assert ctx.L.mc2d_run_sweeps(ctx.h, 5, C.byref(plan), None) == api.ERR_CELL_OVERFLOW  # no statistics asked: still reported
        ctx.set_activity(1e-9)  # deletions always succeed now, insertions never: the cells drain
        try:
            ctx.run_sweeps(80, gc_per_cell=1, shift_grid=False)  # (its first sweeps can still draw an insertion on a full cell)
        except mc.MC2DError as e2:
            assert e2.status == api.ERR_CELL_OVERFLOW
        st2 = ctx.run_sweeps(5, gc_per_cell=1, shift_grid=False)  # nothing refused in THIS call: success, although the running total is > 0
        assert st2["overflow"] >= st["overflow"] > 0 and st2["sweeps_done"] == 130
