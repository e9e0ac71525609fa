"""ctypes bindings for the CPU oracle (oracle/libmc2d_oracle.so) and, when present, the
reference compiled in place (oracle/_ref/libmc2d_ref.so).  TEST INFRASTRUCTURE ONLY — the
product package never imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmc2d_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmc2d_ref.so")

LJ, WCA, YUKAWA, ATHERMAL_STAR, THERMAL_STAR, IDEAL = range(6)
NVT, GCMC = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def _xy(xy):
    a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    return a, a.shape[0]


class OrcCalc(C.Structure):
    _fields_ = [("T", C.c_double), ("press", C.c_double), ("z", C.c_double), ("rcut", C.c_double),
                ("r2cut", C.c_double), ("type", C.c_int), ("f_prime", C.c_double), ("alpha", C.c_double),
                ("f_d_prime", C.c_double), ("kappa", C.c_double)]


def build_oracle(force: bool = False) -> None:
    src = os.path.join(ORACLE_DIR, "mc2d_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle"], stdout=subprocess.DEVNULL)


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        d, i, f = C.c_double, C.c_int, C.c_float
        L.orc_pair_potential.restype = d
        L.orc_pair_potential.argtypes = [d, i, f, f, f, f]
        L.orc_pair_force.restype = d
        L.orc_pair_force.argtypes = [d, i, f, f, f, f]
        L.orc_select_potential.argtypes = [C.c_char_p]
        L.orc_calc_init.argtypes = [C.POINTER(OrcCalc), d, d, i, d, d, d, d, d, d]
        L.orc_apply_pbc.argtypes = [d, d, c_double_p, c_double_p]
        L.orc_min_image_r2.restype = d
        L.orc_min_image_r2.argtypes = [d] * 6
        for name in ("orc_total_energy", "orc_total_virial", "orc_pressure", "orc_total_energy_celllist",
                     "orc_total_virial_celllist", "orc_pressure_celllist"):
            fn = getattr(L, name)
            fn.restype = d
            fn.argtypes = [C.POINTER(OrcCalc), d, d, c_double_p, i]
        L.orc_cell_ids.argtypes = [d, d, d, c_double_p, i, c_int_p]
        L.orc_neighbor_counts.argtypes = [d, d, d, c_double_p, i, c_int_p]
        L.orc_local_energy.restype = d
        L.orc_local_energy.argtypes = [C.POINTER(OrcCalc), d, d, c_double_p, i, i]
        L.orc_point_energy.restype = d
        L.orc_point_energy.argtypes = [C.POINTER(OrcCalc), d, d, c_double_p, i, d, d, i]
        L.orc_cl_create.restype = C.c_void_p
        L.orc_cl_create.argtypes = [d, d, d]
        L.orc_cl_destroy.argtypes = [C.c_void_p]
        L.orc_cl_dims.argtypes = [C.c_void_p, c_int_p, c_int_p, c_double_p, c_double_p]
        L.orc_cl_build.argtypes = [C.c_void_p, c_double_p, i]
        L.orc_cl_neighbors.argtypes = [C.c_void_p, i, c_double_p, c_int_p, c_double_p, i]
        L.orc_cl_neighbors_point.argtypes = [C.c_void_p, d, d, c_double_p, c_int_p, c_double_p, i]
        L.orc_rng_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_rng_uniform01.restype = d
        L.orc_rng_uniform01.argtypes = [C.c_void_p]
        L.orc_rng_uniform.restype = d
        L.orc_rng_uniform.argtypes = [C.c_void_p, d, d]
        L.orc_rng_uniform_int.argtypes = [C.c_void_p, i, i]
        L.orc_initialize_particles_random.argtypes = [d, d, i, C.c_uint, c_double_p]
        L.orc_mc_create.restype = C.c_void_p
        L.orc_mc_create.argtypes = [d, d, c_double_p, i, C.POINTER(OrcCalc), d, d, i, C.c_uint32]
        L.orc_mc_destroy.argtypes = [C.c_void_p]
        L.orc_mc_run.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, c_ll_p, c_int_p, c_double_p, i]
        for name in ("orc_mc_step_celllist", "orc_mc_step_bruteforce", "orc_mc_gc_move", "orc_mc_n"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_mc_update_celllist.argtypes = [C.c_void_p]
        L.orc_mc_energy.restype = d
        L.orc_mc_energy.argtypes = [C.c_void_p]
        L.orc_mc_get_particles.argtypes = [C.c_void_p, c_double_p]
        L.orc_mc_accepted.restype = C.c_longlong
        L.orc_mc_accepted.argtypes = [C.c_void_p]
        L.orc_mc_total.restype = C.c_longlong
        L.orc_mc_total.argtypes = [C.c_void_p]
        L.orc_sanity_frame.argtypes = [d, d, i, i, i, C.c_uint, c_double_p]
        L.orc_best_decomposition.argtypes = [d, d, i, c_int_p, c_int_p]

    # -- pair functions -----------------------------------------------------------------
    def pair_potential(self, r2, ptype, fp=0.0, fdp=0.0, kappa=0.0, alpha=0.0):
        return self.lib.orc_pair_potential(r2, ptype, fp, fdp, kappa, alpha)

    def pair_force(self, r2, ptype, fp=0.0, fdp=0.0, kappa=0.0, alpha=0.0):
        return self.lib.orc_pair_force(r2, ptype, fp, fdp, kappa, alpha)

    def select_potential(self, name: str) -> int:
        return self.lib.orc_select_potential(name.encode())

    def calc(self, T=1.0, press=0.0, ptype=LJ, rcut=2.5, mu=0.0, f=0.0, alpha=0.0, A0=0.0, kappa=0.0) -> OrcCalc:
        c = OrcCalc()
        self.lib.orc_calc_init(C.byref(c), T, press, ptype, rcut, mu, f, alpha, A0, kappa)
        return c

    # -- observables ----------------------------------------------------------------------
    def _obs(self, name, calc, Lx, Ly, xy):
        a, n = _xy(xy)
        return getattr(self.lib, name)(C.byref(calc), Lx, Ly, _dp(a), n)

    def total_energy(self, calc, Lx, Ly, xy):
        return self._obs("orc_total_energy", calc, Lx, Ly, xy)

    def total_virial(self, calc, Lx, Ly, xy):
        return self._obs("orc_total_virial", calc, Lx, Ly, xy)

    def pressure(self, calc, Lx, Ly, xy):
        return self._obs("orc_pressure", calc, Lx, Ly, xy)

    def total_energy_celllist(self, calc, Lx, Ly, xy):
        return self._obs("orc_total_energy_celllist", calc, Lx, Ly, xy)

    def total_virial_celllist(self, calc, Lx, Ly, xy):
        return self._obs("orc_total_virial_celllist", calc, Lx, Ly, xy)

    def pressure_celllist(self, calc, Lx, Ly, xy):
        return self._obs("orc_pressure_celllist", calc, Lx, Ly, xy)

    def cell_ids(self, Lx, Ly, rcut, xy):
        a, n = _xy(xy)
        out = np.empty(n, dtype=np.int32)
        self.lib.orc_cell_ids(Lx, Ly, rcut, _dp(a), n, _ip(out))
        return out

    def cell_dims(self, Lx, Ly, rcut):
        h = self.lib.orc_cl_create(Lx, Ly, rcut)
        nx, ny, dx, dy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        self.lib.orc_cl_dims(h, C.byref(nx), C.byref(ny), C.byref(dx), C.byref(dy))
        self.lib.orc_cl_destroy(h)
        return nx.value, ny.value, dx.value, dy.value

    def neighbor_counts(self, Lx, Ly, rcut, xy):
        a, n = _xy(xy)
        out = np.empty(n, dtype=np.int32)
        self.lib.orc_neighbor_counts(Lx, Ly, rcut, _dp(a), n, _ip(out))
        return out

    def neighbors(self, Lx, Ly, rcut, xy, idx, cap=4096):
        a, n = _xy(xy)
        h = self.lib.orc_cl_create(Lx, Ly, rcut)
        self.lib.orc_cl_build(h, _dp(a), n)
        j = np.empty(cap, dtype=np.int32)
        r2 = np.empty(cap, dtype=np.float64)
        m = self.lib.orc_cl_neighbors(h, idx, _dp(a), _ip(j), _dp(r2), cap)
        self.lib.orc_cl_destroy(h)
        return j[:m].copy(), r2[:m].copy()

    def neighbors_point(self, Lx, Ly, rcut, xy, x, y, cap=4096):
        a, n = _xy(xy)
        h = self.lib.orc_cl_create(Lx, Ly, rcut)
        self.lib.orc_cl_build(h, _dp(a), n)
        j = np.empty(cap, dtype=np.int32)
        r2 = np.empty(cap, dtype=np.float64)
        m = self.lib.orc_cl_neighbors_point(h, x, y, _dp(a), _ip(j), _dp(r2), cap)
        self.lib.orc_cl_destroy(h)
        return j[:m].copy(), r2[:m].copy()

    def local_energy(self, calc, Lx, Ly, xy, idx):
        a, n = _xy(xy)
        return self.lib.orc_local_energy(C.byref(calc), Lx, Ly, _dp(a), n, idx)

    def point_energy(self, calc, Lx, Ly, xy, x, y, exclude=-1):
        a, n = _xy(xy)
        return self.lib.orc_point_energy(C.byref(calc), Lx, Ly, _dp(a), n, x, y, exclude)

    def local_energies(self, calc, Lx, Ly, xy, idxs):
        """local energy of several particles over ONE freshly built list (cheap for big N)."""
        a, n = _xy(xy)
        return np.array([self.lib.orc_local_energy(C.byref(calc), Lx, Ly, _dp(a), n, int(k)) for k in idxs])

    # -- rng / init -------------------------------------------------------------------------
    def rng(self, seed):
        buf = C.create_string_buffer(4 * 624 + 16)
        self.lib.orc_rng_seed(buf, seed)
        return buf

    def initialize_particles_random(self, Lx, Ly, N, seed):
        xy = np.empty((N, 2), dtype=np.float64)
        rc = self.lib.orc_initialize_particles_random(Lx, Ly, N, seed, _dp(xy))
        assert rc == 0
        return xy

    def sanity_frame(self, Lx, Ly, Px, Py, N, seed):
        xy = np.empty((N, 2), dtype=np.float64)
        n = self.lib.orc_sanity_frame(Lx, Ly, Px, Py, N, seed, _dp(xy))
        assert n == N
        return xy

    def best_decomposition(self, Lx, Ly, nranks):
        px, py = C.c_int(), C.c_int()
        self.lib.orc_best_decomposition(Lx, Ly, nranks, C.byref(px), C.byref(py))
        return px.value, py.value

    # -- serial MC ----------------------------------------------------------------------------
    def mc(self, Lx, Ly, xy, calc, rcut, delta, ensemble, seed):
        return _MC(self.lib, "orc_", Lx, Ly, xy, calc=calc, rcut=rcut, delta=delta, ensemble=ensemble, seed=seed)


class _MC:
    """Serial step loop handle (oracle port or real reference, same surface)."""

    def __init__(self, lib, prefix, Lx, Ly, xy, *, calc=None, rcut=None, delta=0.1, ensemble=NVT, seed=1,
                 ref_args=None):
        self.lib, self.p = lib, prefix
        a, n = _xy(xy)
        if prefix == "orc_":
            self.h = lib.orc_mc_create(Lx, Ly, _dp(a), n, C.byref(calc), rcut, delta, ensemble, seed)
        else:
            self.h = lib.ref_mc_create(Lx, Ly, _dp(a), n, *ref_args, delta, ensemble, seed)
        self.h = C.c_void_p(self.h)

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def run(self, nsteps, foutput, fupdate):
        cap = int(nsteps // foutput + 2)
        t = np.zeros(cap, dtype=np.int64)
        n = np.zeros(cap, dtype=np.int32)
        e = np.zeros(cap, dtype=np.float64)
        ns = self._f("mc_run")(self.h, nsteps, foutput, fupdate, t.ctypes.data_as(c_ll_p), _ip(n), _dp(e), cap)
        assert ns >= 0, "ref run needs fOutputStep % fUpdateCell == 0 and nSteps % fOutputStep == 0"
        return t[:ns], n[:ns], e[:ns]

    def step_celllist(self):
        return self._f("mc_step_celllist")(self.h)

    def step_bruteforce(self):
        return self._f("mc_step_bruteforce")(self.h)

    def gc_move(self):
        return self._f("mc_gc_move")(self.h)

    def update_celllist(self):
        self._f("mc_update_celllist")(self.h)

    @property
    def energy(self):
        return self._f("mc_energy")(self.h)

    @property
    def n(self):
        return self._f("mc_n")(self.h)

    def set_delta_V(self, dv):  # reference handle only (NPT)
        self.lib.ref_mc_set_delta_V(self.h, dv)

    @property
    def volume(self):
        return self.lib.ref_mc_volume(self.h)

    def particles(self):
        xy = np.empty((max(self.n, 1), 2), dtype=np.float64)
        self._f("mc_get_particles")(self.h, _dp(xy))
        return xy[: self.n].copy()

    def close(self):
        if self.h:
            self._f("mc_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class Ref:
    """The UNMODIFIED reference (serial_src) behind oracle/ref_shim.cpp."""

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        d, i, f = C.c_double, C.c_int, C.c_float
        L.ref_pair_potential.restype = d
        L.ref_pair_potential.argtypes = [d, i, f, f, f, f]
        L.ref_pair_force.restype = d
        L.ref_pair_force.argtypes = [d, i, f, f, f, f]
        L.ref_select_potential.argtypes = [C.c_char_p]
        L.ref_apply_pbc.argtypes = [d, d, c_double_p, c_double_p]
        L.ref_min_image_r2.restype = d
        L.ref_min_image_r2.argtypes = [d] * 6
        L.ref_cl_dims.argtypes = [d, d, d, c_int_p, c_int_p, c_double_p, c_double_p]
        L.ref_cell_ids.argtypes = [d, d, d, c_double_p, i, c_int_p]
        L.ref_neighbor_counts.argtypes = [d, d, d, c_double_p, i, c_int_p]
        L.ref_neighbors.argtypes = [d, d, d, c_double_p, i, i, c_int_p, c_double_p, i]
        L.ref_neighbors_point.argtypes = [d, d, d, c_double_p, i, d, d, c_int_p, c_double_p, i]
        L.ref_observable.restype = d
        L.ref_observable.argtypes = [i, d, d, i, d, d, d, d, d, d, d, d, c_double_p, i]
        L.ref_local_energy.restype = d
        L.ref_local_energy.argtypes = [d, i, d, d, d, d, d, d, d, c_double_p, i, i]
        L.ref_point_energy.restype = d
        L.ref_point_energy.argtypes = [d, i, d, d, d, d, d, d, d, c_double_p, i, d, d, i]
        L.ref_rng_create.restype = C.c_void_p
        L.ref_rng_create.argtypes = [C.c_uint]
        L.ref_rng_destroy.argtypes = [C.c_void_p]
        L.ref_rng_uniform01.restype = d
        L.ref_rng_uniform01.argtypes = [C.c_void_p]
        L.ref_rng_uniform.restype = d
        L.ref_rng_uniform.argtypes = [C.c_void_p, d, d]
        L.ref_rng_uniform_int.argtypes = [C.c_void_p, i, i]
        L.ref_initialize_particles_random.argtypes = [d, d, i, C.c_uint, c_double_p]
        L.ref_mc_create.restype = C.c_void_p
        L.ref_mc_create.argtypes = [d, d, c_double_p, i, d, d, i, d, d, d, d, d, d, d, i, C.c_uint]
        L.ref_mc_destroy.argtypes = [C.c_void_p]
        L.ref_mc_run.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, c_ll_p, c_int_p, c_double_p, i]
        for name in ("ref_mc_step_celllist", "ref_mc_step_bruteforce", "ref_mc_gc_move", "ref_mc_n"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_mc_update_celllist.argtypes = [C.c_void_p]
        L.ref_mc_energy.restype = d
        L.ref_mc_energy.argtypes = [C.c_void_p]
        L.ref_mc_get_particles.argtypes = [C.c_void_p, c_double_p]
        if hasattr(L, "ref_mc_volume"):
            L.ref_mc_set_delta_V.argtypes = [C.c_void_p, d]
            L.ref_mc_volume.restype = d
            L.ref_mc_volume.argtypes = [C.c_void_p]
        if hasattr(L, "ref_mc_run_exact"):
            L.ref_mc_run_exact.argtypes = [C.c_void_p, C.c_size_t, i, c_double_p]
        L.ref_time_mc_run.restype = d
        L.ref_time_mc_run.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, c_ll_p]
        L.ref_time_moves.restype = d
        L.ref_time_moves.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, i, c_ll_p]
        L.ref_time_energy_virial.restype = d
        L.ref_time_energy_virial.argtypes = [d, i, d, d, d, d, d, d, d, c_double_p, i, c_double_p, c_double_p]
        if hasattr(L, "ref_log_positions"):
            L.ref_log_positions.argtypes = [C.c_char_p, C.c_char_p, i, d, d, c_double_p, C.POINTER(C.c_int), i, C.POINTER(C.c_int)]

    def log_positions(self, path_pos, path_data, dump, Lx, Ly, frames, timesteps):
        """the reference's Logging::logPositions_xyz / _dump (logging.cpp:30-79) for a list of (n, 2) frames"""
        flat = np.ascontiguousarray(np.concatenate([np.asarray(f, dtype=np.float64).reshape(-1, 2) for f in frames]))
        counts = np.array([len(f) for f in frames], dtype=np.int32)
        ts = np.array(timesteps, dtype=np.int32)
        rc = self.lib.ref_log_positions(str(path_pos).encode(), str(path_data).encode(), int(dump), Lx, Ly, _dp(flat),
                                        counts.ctypes.data_as(C.POINTER(C.c_int)), len(frames), ts.ctypes.data_as(C.POINTER(C.c_int)))
        assert rc == 0

    def observable(self, what, Lx, Ly, xy, *, T=1.0, press=0.0, ptype=LJ, rcut=2.5, mu=0.0, f=0.0, alpha=0.0,
                   A0=0.0, kappa=0.0):
        a, n = _xy(xy)
        return self.lib.ref_observable(what, T, press, ptype, rcut, mu, f, alpha, A0, kappa, Lx, Ly, _dp(a), n)

    def cell_ids(self, Lx, Ly, rcut, xy):
        a, n = _xy(xy)
        out = np.empty(n, dtype=np.int32)
        self.lib.ref_cell_ids(Lx, Ly, rcut, _dp(a), n, _ip(out))
        return out

    def cell_dims(self, Lx, Ly, rcut):
        nx, ny, dx, dy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        self.lib.ref_cl_dims(Lx, Ly, rcut, C.byref(nx), C.byref(ny), C.byref(dx), C.byref(dy))
        return nx.value, ny.value, dx.value, dy.value

    def neighbor_counts(self, Lx, Ly, rcut, xy):
        a, n = _xy(xy)
        out = np.empty(n, dtype=np.int32)
        self.lib.ref_neighbor_counts(Lx, Ly, rcut, _dp(a), n, _ip(out))
        return out

    def neighbors(self, Lx, Ly, rcut, xy, idx, cap=4096):
        a, n = _xy(xy)
        j = np.empty(cap, dtype=np.int32)
        r2 = np.empty(cap, dtype=np.float64)
        m = self.lib.ref_neighbors(Lx, Ly, rcut, _dp(a), n, idx, _ip(j), _dp(r2), cap)
        return j[:m].copy(), r2[:m].copy()

    def neighbors_point(self, Lx, Ly, rcut, xy, x, y, cap=4096):
        a, n = _xy(xy)
        j = np.empty(cap, dtype=np.int32)
        r2 = np.empty(cap, dtype=np.float64)
        m = self.lib.ref_neighbors_point(Lx, Ly, rcut, _dp(a), n, x, y, _ip(j), _dp(r2), cap)
        return j[:m].copy(), r2[:m].copy()

    def local_energy(self, Lx, Ly, xy, idx, *, T=1.0, ptype=LJ, rcut=2.5, f=0.0, alpha=0.0, A0=0.0, kappa=0.0):
        a, n = _xy(xy)
        return self.lib.ref_local_energy(T, ptype, rcut, f, alpha, A0, kappa, Lx, Ly, _dp(a), n, idx)

    def point_energy(self, Lx, Ly, xy, x, y, exclude=-1, *, T=1.0, ptype=LJ, rcut=2.5, f=0.0, alpha=0.0, A0=0.0,
                     kappa=0.0):
        a, n = _xy(xy)
        return self.lib.ref_point_energy(T, ptype, rcut, f, alpha, A0, kappa, Lx, Ly, _dp(a), n, x, y, exclude)

    def initialize_particles_random(self, Lx, Ly, N, seed):
        xy = np.empty((N, 2), dtype=np.float64)
        assert self.lib.ref_initialize_particles_random(Lx, Ly, N, seed, _dp(xy)) == 0
        return xy

    def mc(self, Lx, Ly, xy, *, T=1.0, press=0.0, ptype=LJ, rcut=2.5, mu=0.0, f=0.0, alpha=0.0, A0=0.0, kappa=0.0,
           delta=0.1, ensemble=NVT, seed=1):
        return _MC(self.lib, "ref_", Lx, Ly, xy, delta=delta, ensemble=ensemble, seed=seed,
                   ref_args=(T, press, ptype, rcut, mu, f, alpha, A0, kappa))

    def run_exact(self, mc: _MC, nsweeps, gcmc):
        """nsweeps of the reference's moves with an exact neighbour search (ref_shim.cpp: ref_mc_run_exact) -> (N, E)"""
        e = C.c_double()
        n = self.lib.ref_mc_run_exact(mc.h, nsweeps, int(gcmc), C.byref(e))
        return n, e.value

    def time_mc_run(self, mc: _MC, nsteps, fupdate):
        moves = C.c_longlong()
        s = self.lib.ref_time_mc_run(mc.h, nsteps, fupdate, C.byref(moves))
        return s, moves.value

    def time_moves(self, mc: _MC, nsweeps, fupdate, gcmc):
        moves = C.c_longlong()
        s = self.lib.ref_time_moves(mc.h, nsweeps, fupdate, int(gcmc), C.byref(moves))
        return s, moves.value

    def time_energy_virial(self, Lx, Ly, xy, *, T=1.0, ptype=LJ, rcut=2.5, f=0.0, alpha=0.0, A0=0.0, kappa=0.0):
        a, n = _xy(xy)
        U, W = C.c_double(), C.c_double()
        s = self.lib.ref_time_energy_virial(T, ptype, rcut, f, alpha, A0, kappa, Lx, Ly, _dp(a), n, C.byref(U),
                                            C.byref(W))
        return s, U.value, W.value
