"""ctypes binding of libmc2d.so (include/mc2d.h) for tests and bench.py.

This is plumbing: the product is the CUDA library and the C++ host shims in host/.  There is no
CPU fallback anywhere in this package: if libmc2d.so is missing the import of the binding raises,
and on a host without a GPU every compute call returns MC2D_ERR_NO_DEVICE (raised as MC2DError).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MC2D_LIB", os.path.join(PKG, "libmc2d.so"))  # MC2D_LIB: tuning builds

LENNARD_JONES, WCA, YUKAWA, ATHERMAL_STAR, THERMAL_STAR, IDEAL = range(6)
POTENTIAL_NAMES = {"LennardJones": 0, "WCA": 1, "Yukawa": 2, "AthermalStar": 3, "ThermalStar": 4, "Ideal": 5}

(OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_CELL_OVERFLOW, ERR_GEOMETRY, ERR_OUT_OF_BOX, ERR_NCCL, ERR_STATE,
 ERR_PEER_TIMEOUT) = range(10)
UNIQUE_ID_BYTES = 128


class MC2DError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        super().__init__("mc2d status %d: %s" % (status, detail))


class Params(C.Structure):
    _fields_ = [("Lx", C.c_double), ("Ly", C.c_double), ("rcut", C.c_double), ("potential", C.c_int),
                ("f_prime", C.c_float), ("f_d_prime", C.c_float), ("kappa", C.c_float), ("alpha", C.c_float),
                ("temperature", C.c_double), ("activity", C.c_double), ("delta", C.c_double), ("seed", C.c_uint64),
                ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("cell_capacity", C.c_int),
                ("p_insert", C.c_double), ("p_delete", C.c_double), ("cell_margin", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("attempted", C.c_longlong * 3), ("accepted", C.c_longlong * 3), ("n_particles", C.c_longlong),
                ("energy", C.c_double), ("overflow", C.c_longlong), ("sweeps_done", C.c_ulonglong)]

    def as_dict(self):
        return {"attempted": list(self.attempted), "accepted": list(self.accepted), "n_particles": self.n_particles,
                "energy": self.energy, "overflow": self.overflow, "sweeps_done": self.sweeps_done}


class SweepPlan(C.Structure):
    _fields_ = [("disp_per_particle", C.c_int), ("gc_per_cell", C.c_int), ("shift_grid", C.c_int),
                ("disp_per_cell", C.c_int)]


class Trial(C.Structure):
    _fields_ = [("pass_", C.c_int), ("cell", C.c_int), ("seq", C.c_int), ("kind", C.c_int), ("accepted", C.c_int),
                ("reserved", C.c_int), ("old_x", C.c_double), ("old_y", C.c_double), ("new_x", C.c_double),
                ("new_y", C.c_double), ("dE", C.c_double)]


TRIAL_DTYPE = np.dtype([("pass", "<i4"), ("cell", "<i4"), ("seq", "<i4"), ("kind", "<i4"), ("accepted", "<i4"),
                        ("reserved", "<i4"), ("old_x", "<f8"), ("old_y", "<f8"), ("new_x", "<f8"), ("new_y", "<f8"),
                        ("dE", "<f8")])
assert TRIAL_DTYPE.itemsize == C.sizeof(Trial)


class Info(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("col0", C.c_int), ("ncol", C.c_int), ("cell_capacity", C.c_int),
                ("cell_dx", C.c_double), ("cell_dy", C.c_double), ("launches", C.c_longlong),
                ("last_sweep_ms", C.c_double), ("last_energy_ms", C.c_double), ("last_pair_tests", C.c_longlong),
                ("last_pair_evals", C.c_longlong), ("k_sweep_ms", C.c_double), ("k_sweep_launches", C.c_longlong),
                ("k_rebin_ms", C.c_double), ("k_rebin_launches", C.c_longlong),
                ("halo_mode", C.c_int), ("sweep_kernel_lanes", C.c_int),
                ("k_halo_ms", C.c_double), ("k_halo_launches", C.c_longlong),
                ("slot_allocs", C.c_longlong)]


class Tuning(C.Structure):
    """mc2d_tuning: A-B switches (every variant gives the same trajectory); the library reads no environment"""
    _fields_ = [("sweep_lanes", C.c_int), ("sweep_slots", C.c_int), ("sweep_warps_per_cta", C.c_int),
                ("fuse_passes", C.c_int), ("pipeline", C.c_int), ("rebin_lanes", C.c_int),
                ("energy_unstaged", C.c_int), ("halo_nccl", C.c_int), ("halo_overlap", C.c_int),
                ("upload_sort", C.c_int), ("spin_timeout_ms", C.c_int), ("test_upload_lag_us", C.c_int),
                ("reserved", C.c_int * 4)]


# every symbol include/mc2d.h declares (tests check the library exports exactly these)
EXPORTS = [
    "mc2d_status_string", "mc2d_last_error", "mc2d_abi_version", "mc2d_device_count", "mc2d_create", "mc2d_destroy",
    "mc2d_reference_grid", "mc2d_sweep_grid", "mc2d_slab_columns", "mc2d_halo_route", "mc2d_sweep_schedule",
    "mc2d_upload", "mc2d_download", "mc2d_num_particles", "mc2d_run_sweeps", "mc2d_get_stats", "mc2d_reset_counters",
    "mc2d_set_delta", "mc2d_set_activity", "mc2d_total_energy_virial", "mc2d_calc_energy_virial",
    "mc2d_calc_cell_ids", "mc2d_calc_neighbor_counts", "mc2d_calc_point_energies", "mc2d_run_sweep_traced",
    "mc2d_comm_unique_id", "mc2d_comm_init", "mc2d_allreduce_stats", "mc2d_get_info", "mc2d_get_stream",
    "mc2d_synchronize", "mc2d_set_kernel_timing", "mc2d_calc_neighbor_list", "mc2d_calc_energy_virial_bruteforce",
    "mc2d_rescale_try", "mc2d_rescale_finish", "mc2d_probe_point", "mc2d_probe_particle", "mc2d_apply_probe", "mc2d_measure_fp64_peak",
    "mc2d_get_tuning", "mc2d_set_tuning", "mc2d_upload_ex", "mc2d_sample",
]

_lib = None


def load_library() -> C.CDLL:
    """Loads libmc2d.so; raises if it was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmc2d.so not built: run `python -m montecarlo_simulation_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    d, i, ll, vp = C.c_double, C.c_int, C.c_longlong, C.c_void_p
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.mc2d_status_string.restype = C.c_char_p
    L.mc2d_status_string.argtypes = [i]
    L.mc2d_last_error.restype = C.c_char_p
    L.mc2d_last_error.argtypes = [vp]
    L.mc2d_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.mc2d_destroy.argtypes = [vp]
    L.mc2d_reference_grid.argtypes = [d, d, d, ip, ip, dp, dp]
    L.mc2d_sweep_grid.argtypes = [d, d, d, i, ip, ip, dp, dp]
    L.mc2d_slab_columns.argtypes = [i, i, i, ip, ip]
    L.mc2d_halo_route.argtypes = [i, i, i, ip, ip, ip]
    L.mc2d_sweep_schedule.argtypes = [C.c_uint64, C.c_uint64, d, d, i, ip, dp, dp]
    L.mc2d_upload.argtypes = [vp, vp, vp, ll]
    L.mc2d_upload_ex.argtypes = [vp, vp, vp, ll, C.c_uint]
    L.mc2d_sample.argtypes = [vp, i, i, dp, dp, C.POINTER(ll), vp, vp, ll, C.POINTER(ll)]
    L.mc2d_download.argtypes = [vp, vp, vp, ll, C.POINTER(ll)]
    L.mc2d_num_particles.argtypes = [vp, C.POINTER(ll)]
    L.mc2d_run_sweeps.argtypes = [vp, ll, C.POINTER(SweepPlan), C.POINTER(Stats)]
    L.mc2d_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mc2d_reset_counters.argtypes = [vp]
    L.mc2d_set_delta.argtypes = [vp, d]
    L.mc2d_set_activity.argtypes = [vp, d]
    L.mc2d_total_energy_virial.argtypes = [vp, i, i, dp, dp, C.POINTER(ll)]
    L.mc2d_calc_energy_virial.argtypes = [vp, vp, ll, dp, dp]
    L.mc2d_calc_cell_ids.argtypes = [vp, vp, ll, vp]
    L.mc2d_calc_neighbor_counts.argtypes = [vp, vp, ll, vp, vp]
    L.mc2d_calc_point_energies.argtypes = [vp, vp, ll, vp, vp, ll, vp, vp]
    L.mc2d_run_sweep_traced.argtypes = [vp, C.POINTER(SweepPlan), vp, ll, C.POINTER(ll), C.POINTER(Stats)]
    L.mc2d_comm_unique_id.argtypes = [vp]
    L.mc2d_comm_init.argtypes = [vp, vp]
    L.mc2d_allreduce_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mc2d_get_info.argtypes = [vp, C.POINTER(Info)]
    L.mc2d_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.mc2d_synchronize.argtypes = [vp]
    L.mc2d_set_kernel_timing.argtypes = [vp, i]
    L.mc2d_calc_neighbor_list.argtypes = [vp, vp, ll, ll, d, d, vp, vp, i, ip]
    L.mc2d_calc_energy_virial_bruteforce.argtypes = [vp, vp, ll, dp, dp]
    L.mc2d_rescale_try.argtypes = [vp, d, d, dp, dp]
    L.mc2d_rescale_finish.argtypes = [vp, i]
    L.mc2d_probe_point.argtypes = [vp, d, d, dp]
    L.mc2d_probe_particle.argtypes = [vp, ll, dp, dp, dp]
    L.mc2d_apply_probe.argtypes = [vp, i]
    L.mc2d_measure_fp64_peak.argtypes = [vp, dp]
    L.mc2d_get_tuning.argtypes = [vp, C.POINTER(Tuning)]
    L.mc2d_set_tuning.argtypes = [vp, C.POINTER(Tuning)]
    _lib = L
    return L


def calc_params(f=0.0, A0=0.0, kappa=0.0):
    """thermodynamic_calculator.cpp:19-21 (f' and f_d' as doubles; narrowed to float by Params)."""
    f_prime = (2.0 + 9.0 * f * f) / 24.0
    with np.errstate(divide="ignore", invalid="ignore"):
        f_d_prime = float(np.float64(A0 * f * f) / np.float64(kappa))
    return f_prime, f_d_prime


def _f32(v):
    with np.errstate(over="ignore", invalid="ignore"):
        return float(np.float32(v))


def reference_grid(Lx, Ly, rcut):
    L = load_library()
    nx, ny, dx, dy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    rc = L.mc2d_reference_grid(Lx, Ly, rcut, C.byref(nx), C.byref(ny), C.byref(dx), C.byref(dy))
    if rc:
        raise MC2DError(rc, L.mc2d_status_string(rc).decode())
    return nx.value, ny.value, dx.value, dy.value


def sweep_grid(Lx, Ly, rcut, nranks=1):
    L = load_library()
    nx, ny, dx, dy = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    rc = L.mc2d_sweep_grid(Lx, Ly, rcut, nranks, C.byref(nx), C.byref(ny), C.byref(dx), C.byref(dy))
    if rc:
        raise MC2DError(rc, L.mc2d_status_string(rc).decode())
    return nx.value, ny.value, dx.value, dy.value


def slab_columns(nx, nranks, rank):
    L = load_library()
    c0, nc = C.c_int(), C.c_int()
    rc = L.mc2d_slab_columns(nx, nranks, rank, C.byref(c0), C.byref(nc))
    if rc:
        raise MC2DError(rc, L.mc2d_status_string(rc).decode())
    return c0.value, nc.value


def halo_route(colour, rank, nranks):
    L = load_library()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    rc = L.mc2d_halo_route(colour, rank, nranks, C.byref(a), C.byref(b), C.byref(c))
    if rc:
        raise MC2DError(rc, L.mc2d_status_string(rc).decode())
    return bool(a.value), b.value, c.value


def sweep_schedule(seed, sweep, dx, dy, shift_grid=True):
    L = load_library()
    order = (C.c_int * 4)()
    sx, sy = C.c_double(), C.c_double()
    L.mc2d_sweep_schedule(seed, sweep, dx, dy, int(shift_grid), order, C.byref(sx), C.byref(sy))
    return list(order), sx.value, sy.value


def device_count() -> int:
    return load_library().mc2d_device_count()


def comm_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    rc = L.mc2d_comm_unique_id(buf)
    if rc:
        raise MC2DError(rc, "mc2d_comm_unique_id")
    return buf.raw


def bootstrap_comm(ctx, dist, src=0):
    """Multi-GPU bootstrap for torch.distributed hosts: rank `src` creates the 128-byte NCCL unique id, the process group
    ships it (MPI_Bcast in the reference's world), every rank joins (mc2d_comm_init)."""
    ids = [comm_unique_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(ids, src=src)
    ctx.comm_init(ids[0])


class Context:
    """One GPU context (mc2d_ctx).  Mirrors what MonteCarlo + CellList + ThermodynamicCalculator hold."""

    def __init__(self, Lx, Ly, rcut, potential=LENNARD_JONES, T=1.0, activity=0.0, delta=0.1, seed=1, f=0.0,
                 alpha=0.0, A0=0.0, kappa=0.0, device=0, rank=0, nranks=1, cell_capacity=0, p_insert=0.0,
                 p_delete=0.0, cell_margin=0.0, tuning=None):
        self.L = load_library()
        if isinstance(potential, str):
            potential = POTENTIAL_NAMES[potential]
        fp, fdp = calc_params(f, A0, kappa)
        self.params = Params(Lx, Ly, rcut, potential, _f32(fp), _f32(fdp), _f32(kappa), _f32(alpha), T, activity,
                             delta, seed, device, rank, nranks, cell_capacity, p_insert, p_delete, cell_margin)
        self.h = C.c_void_p()
        rc = self.L.mc2d_create(C.byref(self.params), C.byref(self.h))
        if rc:
            detail = self.L.mc2d_last_error(self.h).decode() if self.h else ""
            if self.h:
                self.L.mc2d_destroy(self.h)
                self.h = C.c_void_p()
            raise MC2DError(rc, detail)
        if tuning:
            self.set_tuning(**tuning)

    # -- tuning (mc2d_tuning) -------------------------------------------------------------------
    def tuning(self) -> dict:
        t = Tuning()
        self._check(self.L.mc2d_get_tuning(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Tuning._fields_ if n != "reserved"}

    def set_tuning(self, **kw):
        t = Tuning()
        self._check(self.L.mc2d_get_tuning(self.h, C.byref(t)))
        for k, v in kw.items():
            if k == "reserved" or not hasattr(t, k):
                raise KeyError("unknown tuning field %r" % k)
            setattr(t, k, int(v))
        self._check(self.L.mc2d_set_tuning(self.h, C.byref(t)))

    # -- helpers -------------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise MC2DError(rc, self.L.mc2d_last_error(self.h).decode())

    @staticmethod
    def _xy(xy):
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        return a, a.shape[0]

    def close(self):
        if getattr(self, "h", None):
            self.L.mc2d_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- particles ------------------------------------------------------------------------------
    def upload(self, xy, ids=None, initial_energy=True):
        """initial_energy=False: MC2D_UPLOAD_NO_ENERGY (the caller re-syncs the energy after its sweeps)"""
        a, n = self._xy(xy)
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            idp = ids.ctypes.data
        self._check(self.L.mc2d_upload_ex(self.h, a.ctypes.data, idp, n, 0 if initial_energy else 1))

    def num_particles(self) -> int:
        n = C.c_longlong()
        self._check(self.L.mc2d_num_particles(self.h, C.byref(n)))
        return n.value

    def download_into(self, xy: np.ndarray) -> int:
        """downloads into a caller-owned (e.g. pinned) float64 array of shape (cap, 2); returns the count"""
        assert xy.dtype == np.float64 and xy.flags["C_CONTIGUOUS"] and xy.ndim == 2 and xy.shape[1] == 2
        m = C.c_longlong()
        self._check(self.L.mc2d_download(self.h, xy.ctypes.data, None, xy.shape[0], C.byref(m)))
        return m.value

    def sample_into(self, xy: np.ndarray, resync=True, allreduce=False):
        """mc2d_sample: energy/virial reduction + download into a caller-owned (cap, 2) float64 array -> (U, W, N, count)"""
        assert xy.dtype == np.float64 and xy.flags["C_CONTIGUOUS"] and xy.ndim == 2 and xy.shape[1] == 2
        U, W, N, m = C.c_double(), C.c_double(), C.c_longlong(), C.c_longlong()
        self._check(self.L.mc2d_sample(self.h, int(resync), int(allreduce), C.byref(U), C.byref(W), C.byref(N), xy.ctypes.data,
                                       None, xy.shape[0], C.byref(m)))
        return U.value, W.value, N.value, m.value

    def download(self, with_ids=False):
        n = self.num_particles()
        xy = np.empty((max(n, 1), 2), dtype=np.float64)
        ids = np.empty(max(n, 1), dtype=np.int32) if with_ids else None
        m = C.c_longlong()
        self._check(self.L.mc2d_download(self.h, xy.ctypes.data, ids.ctypes.data if with_ids else None, max(n, 1),
                                         C.byref(m)))
        return (xy[: m.value], ids[: m.value]) if with_ids else xy[: m.value]

    # -- step loop --------------------------------------------------------------------------------
    def run_sweeps(self, nsweeps, disp_per_particle=1, gc_per_cell=0, shift_grid=True, disp_per_cell=0) -> dict:
        plan = SweepPlan(disp_per_particle, gc_per_cell, int(shift_grid), disp_per_cell)
        st = Stats()
        self._check(self.L.mc2d_run_sweeps(self.h, nsweeps, C.byref(plan), C.byref(st)))
        return st.as_dict()

    def run_sweep_traced(self, disp_per_particle=1, gc_per_cell=0, shift_grid=True, cap=1 << 20, disp_per_cell=0):
        plan = SweepPlan(disp_per_particle, gc_per_cell, int(shift_grid), disp_per_cell)
        st = Stats()
        buf = np.zeros(cap, dtype=TRIAL_DTYPE)
        n = C.c_longlong()
        self._check(self.L.mc2d_run_sweep_traced(self.h, C.byref(plan), buf.ctypes.data, cap, C.byref(n), C.byref(st)))
        if n.value > cap:
            raise MC2DError(ERR_INVALID, "trace capacity %d too small for %d trials" % (cap, n.value))
        return buf[: n.value], st.as_dict()

    def stats(self) -> dict:
        st = Stats()
        self._check(self.L.mc2d_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def allreduce_stats(self, st: dict) -> dict:
        s = Stats()
        for k in range(3):
            s.attempted[k] = st["attempted"][k]
            s.accepted[k] = st["accepted"][k]
        s.n_particles, s.energy, s.overflow = st["n_particles"], st["energy"], st["overflow"]
        s.sweeps_done = st["sweeps_done"]
        self._check(self.L.mc2d_allreduce_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_counters(self):
        self._check(self.L.mc2d_reset_counters(self.h))

    def set_delta(self, delta):
        self._check(self.L.mc2d_set_delta(self.h, delta))

    def set_activity(self, z):
        self._check(self.L.mc2d_set_activity(self.h, z))

    def total_energy_virial(self, resync=False, allreduce=False):
        U, W, N = C.c_double(), C.c_double(), C.c_longlong()
        self._check(self.L.mc2d_total_energy_virial(self.h, int(resync), int(allreduce), C.byref(U), C.byref(W),
                                                    C.byref(N)))
        return U.value, W.value, N.value

    # -- calculator hooks on caller-owned vectors (reference grid) --------------------------------
    def calc_energy_virial(self, xy):
        a, n = self._xy(xy)
        U, W = C.c_double(), C.c_double()
        self._check(self.L.mc2d_calc_energy_virial(self.h, a.ctypes.data, n, C.byref(U), C.byref(W)))
        return U.value, W.value

    def calc_pressure(self, xy):
        """P = rho*T + W/(2V), thermodynamic_calculator.cpp:235-242."""
        a, n = self._xy(xy)
        _, W = self.calc_energy_virial(a)
        V = self.params.Lx * self.params.Ly
        return (n / V) * self.params.temperature + W / (2.0 * V)

    def calc_cell_ids(self, xy):
        a, n = self._xy(xy)
        out = np.empty(max(n, 1), dtype=np.int32)
        self._check(self.L.mc2d_calc_cell_ids(self.h, a.ctypes.data, n, out.ctypes.data))
        return out[:n]

    def calc_neighbor_counts(self, xy, with_energy=False):
        a, n = self._xy(xy)
        out = np.zeros(max(n, 1), dtype=np.int32)
        e = np.zeros(max(n, 1), dtype=np.float64) if with_energy else None
        self._check(self.L.mc2d_calc_neighbor_counts(self.h, a.ctypes.data, n, out.ctypes.data,
                                                     e.ctypes.data if with_energy else None))
        return (out[:n], e[:n]) if with_energy else out[:n]

    def calc_point_energies(self, xy, points, exclude=None, with_counts=False):
        a, n = self._xy(xy)
        p, nq = self._xy(points)
        ex = None
        if exclude is not None:
            ex = np.ascontiguousarray(exclude, dtype=np.int32)
            assert ex.shape[0] == nq
        e = np.zeros(max(nq, 1), dtype=np.float64)
        c = np.zeros(max(nq, 1), dtype=np.int32)
        self._check(self.L.mc2d_calc_point_energies(self.h, a.ctypes.data, n, p.ctypes.data,
                                                    ex.ctypes.data if ex is not None else None, nq, e.ctypes.data,
                                                    c.ctypes.data))
        return (e[:nq], c[:nq]) if with_counts else e[:nq]

    def calc_neighbor_list(self, xy, index=None, point=None, cap=4096):
        """CellList::getNeighbors(index) or getNeighbors2(point): (j, r2) in the reference's order."""
        a, n = self._xy(xy)
        j = np.empty(cap, dtype=np.int32)
        r2 = np.empty(cap, dtype=np.float64)
        cnt = C.c_int()
        qx, qy = (0.0, 0.0) if point is None else (float(point[0]), float(point[1]))
        self._check(self.L.mc2d_calc_neighbor_list(self.h, a.ctypes.data, n, -1 if index is None else int(index), qx,
                                                   qy, j.ctypes.data, r2.ctypes.data, cap, C.byref(cnt)))
        return j[: cnt.value].copy(), r2[: cnt.value].copy()

    def calc_energy_virial_bruteforce(self, xy):
        a, n = self._xy(xy)
        U, W = C.c_double(), C.c_double()
        self._check(self.L.mc2d_calc_energy_virial_bruteforce(self.h, a.ctypes.data, n, C.byref(U), C.byref(W)))
        return U.value, W.value

    # -- multi-GPU ----------------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == UNIQUE_ID_BYTES
        self._check(self.L.mc2d_comm_init(self.h, unique_id))

    # -- introspection ------------------------------------------------------------------------------
    def info(self) -> Info:
        o = Info()
        self._check(self.L.mc2d_get_info(self.h, C.byref(o)))
        return o

    def stream(self) -> int:
        s = C.c_void_p()
        self._check(self.L.mc2d_get_stream(self.h, C.byref(s)))
        return s.value

    def rescale_try(self, Lx_new, Ly_new):
        """NPT volume move, first half: (U, W) of the configuration rescaled to the new box (still pending)"""
        U, W = C.c_double(), C.c_double()
        self._check(self.L.mc2d_rescale_try(self.h, Lx_new, Ly_new, C.byref(U), C.byref(W)))
        return U.value, W.value

    def rescale_finish(self, accept: bool):
        self._check(self.L.mc2d_rescale_finish(self.h, int(bool(accept))))

    def probe_point(self, x, y):
        """energy of a test particle at (x, y) against the resident configuration"""
        dU = C.c_double()
        self._check(self.L.mc2d_probe_point(self.h, x, y, C.byref(dU)))
        return dU.value

    def probe_particle(self, index):
        """(x, y, dU) of the index-th particle in download order; dU = minus its energy with its neighbours"""
        x, y, dU = C.c_double(), C.c_double(), C.c_double()
        self._check(self.L.mc2d_probe_particle(self.h, index, C.byref(x), C.byref(y), C.byref(dU)))
        return x.value, y.value, dU.value

    def apply_probe(self, what: int):
        self._check(self.L.mc2d_apply_probe(self.h, what))

    def measure_fp64_peak(self) -> float:
        """measured fp64 FMA throughput of this device, TFLOP/s"""
        t = C.c_double()
        self._check(self.L.mc2d_measure_fp64_peak(self.h, C.byref(t)))
        return t.value

    def set_kernel_timing(self, on: bool):
        self._check(self.L.mc2d_set_kernel_timing(self.h, int(on)))

    def synchronize(self):
        self._check(self.L.mc2d_synchronize(self.h))
