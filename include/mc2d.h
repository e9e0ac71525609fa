/*
 * mc2d.h — C ABI of libmc2d.so, the B200 (sm_100a) implementation of the 2D Monte Carlo hot
 * path of ReyhanehFarimani/MonteCarlo_simulation.
 *
 * The reference has no FFI layer: its boundary is the C++ class API of serial_src/ and mpi_src/
 * linked into one binary.  This header is what those classes bind to when the
 * hot path runs on the GPU; montecarlo_simulation_b200/host/ holds the C++ shims that keep the
 * reference's class names and signatures and forward to these entry points.  Each entry point
 * cites the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns an mc2d_status and never
 * throws; all buffers are caller-owned HOST memory unless a name ends in _dev; one context per
 * GPU; calls on one context are serialised by the caller (the reference is single-threaded per
 * rank, SURVEY.md §8b "Threading").  Positions are packed {double x, y} — exactly the layout of
 * serial_src/initial.h:10-27 `Particle`, so `particles.data()` can be passed as is.
 */
#ifndef MC2D_H
#define MC2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC2D_ABI_VERSION 2

typedef enum {
    MC2D_OK = 0,
    MC2D_ERR_INVALID = 1,       /* bad argument (reference: std::invalid_argument / assert) */
    MC2D_ERR_CUDA = 2,          /* CUDA runtime failure; see mc2d_last_error */
    MC2D_ERR_NO_DEVICE = 3,     /* no CUDA device: there is NO CPU fallback */
    MC2D_ERR_CELL_OVERFLOW = 4, /* a cell needs more than cell_capacity slots */
    MC2D_ERR_GEOMETRY = 5,      /* box too small for the checkerboard (needs >= 2 cells of side >= rcut per direction, even counts per slab) */
    MC2D_ERR_OUT_OF_BOX = 6,    /* a position is NaN or outside [0,Lx] x [0,Ly] */
    MC2D_ERR_NCCL = 7,
    MC2D_ERR_STATE = 8,         /* call out of order (e.g. sweep before upload) */
    MC2D_ERR_PEER_TIMEOUT = 9   /* a neighbour rank did not signal its halo within tuning.spin_timeout_ms (it failed,
                                 * returned early or died): the kernel gave up waiting instead of hanging.  The slab
                                 * contexts of ALL ranks are out of step afterwards and must be destroyed. */
} mc2d_status;

/* PotentialType, in the order of serial_src/potential.h:11-18 */
typedef enum {
    MC2D_LENNARD_JONES = 0,
    MC2D_WCA = 1,
    MC2D_YUKAWA = 2,
    MC2D_ATHERMAL_STAR = 3,
    MC2D_THERMAL_STAR = 4,
    MC2D_IDEAL = 5
} mc2d_potential;

/* Everything the reference spreads over SimulationBox (initial.h:31-38), ThermodynamicCalculator
 * (thermodynamic_calculator.h:186-199) and MonteCarlo (MC.h:76-94). */
typedef struct {
    double Lx, Ly;  /* periodic box */
    double rcut;    /* interaction cutoff; cells have side >= rcut */
    int potential;  /* mc2d_potential */
    /* shape parameters ALREADY derived as thermodynamic_calculator.cpp:19-21 does
     * (f' = (2+9f^2)/24, f_d' = A0 f^2/kappa) and narrowed to float as potential.h:151 does */
    float f_prime, f_d_prime, kappa, alpha;
    double temperature;
    double activity;  /* z; the reference reads it from the input key "mu" (MC.cpp:25) */
    double delta;     /* max displacement per coordinate (MC.cpp:100-101) */
    uint64_t seed;    /* Philox key; one counter-based stream per cell */
    int device;       /* CUDA device ordinal */
    int rank, nranks; /* 1-D slab decomposition in x (replaces mpi_src Decomposition Px x Py) */
    int cell_capacity; /* particle slots per cell; 0 = choose from rcut and a dense-packing bound */
    double p_insert, p_delete; /* per GC draw (MC.cpp:194-215: 0.3 / 0.3, rest no-op); 0,0 = those defaults */
    double cell_margin; /* cells are made >= rcut * (1 + cell_margin) wide: room for the box to shrink under
                         * NPT volume moves without changing the cell counts (mc2d_rescale_try), or fuller cells
                         * for short-range potentials.  0: cells of rcut.  -1: chosen at the first upload so that
                         * a cell holds ~2.5 particles (single rank; never narrower than rcut) */
} mc2d_params;

typedef struct {
    long long attempted[3]; /* displacement, insertion, deletion trials (no-ops are not counted) */
    long long accepted[3];
    long long n_particles;  /* owned by this rank */
    double energy;          /* running potential energy of this rank's moves + its share at the last resync */
    long long overflow;     /* insertions refused because the cell had no free slot (should be 0) */
    unsigned long long sweeps_done;
} mc2d_stats;

/* what one sweep does.  One sweep = 4 colour passes; in each pass every cell of that colour gets
 * disp_per_particle * n_cell displacement trials (MC.cpp:38-44: N trials per sweep) followed by
 * gc_per_cell grand-canonical draws (MC.cpp:45-52), each an insertion / deletion / no-op. */
typedef struct {
    int disp_per_particle; /* 0 disables displacements */
    int gc_per_cell;       /* 0 = NVT */
    int shift_grid;        /* 1: random grid shift before every sweep (ergodicity); 0: frozen grid */
    int disp_per_cell;     /* > 0: every non-empty cell gets exactly this many displacement trials per visit
                            * instead of disp_per_particle * n_cell (a fixed "nselect": same ensemble, all
                            * cells of a warp finish together) */
} mc2d_sweep_plan;

typedef struct mc2d_ctx mc2d_ctx;

const char *mc2d_status_string(int status);
const char *mc2d_last_error(const mc2d_ctx *ctx); /* detail text of the last failure on ctx */
int mc2d_abi_version(void);
/* number of visible CUDA devices (0 on a CPU-only host; never an error) */
int mc2d_device_count(void);

/* -- lifetime (MonteCarlo ctor MC.cpp:4-29 / dtor) ------------------------------------------- */
int mc2d_create(const mc2d_params *params, mc2d_ctx **out);
int mc2d_destroy(mc2d_ctx *ctx);

/* -- pure host geometry (no GPU needed) ------------------------------------------------------ */
/* CellList ctor, cell_list.cpp:5-15: nx = int(Lx/rcut) (>=1), cell_dx = Lx/nx */
int mc2d_reference_grid(double Lx, double Ly, double rcut, int *nx, int *ny, double *cell_dx, double *cell_dy);
/* checkerboard grid: largest nx that is a multiple of 2*nranks with Lx/nx >= rcut (CellListParallel's
 * even-count rule, cell_list_parallel.cpp:64-106, applied per slab); ny largest even with Ly/ny >= rcut */
int mc2d_sweep_grid(double Lx, double Ly, double rcut, int nranks, int *nx, int *ny, double *cell_dx,
                    double *cell_dy);
/* slab of `rank`: first owned cell column and number of owned columns (replaces
 * Decomposition::localBounds, simulatin_box.cpp:91-98, for Px = nranks, Py = 1) */
int mc2d_slab_columns(int nx, int nranks, int rank, int *col0, int *ncol);
/* which boundary column a colour pass modifies and where it goes: colour = px + 2*py.
 * *send_left = 1: the rank's first owned column goes to rank-1 (its right ghost); else the last
 * owned column goes to rank+1.  (replaces the 8-neighbour flushGhostUpdates, particle_exchange.cpp:247-258) */
int mc2d_halo_route(int colour, int rank, int nranks, int *send_left, int *send_to, int *recv_from);
/* the colour order and grid shift of sweep `sweep` — a pure function of (seed, sweep), identical on
 * every rank (replaces bcast_seed_, MC_parallel.cpp:218-230) */
int mc2d_sweep_schedule(uint64_t seed, uint64_t sweep, double cell_dx, double cell_dy, int shift_grid,
                        int colour_order[4], double *shift_x, double *shift_y);

/* -- particles --------------------------------------------------------------------------------
 * upload: takes n packed positions (ids optional, NULL = 0..n-1) and keeps those inside this
 * rank's slab (all of them when nranks == 1); bins them on the checkerboard grid by a key sort +
 * cell-offset pass (replaces CellList::build, cell_list.cpp:28-41, and buildInterior,
 * cell_list_parallel.cpp:116-132) and computes the initial energy (MC.cpp:26). */
int mc2d_upload(mc2d_ctx *ctx, const double *xy, const int *ids, long long n);
/* the same with options.  MC2D_UPLOAD_NO_ENERGY: skip the initial-energy pass (MC.cpp:26) — for callers that re-sync the
 * energy after their sweeps anyway (mc2d_total_energy_virial with resync != 0, the reference's own sampling pattern,
 * MC.cpp:62-71); until then stats.energy is the SUM OF ACCEPTED dE since this upload, not the total energy. */
#define MC2D_UPLOAD_NO_ENERGY 1u
int mc2d_upload_ex(mc2d_ctx *ctx, const double *xy, const int *ids, long long n, unsigned flags);
/* download: owned particles in cell order; *n_out receives the count (cap = capacity of xy/ids in
 * particles; MC2D_ERR_INVALID if too small).  ids may be NULL.  Particles inserted by GC moves
 * carry id -1. */
int mc2d_download(mc2d_ctx *ctx, double *xy, int *ids, long long cap, long long *n_out);
int mc2d_num_particles(mc2d_ctx *ctx, long long *n_local);
/* One sample point (MonteCarlo::recordObservables: Logging::logSimulationData + logPositions, MC.cpp:62-78): the
 * energy/virial reduction of mc2d_total_energy_virial AND the download of mc2d_download in one call — the reduction runs
 * on the device while the positions are still on their way to the host. */
int mc2d_sample(mc2d_ctx *ctx, int resync, int allreduce, double *U, double *W, long long *N, double *xy, int *ids,
                long long cap, long long *n_out);

/* -- the step loop (MonteCarlo::run, MC.cpp:31-78; MonteCarloNVT_MPI::run, MC_parallel.cpp:57-114) */
int mc2d_run_sweeps(mc2d_ctx *ctx, long long nsweeps, const mc2d_sweep_plan *plan, mc2d_stats *stats_out);
int mc2d_get_stats(mc2d_ctx *ctx, mc2d_stats *stats_out);
int mc2d_reset_counters(mc2d_ctx *ctx);
int mc2d_set_delta(mc2d_ctx *ctx, double delta);
int mc2d_set_activity(mc2d_ctx *ctx, double z);

/* full-system energy U, virial W (sum over pairs of r.f) and particle count of the context's
 * current configuration (computeTotalEnergyCellList / computeTotalVirialCellList,
 * thermodynamic_calculator.cpp:194-232, in ONE pass).  resync != 0 stores U as the running energy
 * (MC.cpp:62-71).  Local to this rank unless allreduce != 0 and a communicator is attached
 * (replaces the MPI_Allreduce sites of thermodynamic_calculator_parallel.cpp:41-131). */
int mc2d_total_energy_virial(mc2d_ctx *ctx, int resync, int allreduce, double *U, double *W, long long *N);

/* -- ThermodynamicCalculator on caller-owned particle vectors, on the REFERENCE grid
 * (nx = int(Lx/rcut)), results comparable bit-for-bit / to 1e-10 with the serial reference ---- */
/* computeTotalEnergyCellList + computeTotalVirialCellList (thermodynamic_calculator.cpp:194-232);
 * pressure = rho*T + W/(2V) (:235-242) follows on the host. */
int mc2d_calc_energy_virial(mc2d_ctx *ctx, const double *xy, long long n, double *U, double *W);
/* CellList::cellIndex for each particle (cell_list.cpp:103-109), bit-exact */
int mc2d_calc_cell_ids(mc2d_ctx *ctx, const double *xy, long long n, int *cell_out);
/* getNeighbors(i).size() for each particle (cell_list.cpp:43-70, inclusive cutoff), bit-exact;
 * local_energy_out (optional) = computeLocalEnergy over that list (thermodynamic_calculator.h:103-121) */
int mc2d_calc_neighbor_counts(mc2d_ctx *ctx, const double *xy, long long n, int *count_out,
                              double *local_energy_out);
/* energy of nq trial points against the configuration: computeLocalEnergy over getNeighbors2(point)
 * (cell_list.cpp:72-101) skipping particle exclude[q] when >= 0 (exclude may be NULL).  This is
 * the dE hook: displacement dE = E(new, exclude=i) - E(old_i, exclude=i); insertion dE = E(pt);
 * deletion dE = -E(old_i, exclude=i).  count_out optional. */
int mc2d_calc_point_energies(mc2d_ctx *ctx, const double *xy, long long n, const double *points,
                             const int *exclude, long long nq, double *energy_out, int *count_out);

/* the neighbour LIST of particle query_index (>= 0; self skipped by index, CellList::getNeighbors,
 * cell_list.cpp:43-70) or, with query_index < 0, of the point (qx, qy) (getNeighbors2,
 * cell_list.cpp:72-101): (j, r^2) pairs in the reference's visiting order.  *count_out = list length
 * (may exceed cap; only cap entries are written). */
int mc2d_calc_neighbor_list(mc2d_ctx *ctx, const double *xy, long long n, long long query_index, double qx,
                            double qy, int *out_j, double *out_r2, int cap, int *count_out);
/* brute-force U and W over all pairs (computeTotalEnergy / computeTotalVirial,
 * thermodynamic_calculator.cpp:56-101) — O(N^2), the reference's test truth */
int mc2d_calc_energy_virial_bruteforce(mc2d_ctx *ctx, const double *xy, long long n, double *U, double *W);

/* -- trial log (parity hook for per-trial dE of the sweep kernel) ----------------------------------- */
typedef struct {
    int pass;        /* 0..3 position of the colour pass inside the sweep */
    int cell;        /* global cell id (column-major: cx*ny + cy) */
    int seq;         /* order of the trial inside its cell visit */
    int kind;        /* 0 displacement, 1 insertion, 2 deletion */
    int accepted;
    int reserved;
    double old_x, old_y; /* displacement/deletion: the particle's position before the trial */
    double new_x, new_y; /* displacement: trial position (wrapped into the box); insertion: the point */
    double dE;           /* energy change the kernel used in the acceptance rule (NaN if rejected before evaluation) */
} mc2d_trial;
/* runs ONE sweep and records up to cap trials; *n_out = number of trials that happened */
int mc2d_run_sweep_traced(mc2d_ctx *ctx, const mc2d_sweep_plan *plan, mc2d_trial *trials, long long cap,
                          long long *n_out, mc2d_stats *stats_out);

/* -- multi-GPU (replaces ParticleExchange, particle_exchange.h:33-51, and the Allreduce sites) ---- */
#define MC2D_UNIQUE_ID_BYTES 128
/* rank 0 creates the id, the host program ships it to the other ranks (MPI_Bcast in the reference's
 * world, torch.distributed in bench.py), then every rank calls mc2d_comm_init. */
int mc2d_comm_unique_id(void *id_out /* MC2D_UNIQUE_ID_BYTES */);
int mc2d_comm_init(mc2d_ctx *ctx, const void *unique_id);
/* sums stats over ranks (one ncclAllReduce of a small double vector) */
int mc2d_allreduce_stats(mc2d_ctx *ctx, mc2d_stats *inout);

/* -- NPT volume move (replaces MonteCarlo::npt_move_no_cell_list, serial_src/MC.cpp:153-188) ----
 * mc2d_rescale_try: every position is multiplied by (Lx_new/Lx, Ly_new/Ly) (SimulationBox::recenter,
 *   initial.cpp:27-30) into the spare slot buffer and U, W of the rescaled configuration are computed
 *   there (computeTotalEnergyCellList).  The context keeps its present configuration until
 * mc2d_rescale_finish(accept != 0) switches to the rescaled one (running energy = U_new); accept == 0
 *   drops it.
 * The cell COUNTS are kept (the scaling maps cells onto cells: no re-binning), so the new cells must
 * still be >= rcut wide: otherwise MC2D_ERR_GEOMETRY and nothing is pending — the caller re-creates
 * the context for the new box.  Single-rank contexts only (the reference's MPI path is NVT only). */
int mc2d_rescale_try(mc2d_ctx *ctx, double Lx_new, double Ly_new, double *U_new, double *W_new);
int mc2d_rescale_finish(mc2d_ctx *ctx, int accept);

/* -- single-particle probes on the resident configuration (Gibbs-ensemble particle exchange,
 *    GibbsMC.cpp:340-394; also Widom insertion) ---------------------------------------------------
 * mc2d_probe_point:    dU = energy of a test particle at (x, y) against the configuration
 *                      (getNeighbors2 + computeLocalEnergy, exact reference arithmetic).
 * mc2d_probe_particle: the index-th particle in the order mc2d_download would return them
 *                      (0 <= index < N): its position and dU = MINUS its energy with its neighbours.
 * mc2d_apply_probe:    what != 0 carries out the probed insertion (what = 1) or deletion (what = 2):
 *                      positions, counts and the running energy are updated; 0 forgets the probe.
 * Single-rank contexts; MC2D_ERR_CELL_OVERFLOW if the target cell of an insertion has no free slot. */
int mc2d_probe_point(mc2d_ctx *ctx, double x, double y, double *dU);
int mc2d_probe_particle(mc2d_ctx *ctx, long long index, double *x, double *y, double *dU);
int mc2d_apply_probe(mc2d_ctx *ctx, int what);

/* -- tuning / A-B switches -------------------------------------------------------------------------
 * Every variant gives the same trajectory (tests/test_gpu_parity.py::test_sweep_kernel_variants_are_bit_identical).
 * mc2d_create fills the defaults; mc2d_get_tuning / mc2d_set_tuning read and replace them between calls.  The
 * library reads no environment variables. */
typedef struct {
    int sweep_lanes;          /* -1 auto (4, or 8 for very full cells); 0: k_sweep, one warp per cell; 2, 4, 8: lanes per
                               * cell of k_sweep_p */
    int sweep_slots;          /* 0 auto; candidate slots of shared memory per warp of k_sweep_p (raised to one full cell
                               * neighbourhood) */
    int sweep_warps_per_cta;  /* 0 auto (8) */
    int fuse_passes;          /* 1: the four colour passes of a sweep are ONE cooperative launch; 0: one launch per pass */
    int pipeline;             /* 1: per-column-pair completion counters between the colours; 0: grid barrier */
    int rebin_lanes;          /* 4 (default) or 8 lanes per new cell in the re-bin */
    int energy_unstaged;      /* 1: k_energy_virial_g instead of k_energy_p */
    int halo_nccl;            /* 1: ncclSend/ncclRecv of whole columns instead of stores into the neighbours' memory */
    int halo_overlap;         /* 1: slab-edge work overlaps the interior (second stream / inside the sweep kernel) */
    int upload_sort;          /* 1: every mc2d_upload takes the key-sort path (default: direct binning on re-upload) */
    int spin_timeout_ms;      /* budget of every in-kernel wait on a neighbour's flag word; 0 = default (10 000 ms) */
    int test_upload_lag_us;   /* fault injection for tools/check_multi_gpu.py (odd ranks sleep after the capacity collective
                               * of an upload); honoured only by a library built with -DMC2D_FAULT_INJECTION */
    int reserved[4];
} mc2d_tuning;
int mc2d_get_tuning(mc2d_ctx *ctx, mc2d_tuning *out);
int mc2d_set_tuning(mc2d_ctx *ctx, const mc2d_tuning *in);

/* -- introspection for benchmarks ------------------------------------------------------------ */
typedef struct {
    int nx, ny;            /* checkerboard grid */
    int col0, ncol;        /* owned columns */
    int cell_capacity;
    double cell_dx, cell_dy;
    long long launches;    /* kernels launched by this context so far */
    double last_sweep_ms;  /* device time of the last mc2d_run_sweeps call (CUDA events) */
    double last_energy_ms; /* device time of the last energy/virial reduction */
    long long last_pair_tests; /* directed distance tests of the last energy/virial reduction */
    long long last_pair_evals; /* directed in-cutoff pair evaluations of the last reduction */
    /* per-kernel device time of the last mc2d_run_sweeps call made with kernel timing on
     * (CUDA events around every launch on the context's stream) */
    double k_sweep_ms;
    long long k_sweep_launches;
    double k_rebin_ms;
    long long k_rebin_launches;
    int halo_mode; /* 0: single rank (no halo), 1: ncclSend/ncclRecv, 2: stores into the neighbours' memory (CUDA IPC) */
    int sweep_kernel_lanes; /* lanes per cell of the sweep kernel last launched: 4 or 8 (k_sweep_p), 32 (k_sweep) */
    double k_halo_ms;       /* with kernel timing on: summed halo push + sync time of the last mc2d_run_sweeps */
    long long k_halo_launches;
    long long slot_allocs; /* how often the slot storage was (re-)allocated: an upload whose cells need more slots than
                            * the storage has, or far fewer, re-allocates (and re-maps the neighbours' memory) */
} mc2d_info;
int mc2d_get_info(mc2d_ctx *ctx, mc2d_info *out);
/* on != 0: bracket every sweep / re-bin kernel with CUDA events (adds launch gaps: use it for the
 * roofline measurement, not for the throughput number) */
int mc2d_set_kernel_timing(mc2d_ctx *ctx, int on);
/* raw CUDA stream (cudaStream_t) the context launches on, for callers that time with events */
int mc2d_get_stream(mc2d_ctx *ctx, void **stream_out);
int mc2d_synchronize(mc2d_ctx *ctx);

/* measured fp64 FMA throughput of the context's device in TFLOP/s (a short kernel of independent DFMA chains on
 * every SM): the compute roofline bench.py reports next to the HBM one, which this path never approaches */
int mc2d_measure_fp64_peak(mc2d_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MC2D_H */
