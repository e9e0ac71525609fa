/*
 * mc2d_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's serial 2D Monte Carlo hot path
 * (ReyhanehFarimani/MonteCarlo_simulation, serial_src/).  Every function cites
 * the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (montecarlo_simulation_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this restatement
 *   (a) against the reference's own known-answer tests (unit_test_serial/*.cpp),
 *   (b) bit-for-bit against the reference itself compiled in place into
 *       oracle/_ref/libmc2d_ref.so (see oracle/Makefile, oracle/ref_shim.cpp),
 *   (c) against the sanity_check golden CSV (tests/golden/).
 */
#ifndef MC2D_ORACLE_H
#define MC2D_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PotentialType enum order — serial_src/potential.h:11-18 */
enum {
    ORC_LJ = 0,
    ORC_WCA = 1,
    ORC_YUKAWA = 2,
    ORC_ATHERMAL_STAR = 3,
    ORC_THERMAL_STAR = 4,
    ORC_IDEAL = 5
};

/* Ensemble enum order — serial_src/MC.h:14-19 */
enum { ORC_NVT = 0, ORC_GCMC = 1 };

/* potential.cpp:212-256 (dispatch) and :11-186 (leaf functions).  The four shape
 * parameters are float exactly as in potential.h:151,164. */
double orc_pair_potential(double r2, int type, float f_prime, float f_d_prime, float kappa, float alpha);
double orc_pair_force(double r2, int type, float f_prime, float f_d_prime, float kappa, float alpha);
/* potential.cpp:194-202; returns -1 for an unknown name (reference throws). */
int orc_select_potential(const char *name);

/* thermodynamic_calculator.cpp:19-21: f' = (2+9f^2)/24, f_d' = A0 f^2 / kappa */
void orc_calc_params(double f, double A0, double kappa, double *f_prime, double *f_d_prime);

/* initial.cpp:22-25 and :43-51 */
void orc_apply_pbc(double Lx, double Ly, double *x, double *y);
double orc_min_image_r2(double Lx, double Ly, double x1, double y1, double x2, double y2);

/* The calculator state — thermodynamic_calculator.h:186-199 */
typedef struct {
    double T, press, z; /* "mu_" is consumed as the activity z (calculator.cpp:39-41) */
    double rcut, r2cut;
    int type;
    double f_prime, alpha, f_d_prime, kappa; /* held as double, narrowed per call */
} orc_calc;

void orc_calc_init(orc_calc *c, double T, double press, int type, double rcut, double mu, double f,
                   double alpha, double A0, double kappa);

/* ---- cell list: cell_list.cpp ------------------------------------------- */
typedef struct orc_celllist orc_celllist;
orc_celllist *orc_cl_create(double Lx, double Ly, double rcut); /* cell_list.cpp:5-15 */
void orc_cl_destroy(orc_celllist *cl);
void orc_cl_dims(const orc_celllist *cl, int *nx, int *ny, double *cell_dx, double *cell_dy);
int orc_cl_cell_index(const orc_celllist *cl, double x, double y); /* cell_list.cpp:103-109 */
void orc_cl_build(orc_celllist *cl, const double *xy, int N);      /* cell_list.cpp:28-41 */
/* cell_list.cpp:43-70: returns the number of neighbours; writes up to cap (j, r2) pairs in
 * the reference's visiting order. */
int orc_cl_neighbors(const orc_celllist *cl, int idx, const double *xy, int *out_j, double *out_r2, int cap);
/* cell_list.cpp:72-101 (arbitrary point, self NOT skipped) */
int orc_cl_neighbors_point(const orc_celllist *cl, double x, double y, const double *xy, int *out_j,
                           double *out_r2, int cap);

/* ---- observables: thermodynamic_calculator.cpp -------------------------- */
double orc_total_energy(const orc_calc *c, double Lx, double Ly, const double *xy, int N);          /* :56-77 brute force */
double orc_total_virial(const orc_calc *c, double Lx, double Ly, const double *xy, int N);          /* :80-101 */
double orc_pressure(const orc_calc *c, double Lx, double Ly, const double *xy, int N);              /* :104-111 */
double orc_total_energy_celllist(const orc_calc *c, double Lx, double Ly, const double *xy, int N); /* :194-211 */
double orc_total_virial_celllist(const orc_calc *c, double Lx, double Ly, const double *xy, int N); /* :215-232 */
double orc_pressure_celllist(const orc_calc *c, double Lx, double Ly, const double *xy, int N);     /* :235-242 */

/* parity hooks: per-particle cell id (cellIndex on the reference grid) and the size of
 * getNeighbors(i) on a freshly built list. */
void orc_cell_ids(double Lx, double Ly, double rcut, const double *xy, int N, int *out_cell);
void orc_neighbor_counts(double Lx, double Ly, double rcut, const double *xy, int N, int *out_count);
/* computeLocalEnergy (thermodynamic_calculator.h:103-121) over getNeighbors(idx) of a fresh list */
double orc_local_energy(const orc_calc *c, double Lx, double Ly, const double *xy, int N, int idx);
/* computeLocalEnergy over getNeighbors2(point) of a fresh list; exclude_idx >= 0 drops that
 * particle from the sum (what a displaced particle sees at its trial position). */
double orc_point_energy(const orc_calc *c, double Lx, double Ly, const double *xy, int N, double x, double y,
                        int exclude_idx);

/* ---- RNG: rng.cpp (std::mt19937 + libstdc++ distributions) -------------- */
typedef struct {
    uint32_t mt[624];
    int idx;
} orc_rng;
void orc_rng_seed(orc_rng *r, uint32_t seed);
uint32_t orc_rng_next_u32(orc_rng *r);
double orc_rng_uniform01(orc_rng *r);                 /* rng.cpp:9-11 */
double orc_rng_uniform(orc_rng *r, double a, double b); /* rng.cpp:14-16 */
int orc_rng_uniform_int(orc_rng *r, int lo, int hi);  /* rng.cpp:19-22 */

/* initial.cpp:86-104 random branch: glibc srand(seed+10)/rand() (seed==0 -> time-seeded in
 * the reference; refused here). */
int orc_initialize_particles_random(double Lx, double Ly, int N, unsigned seed, double *xy);

/* ---- the serial step loop: MC.cpp --------------------------------------- */
typedef struct orc_mc orc_mc;
orc_mc *orc_mc_create(double Lx, double Ly, const double *xy, int N, const orc_calc *calc, double rcut,
                      double delta, int ensemble, uint32_t rng_seed); /* MC.cpp:4-29 */
void orc_mc_destroy(orc_mc *mc);
/* MC.cpp:31-78.  Logging is replaced by sampling: at every step where the reference would call
 * recordObservables ((step+1)%fOutputStep==0 || step==nSteps-1) one sample
 * {simulation_step_time, N, running energy} is appended (up to cap).  Returns #samples. */
int orc_mc_run(orc_mc *mc, size_t nSteps, size_t fOutputStep, size_t fUpdateCell, long long *s_time, int *s_N,
               double *s_E, int cap);
int orc_mc_step_celllist(orc_mc *mc);   /* MC.cpp:96-123 */
int orc_mc_step_bruteforce(orc_mc *mc); /* MC.cpp:125-151 */
int orc_mc_gc_move(orc_mc *mc);         /* MC.cpp:189-249 */
void orc_mc_update_celllist(orc_mc *mc); /* MC.cpp:251-253 */
double orc_mc_energy(const orc_mc *mc);
int orc_mc_n(const orc_mc *mc);
void orc_mc_get_particles(const orc_mc *mc, double *xy);
long long orc_mc_accepted(const orc_mc *mc);
long long orc_mc_total(const orc_mc *mc);

/* ---- golden-input generator: sanity_check/main_parallel.cpp:73-80 -------------------------
 * Regenerates the particle set of one sanity_check frame exactly as the MPI run with np ranks on
 * a Px x Py decomposition created it: for rank r, initializeParticles_globalN(... random, seed,
 * id_stride) (mpi_src/initial_mpi.cpp:21-95) drawing from RNG_parallel(seed, rank)
 * (mpi_src/rng_parallel.h:16-34: std::mt19937_64 seeded by std::seed_seq{seed, rank, 3 mixers},
 * generate_canonical<double,53>), tile bounds from Decomposition::localBounds
 * (mpi_src/simulatin_box.cpp:91-98).  Ranks are concatenated in rank order.  xy holds N_global
 * pairs.  Returns N_global. */
int orc_sanity_frame(double Lx, double Ly, int Px, int Py, int N_global, unsigned seed, double *xy);
/* mpi_src/simulatin_box.cpp:100-128 */
void orc_best_decomposition(double Lx, double Ly, int nranks, int *Px, int *Py);

#ifdef __cplusplus
}
#endif
#endif
