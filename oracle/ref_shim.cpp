// ref_shim.cpp — TEST INFRASTRUCTURE.  extern "C" façade over the UNMODIFIED reference
// serial_src/ classes, compiled together with the reference's own .cpp files where they lie
// under /root/reference (see oracle/Makefile) into oracle/_ref/libmc2d_ref.so.
//
// Nothing here restates reference arithmetic: every entry point forwards to the reference's
// own functions so that the C restatement in mc2d_oracle.c can be pinned bit-for-bit, and so
// that bench.py's cpu_baseline / --impl reference leg can time the real reference.
// The symbols mirror mc2d_oracle.h with the prefix ref_ instead of orc_.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

// The reference keeps cellIndex() and grandCanonicalMove() private; the shim needs to call
// them directly.  Access specifiers do not change object layout under the Itanium ABI.
#define private public
#include "cell_list.h"
#include "MC.h"
#undef private
#include "initial.h"
#include "potential.h"
#include "rng.h"
#include "thermodynamic_calculator.h"

namespace {
std::vector<Particle> to_particles(const double *xy, int N) {
    std::vector<Particle> p;
    p.reserve(N > 0 ? N : 0);
    for (int i = 0; i < N; ++i) p.emplace_back(xy[2 * i], xy[2 * i + 1]);
    return p;
}
struct CalcArgs {
    double T, press;
    int type;
    double rcut, mu, f, alpha, A0, kappa;
};
ThermodynamicCalculator make_calc(const CalcArgs &a) {
    return ThermodynamicCalculator(a.T, a.press, static_cast<PotentialType>(a.type), a.rcut, a.mu, a.f, a.alpha,
                                   a.A0, a.kappa);
}
}  // namespace

extern "C" {

double ref_pair_potential(double r2, int type, float fp, float fdp, float kappa, float alpha) {
    return computePairPotential(r2, static_cast<PotentialType>(type), fp, fdp, kappa, alpha);
}
double ref_pair_force(double r2, int type, float fp, float fdp, float kappa, float alpha) {
    return computePairForce(r2, static_cast<PotentialType>(type), fp, fdp, kappa, alpha);
}
int ref_select_potential(const char *name) {
    try {
        return static_cast<int>(selectPotentialType(name));
    } catch (const std::invalid_argument &) {
        return -1;
    }
}
void ref_apply_pbc(double Lx, double Ly, double *x, double *y) {
    SimulationBox box(Lx, Ly);
    Particle p(*x, *y);
    box.applyPBC(p);
    *x = p.x;
    *y = p.y;
}
double ref_min_image_r2(double Lx, double Ly, double x1, double y1, double x2, double y2) {
    SimulationBox box(Lx, Ly);
    return box.minimumImageDistanceSquared(Particle(x1, y1), Particle(x2, y2));
}

void ref_cl_dims(double Lx, double Ly, double rcut, int *nx, int *ny, double *dx, double *dy) {
    SimulationBox box(Lx, Ly);
    CellList cl(box, rcut);
    *nx = cl.nx_;
    *ny = cl.ny_;
    *dx = cl.cell_dx_;
    *dy = cl.cell_dy_;
}
void ref_cell_ids(double Lx, double Ly, double rcut, const double *xy, int N, int *out) {
    SimulationBox box(Lx, Ly);
    CellList cl(box, rcut);
    for (int i = 0; i < N; ++i) out[i] = cl.cellIndex(xy[2 * i], xy[2 * i + 1]);
}
void ref_neighbor_counts(double Lx, double Ly, double rcut, const double *xy, int N, int *out) {
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    CellList cl(box, rcut);
    cl.build(P);
    for (int i = 0; i < N; ++i) out[i] = static_cast<int>(cl.getNeighbors(i, P).size());
}
int ref_neighbors(double Lx, double Ly, double rcut, const double *xy, int N, int idx, int *out_j, double *out_r2,
                  int cap) {
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    CellList cl(box, rcut);
    cl.build(P);
    auto nb = cl.getNeighbors(idx, P);
    for (size_t k = 0; k < nb.size() && (int)k < cap; ++k) {
        out_j[k] = nb[k].first;
        out_r2[k] = nb[k].second;
    }
    return static_cast<int>(nb.size());
}
int ref_neighbors_point(double Lx, double Ly, double rcut, const double *xy, int N, double x, double y, int *out_j,
                        double *out_r2, int cap) {
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    CellList cl(box, rcut);
    cl.build(P);
    auto nb = cl.getNeighbors2(Particle(x, y), P);
    for (size_t k = 0; k < nb.size() && (int)k < cap; ++k) {
        out_j[k] = nb[k].first;
        out_r2[k] = nb[k].second;
    }
    return static_cast<int>(nb.size());
}

// what: 0 U brute force, 1 W brute force, 2 P brute force, 3 U cell list, 4 W cell list, 5 P cell list
double ref_observable(int what, double T, double press, int type, double rcut, double mu, double f, double alpha,
                      double A0, double kappa, double Lx, double Ly, const double *xy, int N) {
    auto calc = make_calc({T, press, type, rcut, mu, f, alpha, A0, kappa});
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    switch (what) {
    case 0: return calc.computeTotalEnergy(P, box);
    case 1: return calc.computeTotalVirial(P, box);
    case 2: return calc.computePressure(P, box);
    case 3: return calc.computeTotalEnergyCellList(P, box);
    case 4: return calc.computeTotalVirialCellList(P, box);
    case 5: return calc.computePressureCellList(P, box);
    }
    return NAN;
}
double ref_local_energy(double T, int type, double rcut, double f, double alpha, double A0, double kappa, double Lx,
                        double Ly, const double *xy, int N, int idx) {
    auto calc = make_calc({T, 0.0, type, rcut, 0.0, f, alpha, A0, kappa});
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    CellList cl(box, rcut);
    cl.build(P);
    auto nb = cl.getNeighbors(idx, P);
    return calc.computeLocalEnergy(idx, P, box, nb);
}
double ref_point_energy(double T, int type, double rcut, double f, double alpha, double A0, double kappa, double Lx,
                        double Ly, const double *xy, int N, double x, double y, int exclude_idx) {
    auto calc = make_calc({T, 0.0, type, rcut, 0.0, f, alpha, A0, kappa});
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    CellList cl(box, rcut);
    cl.build(P);
    auto nb = cl.getNeighbors2(Particle(x, y), P);
    std::vector<std::pair<int, double>> kept;
    for (auto &pr : nb)
        if (pr.first != exclude_idx) kept.push_back(pr);
    return calc.computeLocalEnergy(0, P, box, kept);
}

// ---- RNG ------------------------------------------------------------------
void *ref_rng_create(unsigned seed) { return new RNG(seed); }
void ref_rng_destroy(void *r) { delete static_cast<RNG *>(r); }
double ref_rng_uniform01(void *r) { return static_cast<RNG *>(r)->uniform01(); }
double ref_rng_uniform(void *r, double a, double b) { return static_cast<RNG *>(r)->uniform(a, b); }
int ref_rng_uniform_int(void *r, int lo, int hi) { return static_cast<RNG *>(r)->uniformInt(lo, hi); }

int ref_initialize_particles_random(double Lx, double Ly, int N, unsigned seed, double *xy) {
    if (seed == 0) return -1;
    SimulationBox box(Lx, Ly);
    std::vector<Particle> P;
    initializeParticles(P, box, N, true, seed);
    for (int i = 0; i < N; ++i) {
        xy[2 * i] = P[i].x;
        xy[2 * i + 1] = P[i].y;
    }
    return 0;
}

// ---- MonteCarlo -------------------------------------------------------------
struct RefMC {
    SimulationBox box;
    std::vector<Particle> particles;
    ThermodynamicCalculator calc;
    Logging logger;
    RNG rng;
    std::unique_ptr<MonteCarlo> mc;
    RefMC(double Lx, double Ly, const double *xy, int N, const CalcArgs &a, double delta, int ensemble,
          unsigned seed)
        : box(Lx, Ly), particles(to_particles(xy, N)), calc(make_calc(a)), logger("/dev/null", "/dev/null"),
          rng(seed) {
        mc.reset(new MonteCarlo(box, particles, calc, a.rcut, delta, 0.0, static_cast<Ensemble>(ensemble), logger,
                                rng));
    }
};

void *ref_mc_create(double Lx, double Ly, const double *xy, int N, double T, double press, int type, double rcut,
                    double mu, double f, double alpha, double A0, double kappa, double delta, int ensemble,
                    unsigned seed) {
    return new RefMC(Lx, Ly, xy, N, {T, press, type, rcut, mu, f, alpha, A0, kappa}, delta, ensemble, seed);
}
void ref_mc_destroy(void *h) { delete static_cast<RefMC *>(h); }

// The reference logs instead of sampling.  run() is called in chunks of fOutputStep sweeps;
// each chunk restarts the reference's step counter at 0, so the cell-update cadence equals one
// long run() only when fOutputStep is a multiple of fUpdateCell (asserted).  The observable
// logging at the end of every chunk goes to /dev/null and draws no random numbers.
int ref_mc_run(void *h, size_t nSteps, size_t fOutputStep, size_t fUpdateCell, long long *s_time, int *s_N,
               double *s_E, int cap) {
    RefMC *m = static_cast<RefMC *>(h);
    if (fOutputStep % fUpdateCell != 0 || nSteps % fOutputStep != 0) return -1;
    int ns = 0;
    for (size_t done = 0; done < nSteps; done += fOutputStep) {
        m->mc->run(fOutputStep, fOutputStep, fUpdateCell);
        if (ns < cap) {
            if (s_time) s_time[ns] = m->mc->simulation_step_time;
            if (s_N) s_N[ns] = static_cast<int>(m->particles.size());
            if (s_E) s_E[ns] = m->mc->getEnergy();
        }
        ++ns;
    }
    return ns;
}
int ref_mc_step_celllist(void *h) { return static_cast<RefMC *>(h)->mc->stepCellList() ? 1 : 0; }
int ref_mc_step_bruteforce(void *h) { return static_cast<RefMC *>(h)->mc->stepBruteForce() ? 1 : 0; }
int ref_mc_gc_move(void *h) { return static_cast<RefMC *>(h)->mc->grandCanonicalMove() ? 1 : 0; }
void ref_mc_update_celllist(void *h) { static_cast<RefMC *>(h)->mc->updateCellList(); }
double ref_mc_energy(void *h) { return static_cast<RefMC *>(h)->mc->getEnergy(); }
int ref_mc_n(void *h) { return static_cast<int>(static_cast<RefMC *>(h)->particles.size()); }
// NPT: the reference takes delta_V in the constructor and copies the box (MC.h:76-83); both are private
// (reachable here through the `#define private public` above)
void ref_mc_set_delta_V(void *h, double dv) { static_cast<RefMC *>(h)->mc->delta_V = dv; }
double ref_mc_volume(void *h) { return static_cast<RefMC *>(h)->mc->box_.getV(); }
void ref_mc_get_particles(void *h, double *xy) {
    auto &P = static_cast<RefMC *>(h)->particles;
    for (size_t i = 0; i < P.size(); ++i) {
        xy[2 * i] = P[i].x;
        xy[2 * i + 1] = P[i].y;
    }
}

// ---- timing helpers for bench.py (the reference's own code on one host core) -----------------
// Runs nSteps sweeps through MonteCarlo::run (output disabled except the final mandatory
// sample) and returns wall seconds; *moves receives the number of trial moves the reference
// counted (simulation_step_time delta, MC.cpp:40,47).
double ref_time_mc_run(void *h, size_t nSteps, size_t fUpdateCell, long long *moves) {
    RefMC *m = static_cast<RefMC *>(h);
    long long t0 = m->mc->simulation_step_time;
    auto a = std::chrono::steady_clock::now();
    m->mc->run(nSteps, (size_t)1 << 60, fUpdateCell);
    auto b = std::chrono::steady_clock::now();
    *moves = m->mc->simulation_step_time - t0;
    return std::chrono::duration<double>(b - a).count();
}
// The reference's move functions alone (no logging, no energy re-sync): per sweep N calls of
// displacementMove_cell_list_dE (via stepCellList), one grandCanonicalMove when gcmc != 0, and a
// cell rebuild every fUpdateCell sweeps — the most favourable way to time the reference's
// "attempted moves per second".  Returns wall seconds; *moves = trials performed.
double ref_time_moves(void *h, size_t nSweeps, size_t fUpdateCell, int gcmc, long long *moves) {
    RefMC *m = static_cast<RefMC *>(h);
    long long done = 0;
    auto a = std::chrono::steady_clock::now();
    for (size_t s = 0; s < nSweeps; ++s) {
        size_t N = m->particles.size();
        for (size_t i = 0; i < N; ++i) m->mc->stepCellList();
        done += (long long)N;
        if (gcmc) {
            m->mc->grandCanonicalMove();
            done += 1;
        }
        if (fUpdateCell && s % fUpdateCell == 0) m->mc->updateCellList();
    }
    auto b = std::chrono::steady_clock::now();
    *moves = done;
    return std::chrono::duration<double>(b - a).count();
}
// The reference's sampler without its stale-cell-list shortcut: the cell list is rebuilt before EVERY trial (N x
// [updateCellList + stepCellList], then — for GCMC — a rebuild, one grandCanonicalMove and another rebuild), so that
// every dE is evaluated on a list that knows every particle in its current cell.  Same moves, same acceptance rules,
// same RNG stream as MonteCarlo::run; only the neighbour search is exact (O(N) per trial: small systems only; the
// reference's own stepBruteForce recomputes the TOTAL energy twice per trial, O(N^2), and is no alternative).  Used by
// tools/stat_parity.py where the stale list of the stock loop lets overlaps through (hard cores).
// Returns the particle count; *energy = running energy.
int ref_mc_run_exact(void *h, size_t nSweeps, int gcmc, double *energy) {
    RefMC *m = static_cast<RefMC *>(h);
    for (size_t s = 0; s < nSweeps; ++s) {
        const size_t N = m->particles.size();
        for (size_t i = 0; i < N; ++i) {
            m->mc->updateCellList();
            m->mc->stepCellList();
        }
        if (gcmc) {
            m->mc->updateCellList();
            m->mc->grandCanonicalMove();
            m->mc->updateCellList();
        }
    }
    if (energy) *energy = m->mc->getEnergy();
    return static_cast<int>(m->particles.size());
}
// One computeTotalEnergyCellList + computeTotalVirialCellList pass; returns wall seconds.
double ref_time_energy_virial(double T, int type, double rcut, double f, double alpha, double A0, double kappa,
                              double Lx, double Ly, const double *xy, int N, double *U, double *W) {
    auto calc = make_calc({T, 0.0, type, rcut, 0.0, f, alpha, A0, kappa});
    SimulationBox box(Lx, Ly);
    auto P = to_particles(xy, N);
    auto a = std::chrono::steady_clock::now();
    *U = calc.computeTotalEnergyCellList(P, box);
    *W = calc.computeTotalVirialCellList(P, box);
    auto b = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(b - a).count();
}

// The reference's position logs (logging.cpp:30-79) for `nframes` frames lying behind each other in xy (counts[f]
// particles each): dump != 0 -> logPositions_dump with timesteps[f], else logPositions_xyz.  Golden side of
// tests/test_traj_logger.py::test_serial_position_logs.
int ref_log_positions(const char *path_pos, const char *path_data, int dump, double Lx, double Ly, const double *xy,
                      const int *counts, int nframes, const int *timesteps) {
    try {
        Logging log(path_pos, path_data);
        SimulationBox box(Lx, Ly);
        size_t at = 0;
        for (int f = 0; f < nframes; ++f) {
            std::vector<Particle> P = to_particles(xy + 2 * at, counts[f]);
            at += (size_t)counts[f];
            if (dump) log.logPositions_dump(P, box, timesteps[f]);
            else log.logPositions_xyz(P, box, 0.0);
        }
        log.close();
        return 0;
    } catch (...) {
        return 1;
    }
}

}  // extern "C"
