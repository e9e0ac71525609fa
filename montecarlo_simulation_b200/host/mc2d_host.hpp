// mc2d_host.hpp — C++ host shims that keep the reference's class interfaces for the hot path and
// forward to the C ABI (include/mc2d.h -> libmc2d.so, CUDA sm_100a).
//
// Same class names, constructor signatures, method names and argument meaning as
// ReyhanehFarimani/MonteCarlo_simulation serial_src/, so a driver written against the reference
// (serial_src/main.cpp) or one of its unit tests compiles against this header unchanged:
//
//   Particle, SimulationBox              serial_src/initial.h:10-110
//   PotentialType, selectPotentialType,
//   computePairPotential/Force           serial_src/potential.h:11-164 (host scalar helpers)
//   Input                                serial_src/input.h:14-62
//   RNG                                  serial_src/rng.h:17-54
//   CellList                             serial_src/cell_list.h:13-75          -> mc2d_calc_*
//   ThermodynamicCalculator              serial_src/thermodynamic_calculator.h:33-203 -> mc2d_calc_*
//   Logging                              serial_src/logging.h:14-58
//   Ensemble, MonteCarlo                 serial_src/MC.h:14-98                 -> mc2d_upload / mc2d_run_sweeps
//   MonteCarloNVT_Slabs                  mpi_src/MC_parallel.h:24-88 (MonteCarloNVT_MPI) -> slabs over NCCL
//
// Nothing here computes pair sums on the CPU: every energy, virial, neighbour query and Monte
// Carlo move goes to the GPU; a missing device surfaces as std::runtime_error.
#pragma once
#include <cstddef>
#include <cstdint>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mc2d.h"

// ---- serial_src/initial.h ------------------------------------------------------------------
class Particle {
public:
    double x, y;
    Particle(double x_init = 0.0, double y_init = 0.0) : x(x_init), y(y_init) {}
    void updatePosition(double dx, double dy) { x += dx; y += dy; }
};
static_assert(sizeof(Particle) == 2 * sizeof(double), "Particle must be a packed {double x, y}");

class SimulationBox {
    double Lx, Ly, invLx, invLy, volume;

public:
    SimulationBox(double Lx_, double Ly_);
    void applyPBC(Particle &p) const;
    double minimumImageDistance(const Particle &p1, const Particle &p2) const;
    double minimumImageDistanceSquared(const Particle &p1, const Particle &p2) const;
    double getLx() const { return Lx; }
    double getLy() const { return Ly; }
    double getV() const { return volume; }
    void setLx(double lx);
    void setLy(double ly);
    void setV(double v);
    void recenter(Particle &p, double lx_old, double ly_old);
};

void initializeParticles(std::vector<Particle> &particles, const SimulationBox &box, int N, bool random = true,
                         unsigned int seed = 0);
void initializeParticles_from_file(std::vector<Particle> &particles, const SimulationBox &box, int N,
                                   const std::string &filename_data);

// ---- serial_src/potential.h ------------------------------------------------------------------
enum class PotentialType { LennardJones, WCA, Yukawa, AthermalStar, ThermalStar, Ideal };
PotentialType selectPotentialType(const std::string &potentialName);
double computePairPotential(double r2, PotentialType type, float f_prime, float f_d_prime, float kappa, float alpha);
double computePairForce(double r2, PotentialType type, float f_prime, float f_d_prime, float kappa, float alpha);

// ---- serial_src/input.h ------------------------------------------------------------------------
class Input {
public:
    explicit Input(const std::string &filename);
    void parseCommandLineArgs(int argc, char *argv[]);
    double getConstant(const std::string &key) const;
    std::string getFilename(const std::string &key) const;
    void setConstant(const std::string &key, double value);
    void setFilename(const std::string &key, const std::string &value);

private:
    std::map<std::string, double> constants;
    std::map<std::string, std::string> filenames;
    void parseInputFile(const std::string &filename);
    void store(const std::string &key, const std::string &value);
};

// ---- serial_src/rng.h --------------------------------------------------------------------------
class RNG {
public:
    explicit RNG(unsigned seed);
    double uniform01();
    double uniform(double a, double b);
    int uniformInt(int low, int high);

private:
    std::mt19937 engine_;
    std::uniform_real_distribution<double> dist01_;
};

// ---- GPU context shared by the shims ----------------------------------------------------------------
namespace mc2d {
struct PotentialSpec {
    PotentialType type = PotentialType::LennardJones;
    double f_prime = 0, f_d_prime = 0, kappa = 0, alpha = 0;  // as ThermodynamicCalculator derives them
};
// RAII owner of an mc2d_ctx; throws std::runtime_error with mc2d_last_error on failure.
class Context {
public:
    Context(double Lx, double Ly, double rcut, const PotentialSpec &pot, double T, double z, double delta,
            std::uint64_t seed, int device = 0, int rank = 0, int nranks = 1, int cell_capacity = 0,
            double cell_margin = 0.0);
    ~Context();
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    mc2d_ctx *get() const { return ctx_; }
    void check(int status) const;
    double Lx() const { return Lx_; }
    double Ly() const { return Ly_; }

private:
    mc2d_ctx *ctx_ = nullptr;
    double Lx_, Ly_;
};
}  // namespace mc2d

// ---- serial_src/cell_list.h --------------------------------------------------------------------------
class CellList {
public:
    CellList(SimulationBox &box, double rcut);
    void adjust_box(SimulationBox &box);
    void build(const std::vector<Particle> &particles);
    std::vector<std::pair<int, double>> getNeighbors(int idx, const std::vector<Particle> &particles) const;
    std::vector<std::pair<int, double>> getNeighbors2(const Particle &p, const std::vector<Particle> &particles) const;
    // parity hooks beyond the reference surface
    std::vector<int> cellIds(const std::vector<Particle> &particles) const;
    std::vector<int> neighborCounts(const std::vector<Particle> &particles) const;
    int nx() const { return nx_; }
    int ny() const { return ny_; }

private:
    SimulationBox &box_;
    double rcut_;
    int nx_ = 1, ny_ = 1;
    mutable std::unique_ptr<mc2d::Context> gpu_;
    mc2d::Context &gpu() const;
};

// ---- serial_src/thermodynamic_calculator.h ---------------------------------------------------------------
class ThermodynamicCalculator {
public:
    ThermodynamicCalculator(double temperature, double press, PotentialType potentialType, double rcut,
                            double mu = 0.0, double f = 0.0, double alpha = 0.0, double A_0 = 0.0, double kappa = 0.0);
    size_t getNumParticles(const std::vector<Particle> &particles) const { return particles.size(); }
    double getTemperature() const { return temperature_; }
    double getPressure() const { return press_; }
    double getActivity() const { return mu_; }
    double getVolume(const SimulationBox &box) const { return box.getV(); }
    double computeDensity(const std::vector<Particle> &particles, const SimulationBox &box) const;
    // brute force over all pairs (GPU, O(N^2); the reference's test truth)
    double computeTotalEnergy(const std::vector<Particle> &particles, const SimulationBox &box) const;
    double computeTotalVirial(const std::vector<Particle> &particles, const SimulationBox &box) const;
    double computePressure(const std::vector<Particle> &particles, const SimulationBox &box) const;
    // sum of pair energies over a neighbour list the caller already holds (thermodynamic_calculator.h:103-121)
    template <typename NeighborList>
    double computeLocalEnergy(size_t, const std::vector<Particle> &, const SimulationBox &,
                              const NeighborList &neighbors) const {
        double U = 0.0;
        for (const auto &pr : neighbors)
            U += computePairPotential(pr.second, potentialType_, (float)f_prime_, (float)f_d_prime_, (float)kappa_,
                                      (float)alpha_);
        return U;
    }
    double computeTailCorrectionEnergy2D(const std::vector<Particle> &particles, const SimulationBox &box) const;
    double computeTailCorrectionPressure2D(const std::vector<Particle> &particles, const SimulationBox &box) const;
    // cell-list reductions (GPU, one pass gives both U and W)
    double computeTotalEnergyCellList(const std::vector<Particle> &particles, SimulationBox &box) const;
    double computeTotalVirialCellList(const std::vector<Particle> &particles, SimulationBox &box) const;
    double computePressureCellList(const std::vector<Particle> &particles, SimulationBox &box) const;
    // both at once (what Logging needs per sample)
    void computeEnergyVirialCellList(const std::vector<Particle> &particles, const SimulationBox &box, double &U,
                                     double &W) const;

    double getCutoff() const { return rcut_; }
    mc2d::PotentialSpec potentialSpec() const { return {potentialType_, f_prime_, f_d_prime_, kappa_, alpha_}; }

private:
    double temperature_, press_, mu_, rcut_, r2cut_;
    PotentialType potentialType_;
    double f_prime_, alpha_, f_d_prime_, kappa_;
    mutable std::unique_ptr<mc2d::Context> gpu_;
    mc2d::Context &gpu(const SimulationBox &box) const;
};

// ---- serial_src/logging.h -------------------------------------------------------------------------------------
class Logging {
public:
    Logging(const std::string &filename_position, const std::string &filename_data);
    ~Logging();
    void logPositions_xyz(const std::vector<Particle> &particles, const SimulationBox &box, double r2cut);
    void logPositions_dump(const std::vector<Particle> &particles, const SimulationBox &box, int timestep);
    void logSimulationData(const std::vector<Particle> &particles, SimulationBox &box,
                           const ThermodynamicCalculator &cal, int timestep);
    void close();

private:
    std::ofstream outFile_position, outFile_data;
    std::string filename_position, filename_data;
};

// ---- serial_src/MC.h ----------------------------------------------------------------------------------------------
enum class Ensemble { NVT, GCMC, NPT, Gibbs };

class MonteCarlo {
public:
    MonteCarlo(SimulationBox &box, std::vector<Particle> &particles, ThermodynamicCalculator &calc, double rcut,
               double delta, double delta_V, Ensemble ensemble, Logging &logger, RNG &rng);
    ~MonteCarlo();
    // nSteps sweeps; observables every fOutputStep sweeps and at the end; energy re-sync every fUpdateCell
    int run(size_t nSteps, size_t fOutputStep, size_t fUpdateCell);
    bool stepCellList();    // test hook: ONE SWEEP of displacement trials (the reference does one trial)
    bool stepBruteForce();  // same path (the GPU has no separate brute-force move)
    double getEnergy() const;
    const std::vector<Particle> &getParticles() const;  // synchronises the caller's vector with the device
    void updateCellList();                               // no-op: device cells are always exact
    bool nptMove();  // one volume move (MC.cpp:153-188): ln V step of width delta_V, isotropic rescaling
    const SimulationBox &getBox() const { return box_; }
    // GPU-specific knobs
    void setGcDrawsPerCell(int n) { gc_per_cell_ = n; }
    mc2d_stats stats() const;
    double acceptanceRatio() const;

private:
    SimulationBox box_;
    std::vector<Particle> &particles_;
    ThermodynamicCalculator &calc_;
    Logging &logger_;
    Ensemble ensemble_;
    RNG &rng_;
    double rcut_, delta_, delta_V_;
    std::uint64_t seed_ = 0;
    int regrids_ = 0;  // contexts re-created because the box left the range one cell grid can serve
    int gc_per_cell_ = 1;
    long long simulation_step_time = 0;
    std::unique_ptr<mc2d::Context> gpu_;
    mutable bool host_stale_ = false;
    void sync_host() const;
    void recordObservables();
    void regrid(const SimulationBox &nb, std::unique_ptr<mc2d::Context> &out, std::vector<Particle> &scaled);
};

// ---- serial_src/GibbsMC.h ------------------------------------------------------------------------------------------
// Two boxes, each resident on the GPU in its own context.  A step is, as in GibbsMC.cpp:207-270, with probability
// 1/2 a displacement sweep of both boxes, 1/4 one particle exchange, 1/4 one coupled volume change; the
// exchange uses mc2d_probe_point / mc2d_probe_particle / mc2d_apply_probe, the volume change mc2d_rescale_*.
class GibbsMonteCarlo {
public:
    GibbsMonteCarlo(SimulationBox &box1, SimulationBox &box2, std::vector<Particle> particles1,
                    std::vector<Particle> particles2, ThermodynamicCalculator &calc1, ThermodynamicCalculator &calc2,
                    double rcut, double delta, double delta_V, Logging &logger1, Logging &logger2, RNG &rng);
    ~GibbsMonteCarlo();
    int run(size_t nSteps, size_t fOutputStep, size_t fUpdateCell);
    int equlibrateNVT(size_t nSteps, size_t fOutputStep, size_t fUpdateCell);
    int equlibrateNPT(size_t nSteps, size_t fOutputStep, size_t fUpdateCell);
    int equlibratemuVT(size_t nSteps, size_t fOutputStep, size_t fUpdateCell);
    double w_1 = 0, w_2 = 0;  // Widom accumulators of the exchange move (GibbsMC.cpp:352,375)
    // false (default): particle exchange accepts with the detailed-balance ratio N_F V_T / ((N_T + 1) V_F);
    // true: with the reference's own expression V_F (N_T + 1) / (V_T (N_F + 1)) (GibbsMC.cpp:359,382) — for A/B runs
    // against the reference's output only: with it an ideal gas collapses into one box.
    bool reference_exchange_ratio = false;
    // state, for tests and drivers
    const SimulationBox &getBox(int which) const;
    long long getNumParticles(int which) const;
    double getEnergy(int which) const;
    const std::vector<Particle> &getParticles(int which);
    bool particle_exchange();
    bool volume_change();

private:
    struct Side;
    std::unique_ptr<Side> s_[2];
    Logging *logger_[2];
    RNG &rng_;
    double rcut_, delta_, delta_V_, beta_;
    long long simulation_step_time = 0;
    enum class Loop { NVT, NPT, muVT, Gibbs };
    int loop(Loop kind, size_t nSteps, size_t fOutputStep, size_t fUpdateCell);
    void sweep_both(long long &acc, long long &tot);
    void recordObservables();
};

// ---- mpi_src/MC_parallel.h (MonteCarloNVT_MPI) on GPU slabs ---------------------------------------------------------
// One object per rank / GPU.  The communicator bootstrap is a callback that broadcasts a byte buffer
// from rank 0 (MPI_Bcast in the reference's world), so this header needs no MPI.
class MonteCarloNVT_Slabs {
public:
    struct Params {
        double rcut;
        double delta;
        int out_every{0};
        int gc_per_cell{0};  // 0 = NVT like the reference's MPI driver; > 0 adds GCMC (the reference cannot)
    };
    using BcastFn = std::function<void(void *buf, std::size_t bytes, int root)>;
    MonteCarloNVT_Slabs(int rank, int nranks, int device, SimulationBox &box, ThermodynamicCalculator &thermo,
                        std::vector<Particle> &owned, std::uint64_t seed, const Params &p, const BcastFn &bcast);
    ~MonteCarloNVT_Slabs();
    double run(std::size_t nsweeps, int start_timestep = 0);  // returns the GLOBAL acceptance ratio
    double totalEnergy();  // global (all-reduced), thermodynamic_calculator_parallel.cpp:56-94
    double totalVirial();
    double pressure();
    long long numParticlesGlobal();
    void download(std::vector<Particle> &owned);

private:
    int rank_, nranks_;
    SimulationBox &box_;
    ThermodynamicCalculator &thermo_;
    std::vector<Particle> &owned_;
    Params p_;
    std::unique_ptr<mc2d::Context> gpu_;
};
