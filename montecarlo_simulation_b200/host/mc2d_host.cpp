// mc2d_host.cpp — implementation of the reference-interface shims over the C ABI (see
// mc2d_host.hpp).  Host code only does what the reference's host code does around the hot path
// (parsing, logging, bookkeeping); every pair sum and every Monte Carlo move runs in libmc2d.so.
#include "mc2d_host.hpp"

#include <climits>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <sstream>

// =================================================================================================
// SimulationBox, initialisers (serial_src/initial.cpp)
// =================================================================================================
SimulationBox::SimulationBox(double Lx_, double Ly_)
    : Lx(Lx_), Ly(Ly_), invLx(1.0 / Lx_), invLy(1.0 / Ly_), volume(Lx_ * Ly_) {}

void SimulationBox::applyPBC(Particle &p) const {  // initial.cpp:22-25
    p.x -= Lx * std::floor(p.x * invLx);
    p.y -= Ly * std::floor(p.y * invLy);
}
double SimulationBox::minimumImageDistanceSquared(const Particle &a, const Particle &b) const {  // initial.cpp:43-51
    double dx = a.x - b.x, dy = a.y - b.y;
    dx -= Lx * std::round(dx * invLx);
    dy -= Ly * std::round(dy * invLy);
    return dx * dx + dy * dy;
}
double SimulationBox::minimumImageDistance(const Particle &a, const Particle &b) const {
    return std::sqrt(minimumImageDistanceSquared(a, b));
}
void SimulationBox::setLx(double lx) { Lx = lx; invLx = 1 / Lx; volume = Lx * Ly; }
void SimulationBox::setLy(double ly) { Ly = ly; invLy = 1 / Ly; volume = Lx * Ly; }
void SimulationBox::setV(double v) {  // keeps the aspect ratio, initial.cpp:76-83
    const double ratio = Lx / Ly;
    Lx = std::sqrt(ratio * v);
    Ly = std::sqrt(v / ratio);
    invLx = 1 / Lx;
    invLy = 1 / Ly;
    volume = Lx * Ly;
}
void SimulationBox::recenter(Particle &p, double lx_old, double ly_old) {
    p.x *= Lx / lx_old;
    p.y *= Ly / ly_old;
}

void initializeParticles(std::vector<Particle> &particles, const SimulationBox &box, int N, bool random,
                         unsigned int seed) {  // initial.cpp:86-123
    particles.clear();
    particles.reserve(N > 0 ? N : 0);
    if (random) {
        srand(seed == 0 ? static_cast<unsigned>(time(nullptr)) : seed + 10);
        for (int i = 0; i < N; ++i) {
            const double x = static_cast<double>(rand()) / RAND_MAX * box.getLx();
            const double y = static_cast<double>(rand()) / RAND_MAX * box.getLy();
            particles.emplace_back(x, y);
        }
        return;
    }
    const int side = static_cast<int>(std::sqrt(N));
    const double sx = box.getLx() / side, sy = box.getLy() / side;
    for (int i = 0; i < side; ++i)
        for (int j = 0; j < side; ++j) particles.emplace_back(i * sx, j * sy);
}

void initializeParticles_from_file(std::vector<Particle> &particles, const SimulationBox &box, int N,
                                   const std::string &filename_data) {  // initial.cpp:125-168
    particles.clear();
    particles.reserve(N > 0 ? N : 0);
    std::ifstream in(filename_data);
    if (!in.is_open()) {
        std::cerr << "Error: Could not open file " << filename_data << std::endl;
        return;
    }
    std::string line;
    std::getline(in, line);  // particle count
    std::getline(in, line);  // comment
    int taken = 0;
    while (taken < N && std::getline(in, line)) {
        std::istringstream iss(line);
        std::string type;
        double x, y;
        if (!(iss >> type >> x >> y)) {
            std::cerr << "Error reading line: " << line << std::endl;
            continue;
        }
        if (x >= 0 && x <= box.getLx() && y >= 0 && y <= box.getLy()) {
            particles.emplace_back(x, y);
            ++taken;
        } else {
            std::cerr << "Warning: Particle out of bounds (" << x << ", " << y << ")" << std::endl;
        }
    }
}

// =================================================================================================
// pair functions as host scalars (serial_src/potential.cpp) — used by computeLocalEnergy on a
// caller-held neighbour list and by callers of the free functions; the GPU has its own copies
// =================================================================================================
PotentialType selectPotentialType(const std::string &name) {
    static const std::pair<const char *, PotentialType> table[] = {
        {"LennardJones", PotentialType::LennardJones}, {"WCA", PotentialType::WCA},
        {"Yukawa", PotentialType::Yukawa},             {"AthermalStar", PotentialType::AthermalStar},
        {"ThermalStar", PotentialType::ThermalStar},   {"Ideal", PotentialType::Ideal}};
    for (const auto &e : table)
        if (name == e.first) return e.second;
    throw std::invalid_argument("Unknown potential type: " + name);
}

namespace {
struct UW {
    double u, w;
};
// one evaluation gives both; expressions keep the reference's operand order and float operands
UW pair_uw(double r2, PotentialType t, float fp, float fdp, float kappa, float alpha) {
    switch (t) {
    case PotentialType::LennardJones: {
        const double r2inv = 1.0 / r2, r6inv = r2inv * r2inv * r2inv;
        return {4.0 * r6inv * (r6inv - 1.0), 48.0 * r6inv * (r6inv - 0.5)};
    }
    case PotentialType::WCA: {
        if (!(r2 < 1.2599)) return {0.0, 0.0};  // the literal cutoff of potential.cpp:62
        const double r2inv = 1.0 / r2, r6inv = r2inv * r2inv * r2inv;
        return {4.0 * r6inv * (r6inv - 1.0) + 1.0, 48.0 * r6inv * (r6inv - 0.5)};
    }
    case PotentialType::Yukawa: {  // epsilon = kappa = 1 hard-coded (potential.cpp:95-113)
        const double r = std::sqrt(r2);
        return {1.0 * std::exp(-1.0 * r) / r, 1.0 * (1.0 + 1.0 / r) * std::exp(-1.0 * r)};
    }
    case PotentialType::AthermalStar: {
        if (r2 < 1) return {fp * (-std::log(std::sqrt(r2)) + alpha), fp};
        return {fp * (alpha * std::exp((1 - r2) / 2 / alpha)), fp * r2 * std::exp((1 - r2) / 2 / alpha)};
    }
    case PotentialType::ThermalStar: {
        const double r = std::sqrt(r2);
        const double e2 = fdp * std::exp(-kappa * r);
        const double f = fdp * kappa * std::exp(-kappa * r) * r;  // float*float first, as the reference
        if (r2 < 1) return {fp * (-std::log(r) + alpha) - e2, fp - f};
        return {fp * (alpha * std::exp((1 - r2) / 2 / alpha)) - e2, fp * r2 * std::exp((1 - r2) / 2 / alpha) - f};
    }
    case PotentialType::Ideal: return {0.0, 0.0};
    }
    throw std::invalid_argument("Unknown potential type");
}
}  // namespace

double computePairPotential(double r2, PotentialType t, float fp, float fdp, float kappa, float alpha) {
    return pair_uw(r2, t, fp, fdp, kappa, alpha).u;
}
double computePairForce(double r2, PotentialType t, float fp, float fdp, float kappa, float alpha) {
    return pair_uw(r2, t, fp, fdp, kappa, alpha).w;
}

// =================================================================================================
// Input (serial_src/input.cpp)
// =================================================================================================
Input::Input(const std::string &filename) { parseInputFile(filename); }

void Input::store(const std::string &key, const std::string &value) {
    try {  // numeric -> constants, anything std::stod rejects -> filenames (input.cpp:27-32)
        constants[key] = std::stod(value);
    } catch (std::invalid_argument &) {
        filenames[key] = value;
    }
}
void Input::parseInputFile(const std::string &filename) {
    std::ifstream in(filename);
    if (!in) throw std::runtime_error("Could not open input file: " + filename);
    std::string line, key, value;
    while (std::getline(in, line)) {
        std::istringstream iss(line);
        if (iss >> key >> value) store(key, value);
    }
}
void Input::parseCommandLineArgs(int argc, char *argv[]) {
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        const auto eq = arg.find('=');
        if (eq != std::string::npos) store(arg.substr(0, eq), arg.substr(eq + 1));
    }
}
double Input::getConstant(const std::string &key) const {
    const auto it = constants.find(key);
    return it == constants.end() ? 0.0 : it->second;  // a missing key is 0.0 (input.cpp:56-63)
}
std::string Input::getFilename(const std::string &key) const {
    const auto it = filenames.find(key);
    return it == filenames.end() ? std::string() : it->second;
}
void Input::setConstant(const std::string &key, double value) { constants[key] = value; }
void Input::setFilename(const std::string &key, const std::string &value) { filenames[key] = value; }

// =================================================================================================
// RNG (serial_src/rng.cpp) — host-side only: initial conditions and the Philox seed
// =================================================================================================
RNG::RNG(unsigned seed) : engine_(seed), dist01_(0.0, 1.0) {}
double RNG::uniform01() { return dist01_(engine_); }
double RNG::uniform(double a, double b) { return a + (b - a) * uniform01(); }
int RNG::uniformInt(int low, int high) {
    std::uniform_int_distribution<int> dist(low, high);
    return dist(engine_);
}

// =================================================================================================
// mc2d::Context
// =================================================================================================
namespace mc2d {
Context::Context(double Lx, double Ly, double rcut, const PotentialSpec &pot, double T, double z, double delta,
                 std::uint64_t seed, int device, int rank, int nranks, int cell_capacity, double cell_margin)
    : Lx_(Lx), Ly_(Ly) {
    mc2d_params p{};
    p.Lx = Lx;
    p.Ly = Ly;
    p.rcut = rcut;
    p.potential = static_cast<int>(pot.type);  // same order as potential.h:11-18
    p.f_prime = static_cast<float>(pot.f_prime);
    p.f_d_prime = static_cast<float>(pot.f_d_prime);
    p.kappa = static_cast<float>(pot.kappa);
    p.alpha = static_cast<float>(pot.alpha);
    p.temperature = T;
    p.activity = z;
    p.delta = delta;
    p.seed = seed;
    p.device = device;
    p.rank = rank;
    p.nranks = nranks;
    p.cell_capacity = cell_capacity;
    p.cell_margin = cell_margin;
    const int rc = mc2d_create(&p, &ctx_);
    if (rc != MC2D_OK) {
        const std::string msg = std::string(mc2d_status_string(rc)) + ": " + (ctx_ ? mc2d_last_error(ctx_) : "");
        if (ctx_) mc2d_destroy(ctx_);
        ctx_ = nullptr;
        if (rc == MC2D_ERR_INVALID) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }
}
Context::~Context() {
    if (ctx_) mc2d_destroy(ctx_);
}
void Context::check(int status) const {
    if (status == MC2D_OK) return;
    const std::string msg = std::string(mc2d_status_string(status)) + ": " + mc2d_last_error(ctx_);
    if (status == MC2D_ERR_INVALID) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
}  // namespace mc2d

static const double *raw(const std::vector<Particle> &p) { return reinterpret_cast<const double *>(p.data()); }

// =================================================================================================
// CellList (serial_src/cell_list.cpp) on the GPU
// =================================================================================================
CellList::CellList(SimulationBox &box, double rcut) : box_(box), rcut_(rcut) { adjust_box(box); }

void CellList::adjust_box(SimulationBox &box) {
    box_ = box;
    mc2d_reference_grid(box_.getLx(), box_.getLy(), rcut_, &nx_, &ny_, nullptr, nullptr);
    gpu_.reset();
}
mc2d::Context &CellList::gpu() const {
    if (!gpu_ || gpu_->Lx() != box_.getLx() || gpu_->Ly() != box_.getLy()) {
        mc2d::PotentialSpec spec;
        spec.type = PotentialType::Ideal;  // neighbour queries do not depend on the potential
        gpu_.reset(new mc2d::Context(box_.getLx(), box_.getLy(), rcut_, spec, 1.0, 0.0, 0.0, 0));
    }
    return *gpu_;
}
// The device list is rebuilt from the caller's vector on every query (the reference's own
// calculators do the same, thermodynamic_calculator.cpp:198-199), so build() has nothing to keep.
void CellList::build(const std::vector<Particle> &) {}

static std::vector<std::pair<int, double>> neighbor_query(mc2d::Context &g, const std::vector<Particle> &P,
                                                          long long idx, double x, double y) {
    int cap = 256, count = 0;
    std::vector<int> j;
    std::vector<double> r2;
    for (;;) {
        j.resize(cap);
        r2.resize(cap);
        g.check(mc2d_calc_neighbor_list(g.get(), raw(P), (long long)P.size(), idx, x, y, j.data(), r2.data(), cap, &count));
        if (count <= cap) break;
        cap = count;
    }
    std::vector<std::pair<int, double>> out;
    out.reserve(count);
    for (int k = 0; k < count; ++k) out.emplace_back(j[k], r2[k]);
    return out;
}
std::vector<std::pair<int, double>> CellList::getNeighbors(int idx, const std::vector<Particle> &particles) const {
    return neighbor_query(gpu(), particles, idx, 0.0, 0.0);
}
std::vector<std::pair<int, double>> CellList::getNeighbors2(const Particle &p,
                                                            const std::vector<Particle> &particles) const {
    return neighbor_query(gpu(), particles, -1, p.x, p.y);
}
std::vector<int> CellList::cellIds(const std::vector<Particle> &particles) const {
    std::vector<int> out(particles.size());
    gpu().check(mc2d_calc_cell_ids(gpu().get(), raw(particles), (long long)particles.size(), out.data()));
    return out;
}
std::vector<int> CellList::neighborCounts(const std::vector<Particle> &particles) const {
    std::vector<int> out(particles.size());
    gpu().check(mc2d_calc_neighbor_counts(gpu().get(), raw(particles), (long long)particles.size(), out.data(), nullptr));
    return out;
}

// =================================================================================================
// ThermodynamicCalculator (serial_src/thermodynamic_calculator.cpp) on the GPU
// =================================================================================================
ThermodynamicCalculator::ThermodynamicCalculator(double temperature, double press, PotentialType potentialType,
                                                 double rcut, double mu, double f, double alpha, double A_0,
                                                 double kappa)
    : temperature_(temperature), press_(press), mu_(mu), rcut_(rcut), r2cut_(rcut * rcut),
      potentialType_(potentialType), f_prime_((2.0 + 9.0 * f * f) / 24.0), alpha_(alpha),
      f_d_prime_(A_0 * f * f / kappa), kappa_(kappa) {
    if (!(rcut_ >= 0.0)) throw std::invalid_argument("Cutoff distance must be non-negative");
}

mc2d::Context &ThermodynamicCalculator::gpu(const SimulationBox &box) const {
    if (!gpu_ || gpu_->Lx() != box.getLx() || gpu_->Ly() != box.getLy())
        gpu_.reset(new mc2d::Context(box.getLx(), box.getLy(), rcut_, potentialSpec(),
                                     temperature_ > 0 ? temperature_ : 1.0, mu_ > 0 ? mu_ : 0.0, 0.0, 0));
    return *gpu_;
}
double ThermodynamicCalculator::computeDensity(const std::vector<Particle> &particles, const SimulationBox &box) const {
    return static_cast<double>(particles.size()) / box.getV();
}
void ThermodynamicCalculator::computeEnergyVirialCellList(const std::vector<Particle> &particles,
                                                          const SimulationBox &box, double &U, double &W) const {
    mc2d::Context &g = gpu(box);
    g.check(mc2d_calc_energy_virial(g.get(), raw(particles), (long long)particles.size(), &U, &W));
}
double ThermodynamicCalculator::computeTotalEnergyCellList(const std::vector<Particle> &particles,
                                                           SimulationBox &box) const {
    double U, W;
    computeEnergyVirialCellList(particles, box, U, W);
    return U;
}
double ThermodynamicCalculator::computeTotalVirialCellList(const std::vector<Particle> &particles,
                                                           SimulationBox &box) const {
    double U, W;
    computeEnergyVirialCellList(particles, box, U, W);
    return W;
}
double ThermodynamicCalculator::computePressureCellList(const std::vector<Particle> &particles,
                                                        SimulationBox &box) const {
    const double V = box.getV();  // P = rho T + W / (2V), thermodynamic_calculator.cpp:235-242
    return computeDensity(particles, box) * temperature_ + computeTotalVirialCellList(particles, box) / (2.0 * V);
}
double ThermodynamicCalculator::computeTotalEnergy(const std::vector<Particle> &particles,
                                                   const SimulationBox &box) const {
    double U, W;
    mc2d::Context &g = gpu(box);
    g.check(mc2d_calc_energy_virial_bruteforce(g.get(), raw(particles), (long long)particles.size(), &U, &W));
    return U;
}
double ThermodynamicCalculator::computeTotalVirial(const std::vector<Particle> &particles,
                                                   const SimulationBox &box) const {
    double U, W;
    mc2d::Context &g = gpu(box);
    g.check(mc2d_calc_energy_virial_bruteforce(g.get(), raw(particles), (long long)particles.size(), &U, &W));
    return W;
}
double ThermodynamicCalculator::computePressure(const std::vector<Particle> &particles,
                                                const SimulationBox &box) const {
    const double V = box.getV();
    return computeDensity(particles, box) * temperature_ + computeTotalVirial(particles, box) / (2.0 * V);
}
// 2D tail corrections (thermodynamic_calculator.cpp:114-190; no caller in the reference): closed forms of
// pi*rho*Int_rc^inf u(r) r dr and (pi*rho^2/2)*Int_rc^inf (r.f) r dr
double ThermodynamicCalculator::computeTailCorrectionEnergy2D(const std::vector<Particle> &particles,
                                                              const SimulationBox &box) const {
    const double pref = M_PI * computeDensity(particles, box);
    const double s2 = 1.0 / r2cut_, s4 = s2 * s2, s10 = s4 * s4 * s2;
    const double star = -std::exp((1.0 - r2cut_) / (2.0 * alpha_)) * alpha_ * alpha_ * f_prime_;
    switch (potentialType_) {
    case PotentialType::LennardJones: return pref * 4.0 * (s4 / 4.0 - s10 / 10.0);
    case PotentialType::AthermalStar: return pref * star;
    case PotentialType::ThermalStar: {
        const double r = std::sqrt(r2cut_);
        return pref * (star + f_d_prime_ / (kappa_ * kappa_) * std::exp(-kappa_ * r) * (kappa_ * r + 1.0));
    }
    default: return 0.0;
    }
}
double ThermodynamicCalculator::computeTailCorrectionPressure2D(const std::vector<Particle> &particles,
                                                                const SimulationBox &box) const {
    const double rho = computeDensity(particles, box), pref = M_PI * rho * rho / 2.0;
    const double s2 = 1.0 / r2cut_, s4 = s2 * s2, s10 = s4 * s4 * s2;
    const double star = f_prime_ * alpha_ * (r2cut_ + 2.0 * alpha_) *
                        std::exp(1.0 / (2.0 * alpha_) - r2cut_ / (2.0 * alpha_));
    switch (potentialType_) {
    case PotentialType::LennardJones: return pref * 48.0 * (s4 / 8.0 - s10 / 10.0);
    case PotentialType::AthermalStar: return pref * star;
    case PotentialType::ThermalStar: {
        const double r = std::sqrt(r2cut_);
        return pref * (star - f_d_prime_ / (kappa_ * kappa_) * std::exp(-kappa_ * r) *
                                  (kappa_ * kappa_ * r2cut_ + 2.0 * kappa_ * r + 2.0));
    }
    default: return 0.0;
    }
}

// =================================================================================================
// Logging (serial_src/logging.cpp) — formats kept byte for byte: PostProcessing parses them
// =================================================================================================
Logging::Logging(const std::string &fp, const std::string &fd) : filename_position(fp), filename_data(fd) {
    outFile_position.open(fp);
    outFile_data.open(fd);
    if (!outFile_position) throw std::runtime_error("Could not open file for logging positions: " + fp);
    if (!outFile_data) throw std::runtime_error("Could not open file for logging data: " + fd);
}
Logging::~Logging() { close(); }
void Logging::close() {
    if (outFile_position.is_open()) outFile_position.close();
    if (outFile_data.is_open()) outFile_data.close();
}
void Logging::logPositions_xyz(const std::vector<Particle> &particles, const SimulationBox &, double) {
    outFile_position << particles.size() << "\n";  // logging.cpp:30-42: count line, then "1 x y"
    for (const auto &p : particles) outFile_position << "1" << " " << p.x << " " << p.y << "\n";
    outFile_position.flush();
}
void Logging::logPositions_dump(const std::vector<Particle> &particles, const SimulationBox &box, int timestep) {
    if (!outFile_position.is_open()) throw std::runtime_error("Position log file is not open");
    std::ofstream &o = outFile_position;  // LAMMPS dump, logging.cpp:44-79
    o << "ITEM: TIMESTEP\n" << timestep << "\nITEM: NUMBER OF ATOMS\n" << particles.size() << "\n";
    o << "ITEM: BOX BOUNDS pp pp ff\n" << 0.0 << " " << box.getLx() << "\n" << 0.0 << " " << box.getLy() << "\n"
      << 0.0 << " " << 0.0 << "\n";
    o << "ITEM: ATOMS id x y z\n";
    int id = 1;
    for (const auto &p : particles) o << id++ << " " << p.x << " " << p.y << " " << 0.0 << "\n";
    o.flush();
}
void Logging::logSimulationData(const std::vector<Particle> &particles, SimulationBox &box,
                                const ThermodynamicCalculator &cal, int timestep) {
    double U, W;  // one GPU pass gives what the reference computes in two (logging.cpp:80-90)
    cal.computeEnergyVirialCellList(particles, box, U, W);
    const double V = box.getV();
    const double P = cal.computeDensity(particles, box) * cal.getTemperature() + W / (2.0 * V);
    outFile_data << "Timestep:" << timestep << " " << std::fixed << std::setprecision(5) << ",\tEnergy:" << U
                 << ",\tTempreture:" << cal.getTemperature() << ",\tPressure:" << P << ",\tVolume:" << V
                 << ",\tdensity of particles:" << cal.getNumParticles(particles) / V
                 << ",\tNumberofParticles:" << cal.getNumParticles(particles) << "\n";
    outFile_data.flush();
}

// =================================================================================================
// MonteCarlo (serial_src/MC.cpp) on the GPU
// =================================================================================================
// NPT: cells 8 % wider than rcut, so the cell grid of a context serves volumes down to -14 % before the shim
// has to re-create it (and it re-creates it when cells have grown 25 % beyond rcut)
static constexpr double kNptCellMargin = 0.08;

MonteCarlo::MonteCarlo(SimulationBox &box, std::vector<Particle> &particles, ThermodynamicCalculator &calc,
                       double rcut, double delta, double delta_V, Ensemble ensemble, Logging &logger, RNG &rng)
    : box_(box), particles_(particles), calc_(calc), logger_(logger), ensemble_(ensemble), rng_(rng), rcut_(rcut),
      delta_(delta), delta_V_(delta_V) {
    if (ensemble == Ensemble::Gibbs)
        throw std::runtime_error("the Gibbs ensemble is not on the GPU hot path (use the reference build)");
    // the Philox key comes from the caller's generator, so a fixed RNG seed fixes the run
    const std::uint64_t hi = static_cast<std::uint32_t>(rng.uniformInt(0, INT_MAX));
    const std::uint64_t lo = static_cast<std::uint32_t>(rng.uniformInt(0, INT_MAX));
    seed_ = (hi << 32) | lo;
    gpu_.reset(new mc2d::Context(box.getLx(), box.getLy(), rcut, calc.potentialSpec(), calc.getTemperature(),
                                 calc.getActivity() > 0 ? calc.getActivity() : 0.0, delta, seed_, 0, 0, 1, 0,
                                 ensemble == Ensemble::NPT ? kNptCellMargin : 0.0));
    gpu_->check(mc2d_upload(gpu_->get(), raw(particles_), nullptr, (long long)particles_.size()));  // + energy, MC.cpp:26
}
MonteCarlo::~MonteCarlo() = default;

mc2d_stats MonteCarlo::stats() const {
    mc2d_stats st{};
    gpu_->check(mc2d_get_stats(gpu_->get(), &st));
    return st;
}
double MonteCarlo::getEnergy() const { return stats().energy; }
double MonteCarlo::acceptanceRatio() const {
    const mc2d_stats s = stats();
    const double att = double(s.attempted[0] + s.attempted[1] + s.attempted[2]);
    return att > 0 ? double(s.accepted[0] + s.accepted[1] + s.accepted[2]) / att : 0.0;
}
void MonteCarlo::sync_host() const {
    if (!host_stale_) return;
    long long n = 0;
    gpu_->check(mc2d_num_particles(gpu_->get(), &n));
    particles_.resize((size_t)n);
    gpu_->check(mc2d_download(gpu_->get(), reinterpret_cast<double *>(particles_.data()), nullptr, n, &n));
    host_stale_ = false;
}
const std::vector<Particle> &MonteCarlo::getParticles() const {
    sync_host();
    return particles_;
}
void MonteCarlo::updateCellList() {}

void MonteCarlo::recordObservables() {
    sync_host();
    logger_.logSimulationData(particles_, box_, calc_, static_cast<int>(simulation_step_time));
    logger_.logPositions_dump(particles_, box_, static_cast<int>(simulation_step_time));
}

int MonteCarlo::run(size_t nSteps, size_t fOutputStep, size_t fUpdateCell) {
    if (fOutputStep == 0 || fUpdateCell == 0) throw std::invalid_argument("output / cell-update period must be > 0");
    const mc2d_sweep_plan plan{1, ensemble_ == Ensemble::GCMC ? gc_per_cell_ : 0, 1, 0};
    const mc2d_stats before = stats();
    mc2d_stats st = before;
    size_t step = 0;
    while (step < nSteps) {
        // the next sweep index at which the reference does something on the host (MC.cpp:62-73):
        // re-sync when step % fUpdateCell == 0, output when (step+1) % fOutputStep == 0 or at the end
        size_t stop = nSteps;
        if (ensemble_ == Ensemble::NPT) stop = step + 1;  // a volume move may follow every sweep (MC.cpp:53-60)
        const size_t next_out = ((step / fOutputStep) + 1) * fOutputStep;       // first s+1 multiple
        const size_t next_upd = (step % fUpdateCell == 0) ? step + 1 : ((step / fUpdateCell) + 1) * fUpdateCell + 1;
        stop = std::min(stop, std::min(next_out, next_upd));
        const long long attempted0 = st.attempted[0] + st.attempted[1] + st.attempted[2];
        gpu_->check(mc2d_run_sweeps(gpu_->get(), (long long)(stop - step), &plan, &st));
        simulation_step_time += st.attempted[0] + st.attempted[1] + st.attempted[2] - attempted0;
        host_stale_ = true;
        const size_t last = stop - 1;  // index of the sweep just finished
        if (ensemble_ == Ensemble::NPT) {
            if (rng_.uniform01() > 0.3) (void)nptMove();
            simulation_step_time += 1;
            gpu_->check(mc2d_get_stats(gpu_->get(), &st));
        }
        if (last % fUpdateCell == 0) {
            double U, W;
            long long N;
            gpu_->check(mc2d_total_energy_virial(gpu_->get(), 1, 0, &U, &W, &N));
            if (std::abs(st.energy - U) > 1e-6 * std::max(1.0, std::abs(U)))
                std::cout << "running energy drifted from the recomputed one by " << st.energy - U << std::endl;
        }
        if ((last + 1) % fOutputStep == 0 || last == nSteps - 1) recordObservables();
        step = stop;
    }
    const long long acc = st.accepted[0] + st.accepted[1] + st.accepted[2] -
                          (before.accepted[0] + before.accepted[1] + before.accepted[2]);
    const long long att = st.attempted[0] + st.attempted[1] + st.attempted[2] -
                          (before.attempted[0] + before.attempted[1] + before.attempted[2]);
    return att > 0 ? static_cast<int>(acc / att) : 0;  // the reference's integer ratio (MC.cpp:76)
}

// A new context for box nb holding the configuration rescaled to it (the slow way of a volume move: download,
// scale on the host, re-bin on a fresh cell grid).  Used when the box has left the range the present cell
// grid can serve (cells narrower than rcut, or so wide that the sweeps waste work).
void MonteCarlo::regrid(const SimulationBox &nb, std::unique_ptr<mc2d::Context> &out, std::vector<Particle> &scaled) {
    sync_host();
    scaled = particles_;
    const double rx = nb.getLx() / box_.getLx(), ry = nb.getLy() / box_.getLy();
    for (Particle &p : scaled) {  // SimulationBox::recenter, initial.cpp:27-30
        p.x *= rx;
        p.y *= ry;
        if (p.x >= nb.getLx()) p.x = std::nextafter(nb.getLx(), 0.0);
        if (p.y >= nb.getLy()) p.y = std::nextafter(nb.getLy(), 0.0);
    }
    ++regrids_;  // a fresh Philox key: the new context starts its sweep counter at zero
    out.reset(new mc2d::Context(nb.getLx(), nb.getLy(), rcut_, calc_.potentialSpec(), calc_.getTemperature(),
                                calc_.getActivity() > 0 ? calc_.getActivity() : 0.0, delta_,
                                seed_ + 0x9E3779B97F4A7C15ull * (std::uint64_t)regrids_, 0, 0, 1, 0, kNptCellMargin));
    out->check(mc2d_upload(out->get(), raw(scaled), nullptr, (long long)scaled.size()));
}

bool MonteCarlo::nptMove() {
    const double beta = 1.0 / calc_.getTemperature(), press = calc_.getPressure();
    const double V_o = box_.getV();
    mc2d_stats st{};
    gpu_->check(mc2d_get_stats(gpu_->get(), &st));
    const double U_old = st.energy;
    const double N = (double)st.n_particles;
    const double lnvn = std::log(V_o) + rng_.uniform(-0.5, 0.5) * delta_V_;
    const double V_n = std::exp(lnvn);
    SimulationBox nb = box_;
    nb.setV(V_n);
    double U_new = 0, W_new = 0;
    std::unique_ptr<mc2d::Context> fresh;
    std::vector<Particle> scaled;
    const int rc = mc2d_rescale_try(gpu_->get(), nb.getLx(), nb.getLy(), &U_new, &W_new);
    if (rc == MC2D_ERR_GEOMETRY) {  // the present cell grid cannot serve the smaller box
        regrid(nb, fresh, scaled);
        mc2d_stats s2{};
        fresh->check(mc2d_get_stats(fresh->get(), &s2));
        U_new = s2.energy;
    } else {
        gpu_->check(rc);
    }
    const double arg = -beta * ((U_new - U_old) + press * (V_n - V_o) - (N + 1) * std::log(V_n / V_o) / beta);
    const bool accept = rng_.uniform01() < std::exp(arg);  // MC.cpp:171
    if (fresh) {
        if (accept) {
            gpu_ = std::move(fresh);
            particles_ = scaled;
            host_stale_ = false;
        }
    } else {
        gpu_->check(mc2d_rescale_finish(gpu_->get(), accept ? 1 : 0));
        if (accept) host_stale_ = true;
    }
    if (accept) {
        box_ = nb;
        // cells much wider than rcut waste pair tests: re-create the grid (no acceptance step, same configuration)
        mc2d_info info{};
        gpu_->check(mc2d_get_info(gpu_->get(), &info));
        if (!fresh && std::min(info.cell_dx, info.cell_dy) > 1.25 * rcut_ && std::min(box_.getLx(), box_.getLy()) > 8 * rcut_) {
            std::unique_ptr<mc2d::Context> finer;
            regrid(box_, finer, scaled);
            gpu_ = std::move(finer);
            particles_ = scaled;
            host_stale_ = false;
        }
    }
    return accept;
}

bool MonteCarlo::stepCellList() {
    const mc2d_sweep_plan plan{1, 0, 1, 0};
    mc2d_stats a = stats(), b{};
    gpu_->check(mc2d_run_sweeps(gpu_->get(), 1, &plan, &b));
    host_stale_ = true;
    return b.accepted[0] > a.accepted[0];
}
bool MonteCarlo::stepBruteForce() { return stepCellList(); }

// =================================================================================================
// GibbsMonteCarlo — serial_src/GibbsMC.cpp on two GPU contexts
// =================================================================================================
struct GibbsMonteCarlo::Side {
    SimulationBox box;
    std::vector<Particle> particles;
    ThermodynamicCalculator &calc;
    std::unique_ptr<mc2d::Context> gpu;
    bool host_stale = false;
    std::uint64_t seed;
    int regrids = 0;
    double rcut, delta;

    Side(const SimulationBox &b, std::vector<Particle> p, ThermodynamicCalculator &c, double rcut_, double delta_,
         std::uint64_t seed_)
        : box(b), particles(std::move(p)), calc(c), seed(seed_), rcut(rcut_), delta(delta_) {
        make_context(box, particles, gpu);
    }
    void make_context(const SimulationBox &b, const std::vector<Particle> &p, std::unique_ptr<mc2d::Context> &out) {
        out.reset(new mc2d::Context(b.getLx(), b.getLy(), rcut, calc.potentialSpec(), calc.getTemperature(), 0.0, delta,
                                    seed + 0x9E3779B97F4A7C15ull * (std::uint64_t)regrids, 0, 0, 1, 0, kNptCellMargin));
        out->check(mc2d_upload(out->get(), raw(p), nullptr, (long long)p.size()));
        ++regrids;
    }
    mc2d_stats stats() const {
        mc2d_stats st{};
        gpu->check(mc2d_get_stats(gpu->get(), &st));
        return st;
    }
    void sync_host() {
        if (!host_stale) return;
        long long n = 0;
        gpu->check(mc2d_num_particles(gpu->get(), &n));
        particles.resize((size_t)n);
        gpu->check(mc2d_download(gpu->get(), reinterpret_cast<double *>(particles.data()), nullptr, n, &n));
        host_stale = false;
    }
    // U of the configuration rescaled to nb; `fresh` is set when the present cell grid cannot serve nb
    double try_box(const SimulationBox &nb, std::unique_ptr<mc2d::Context> &fresh, std::vector<Particle> &scaled) {
        double U = 0, W = 0;
        const int rc = mc2d_rescale_try(gpu->get(), nb.getLx(), nb.getLy(), &U, &W);
        if (rc == MC2D_ERR_GEOMETRY) {
            sync_host();
            scaled = particles;
            const double rx = nb.getLx() / box.getLx(), ry = nb.getLy() / box.getLy();
            for (Particle &p : scaled) {
                p.x *= rx;
                p.y *= ry;
                if (p.x >= nb.getLx()) p.x = std::nextafter(nb.getLx(), 0.0);
                if (p.y >= nb.getLy()) p.y = std::nextafter(nb.getLy(), 0.0);
            }
            make_context(nb, scaled, fresh);
            mc2d_stats st{};
            fresh->check(mc2d_get_stats(fresh->get(), &st));
            return st.energy;
        }
        gpu->check(rc);
        return U;
    }
    void finish_box(bool accept, const SimulationBox &nb, std::unique_ptr<mc2d::Context> &fresh, std::vector<Particle> &scaled) {
        if (fresh) {
            if (accept) {
                gpu = std::move(fresh);
                particles = scaled;
                host_stale = false;
            }
        } else {
            gpu->check(mc2d_rescale_finish(gpu->get(), accept ? 1 : 0));
            if (accept) host_stale = true;
        }
        if (accept) box = nb;
    }
};

GibbsMonteCarlo::GibbsMonteCarlo(SimulationBox &box1, SimulationBox &box2, std::vector<Particle> particles1,
                                 std::vector<Particle> particles2, ThermodynamicCalculator &calc1,
                                 ThermodynamicCalculator &calc2, double rcut, double delta, double delta_V,
                                 Logging &logger1, Logging &logger2, RNG &rng)
    : rng_(rng), rcut_(rcut), delta_(delta), delta_V_(delta_V), beta_(1.0 / calc1.getTemperature()) {
    const std::uint64_t hi = static_cast<std::uint32_t>(rng.uniformInt(0, INT_MAX));
    const std::uint64_t lo = static_cast<std::uint32_t>(rng.uniformInt(0, INT_MAX));
    const std::uint64_t seed = (hi << 32) | lo;
    s_[0].reset(new Side(box1, std::move(particles1), calc1, rcut, delta, seed));
    s_[1].reset(new Side(box2, std::move(particles2), calc2, rcut, delta, seed ^ 0xD1B54A32D192ED03ull));
    logger_[0] = &logger1;
    logger_[1] = &logger2;
}
GibbsMonteCarlo::~GibbsMonteCarlo() = default;

const SimulationBox &GibbsMonteCarlo::getBox(int which) const { return s_[which]->box; }
long long GibbsMonteCarlo::getNumParticles(int which) const { return s_[which]->stats().n_particles; }
double GibbsMonteCarlo::getEnergy(int which) const { return s_[which]->stats().energy; }
const std::vector<Particle> &GibbsMonteCarlo::getParticles(int which) {
    s_[which]->sync_host();
    return s_[which]->particles;
}

void GibbsMonteCarlo::recordObservables() {  // GibbsMC.cpp:467-472
    for (int w = 0; w < 2; ++w) {
        s_[w]->sync_host();
        logger_[w]->logSimulationData(s_[w]->particles, s_[w]->box, s_[w]->calc, static_cast<int>(simulation_step_time));
        logger_[w]->logPositions_dump(s_[w]->particles, s_[w]->box, static_cast<int>(simulation_step_time));
    }
}

// one displacement trial per particle in both boxes (particle_displacement_1/2 x N, GibbsMC.cpp:214-231)
void GibbsMonteCarlo::sweep_both(long long &acc, long long &tot) {
    const mc2d_sweep_plan plan{1, 0, 1, 0};
    for (int w = 0; w < 2; ++w) {
        const mc2d_stats a = s_[w]->stats();
        if (a.n_particles == 0) continue;
        mc2d_stats b{};
        s_[w]->gpu->check(mc2d_run_sweeps(s_[w]->gpu->get(), 1, &plan, &b));
        s_[w]->host_stale = true;
        acc += b.accepted[0] - a.accepted[0];
        tot += b.attempted[0] - a.attempted[0];
        simulation_step_time += b.attempted[0] - a.attempted[0];
    }
}

// GibbsMC.cpp:340-394: a random point in the receiving box, a random particle of the donor box
bool GibbsMonteCarlo::particle_exchange() {
    const int to = rng_.uniform01() > 0.5 ? 0 : 1, from = 1 - to;  // > 0.5: remove from box two, insert into box one
    Side &T = *s_[to], &F = *s_[from];
    const double N_from = (double)F.stats().n_particles;
    if (N_from < 50) return false;
    const double N_to = (double)T.stats().n_particles;
    const double x = rng_.uniform(0, T.box.getLx()), y = rng_.uniform(0, T.box.getLy());
    double dU_to = 0, dU_from = 0, px = 0, py = 0;
    T.gpu->check(mc2d_probe_point(T.gpu->get(), std::min(x, std::nextafter(T.box.getLx(), 0.0)),
                                  std::min(y, std::nextafter(T.box.getLy(), 0.0)), &dU_to));
    (to == 0 ? w_1 : w_2) += T.box.getV() * std::exp(-beta_ * dU_to) / (N_to + 1);
    const long long idx = rng_.uniformInt(0, (int)N_from - 1);
    F.gpu->check(mc2d_probe_particle(F.gpu->get(), idx, &px, &py, &dU_from));
    // Acceptance of moving one particle from box F to box T (Frenkel & Smit, Eq. 8.3.3):
    //   N_F V_T / ((N_T + 1) V_F) exp(-beta dU).
    // GibbsMC.cpp:359,382 has the ratio upside down, V_F (N_T + 1) / (V_T (N_F + 1)): with it an ideal gas
    // collapses into one box.  A reference defect, not copied.
    // reference_exchange_ratio = true reproduces the reference's expression for output-parity runs.
    const double ratio = reference_exchange_ratio ? F.box.getV() * (N_to + 1) / (T.box.getV() * (N_from + 1))
                                                  : N_from * T.box.getV() / ((N_to + 1) * F.box.getV());
    const double arg = std::exp(-beta_ * (dU_to + dU_from)) * ratio;
    const bool accept = rng_.uniform01() < arg;
    if (accept) {
        T.gpu->check(mc2d_apply_probe(T.gpu->get(), 1));
        F.gpu->check(mc2d_apply_probe(F.gpu->get(), 2));
        T.host_stale = F.host_stale = true;
    } else {
        T.gpu->check(mc2d_apply_probe(T.gpu->get(), 0));
        F.gpu->check(mc2d_apply_probe(F.gpu->get(), 0));
    }
    return accept;
}

// GibbsMC.cpp:396-460: the total volume is kept, ln(V1 / V2) takes a uniform step of width delta_V
bool GibbsMonteCarlo::volume_change() {
    Side &A = *s_[0], &B = *s_[1];
    const double V1 = A.box.getV(), V2 = B.box.getV();
    const mc2d_stats sa = A.stats(), sb = B.stats();
    const double U_old = sa.energy + sb.energy;
    const double dlnV = rng_.uniform(-0.5, 0.5) * delta_V_;
    const double q = std::exp(std::log(V1 / V2) + dlnV);
    const double V1n = (V1 + V2) * q / (1 + q), V2n = (V1 + V2) - V1n;
    SimulationBox b1 = A.box, b2 = B.box;
    b1.setV(V1n);
    b2.setV(V2n);
    std::unique_ptr<mc2d::Context> f1, f2;
    std::vector<Particle> sc1, sc2;
    const double U_new = A.try_box(b1, f1, sc1) + B.try_box(b2, f2, sc2);
    double arg = std::exp(-beta_ * (U_new - U_old));
    arg *= std::pow(V1n / V1, (double)sa.n_particles + 1) * std::pow(V2n / V2, (double)sb.n_particles + 1);
    const bool accept = rng_.uniform01() < arg;
    A.finish_box(accept, b1, f1, sc1);
    B.finish_box(accept, b2, f2, sc2);
    return accept;
}

int GibbsMonteCarlo::loop(Loop kind, size_t nSteps, size_t fOutputStep, size_t fUpdateCell) {
    if (fOutputStep == 0 || fUpdateCell == 0) throw std::invalid_argument("output / cell-update period must be > 0");
    long long acc = 0, tot = 0;
    for (size_t step = 0; step < nSteps; ++step) {
        const double r = rng_.uniform01();  // drawn in every loop of the reference, used by two of them
        bool single = false, ok = false;
        switch (kind) {
        case Loop::NVT: sweep_both(acc, tot); break;
        case Loop::NPT:
            sweep_both(acc, tot);
            ok = volume_change();
            single = true;
            break;
        case Loop::muVT:
            if (r < 0.5) sweep_both(acc, tot);
            else { ok = particle_exchange(); single = true; }
            break;
        case Loop::Gibbs:
            if (r < 0.5) sweep_both(acc, tot);
            else if (r < 0.75) { ok = particle_exchange(); single = true; }
            else { ok = volume_change(); single = true; }
            break;
        }
        if (single) {
            acc += ok ? 1 : 0;
            tot += 1;
            simulation_step_time += 1;
        }
        if (step % fUpdateCell == 0) {  // energy re-sync of both boxes (GibbsMC.cpp:249-263)
            for (int w = 0; w < 2; ++w) {
                double U, W;
                long long N;
                s_[w]->gpu->check(mc2d_total_energy_virial(s_[w]->gpu->get(), 1, 0, &U, &W, &N));
            }
        }
        if ((step + 1) % fOutputStep == 0 || step == nSteps - 1) recordObservables();
    }
    return tot > 0 ? static_cast<int>(acc / tot) : 0;  // the reference's integer ratio
}
int GibbsMonteCarlo::run(size_t n, size_t fo, size_t fu) { return loop(Loop::Gibbs, n, fo, fu); }
int GibbsMonteCarlo::equlibrateNVT(size_t n, size_t fo, size_t fu) { return loop(Loop::NVT, n, fo, fu); }
int GibbsMonteCarlo::equlibrateNPT(size_t n, size_t fo, size_t fu) { return loop(Loop::NPT, n, fo, fu); }
int GibbsMonteCarlo::equlibratemuVT(size_t n, size_t fo, size_t fu) { return loop(Loop::muVT, n, fo, fu); }

// =================================================================================================
// MonteCarloNVT_Slabs — MonteCarloNVT_MPI (mpi_src/MC_parallel.cpp) on one GPU per rank
// =================================================================================================
MonteCarloNVT_Slabs::MonteCarloNVT_Slabs(int rank, int nranks, int device, SimulationBox &box,
                                         ThermodynamicCalculator &thermo, std::vector<Particle> &owned,
                                         std::uint64_t seed, const Params &p, const BcastFn &bcast)
    : rank_(rank), nranks_(nranks), box_(box), thermo_(thermo), owned_(owned), p_(p) {
    gpu_.reset(new mc2d::Context(box.getLx(), box.getLy(), p.rcut, thermo.potentialSpec(), thermo.getTemperature(),
                                 thermo.getActivity() > 0 ? thermo.getActivity() : 0.0, p.delta, seed, device, rank,
                                 nranks));
    if (nranks > 1) {
        unsigned char id[MC2D_UNIQUE_ID_BYTES] = {0};
        if (rank == 0 && mc2d_comm_unique_id(id) != MC2D_OK) throw std::runtime_error("cannot create an NCCL id");
        bcast(id, sizeof id, 0);  // MPI_Bcast in the reference's world
        gpu_->check(mc2d_comm_init(gpu_->get(), id));
    }
    gpu_->check(mc2d_upload(gpu_->get(), raw(owned_), nullptr, (long long)owned_.size()));
}
MonteCarloNVT_Slabs::~MonteCarloNVT_Slabs() = default;

double MonteCarloNVT_Slabs::run(std::size_t nsweeps, int) {
    const mc2d_sweep_plan plan{1, p_.gc_per_cell, 1, 0};
    mc2d_stats st{};
    gpu_->check(mc2d_run_sweeps(gpu_->get(), (long long)nsweeps, &plan, &st));
    gpu_->check(mc2d_allreduce_stats(gpu_->get(), &st));  // MC_parallel.cpp:111-112
    const double att = double(st.attempted[0] + st.attempted[1] + st.attempted[2]);
    return att > 0 ? double(st.accepted[0] + st.accepted[1] + st.accepted[2]) / att : 0.0;
}
double MonteCarloNVT_Slabs::totalEnergy() {
    double U, W;
    long long N;
    gpu_->check(mc2d_total_energy_virial(gpu_->get(), 0, 1, &U, &W, &N));
    return U;
}
double MonteCarloNVT_Slabs::totalVirial() {
    double U, W;
    long long N;
    gpu_->check(mc2d_total_energy_virial(gpu_->get(), 0, 1, &U, &W, &N));
    return W;
}
long long MonteCarloNVT_Slabs::numParticlesGlobal() {
    double U, W;
    long long N;
    gpu_->check(mc2d_total_energy_virial(gpu_->get(), 0, 1, &U, &W, &N));
    return N;
}
double MonteCarloNVT_Slabs::pressure() {  // thermodynamic_calculator_parallel.cpp:136-145
    double U, W;
    long long N;
    gpu_->check(mc2d_total_energy_virial(gpu_->get(), 0, 1, &U, &W, &N));
    const double V = box_.getV();
    return (double(N) / V) * thermo_.getTemperature() + W / (2.0 * V);
}
void MonteCarloNVT_Slabs::download(std::vector<Particle> &owned) {
    long long n = 0;
    gpu_->check(mc2d_num_particles(gpu_->get(), &n));
    owned.resize((size_t)n);
    gpu_->check(mc2d_download(gpu_->get(), reinterpret_cast<double *>(owned.data()), nullptr, n, &n));
}
