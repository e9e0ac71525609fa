// mc2d_slabs — the mpi_src/main.cpp-equivalent driver on GPU slabs: `mc2d_slabs input.txt [nranks]`.
//
// The reference's MPI driver (mpi_src/main.cpp:28-196) starts one rank per tile under mpirun.  This
// image has no MPI, and the path does not need one: the only rank-to-rank traffic outside the GPU
// library is the 128-byte communicator id, a barrier and turn-taking for the log files.  The driver
// therefore forks one process per GPU itself and gives them a small shared-memory page for those
// three things; everything else (halo, migration, reductions) happens inside libmc2d over NVLink/NCCL.
//
// Same input keys as mpi_src/main.cpp:52-78 (Lx Ly N rcut T nSteps eSteps outputFreq cellUpdateFreq
// potential out_xyz out_data delta; optional mu f alpha A_0 kappa seed ensemble), same outputs:
//   <basename>.thermo.csv  header "timestep,N_global,rho,T,U,W,P,V", 17 significant digits
//                          (logging_data_mpi.cpp:56-67)
//   <basename>.dump        LAMMPS dump frames "id type x y z", type = rank + 1
//                          (logging_traj_mpi.cpp:268-303, Mode::MPIIO with rank_as_type — what mpi_src/main.cpp:164
//                          asks for; written by LoggingTrajSlabs, slab_traj_logger.hpp).  The optional key
//                          `traj_mode perrank | gather | shared` selects the class's other file modes.
// Differences: ensemble GCMC is accepted (mu is the activity, as in serial_src; the reference's MPI
// driver is NVT only); `init lattice` starts from a jittered square lattice instead of uniformly
// random points (which overlap and give LJ energies of 1e30); rcut is never shrunk silently.
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "mc2d.h"
#include "mc2d_host.hpp"
#include "slab_ranks.hpp"
#include "slab_traj_logger.hpp"

namespace {

using mc2d_slabs::LoggingTrajSlabs;
using mc2d_slabs::Ranks;
using mc2d_slabs::Shared;

double get_or(Input &in, const std::string &key, double dflt) {
    try {
        return in.getConstant(key);
    } catch (...) {
        return dflt;
    }
}

std::string strip_ext(const std::string &s) {
    const auto pos = s.rfind('.');
    return pos == std::string::npos ? s : s.substr(0, pos);
}

// the global start configuration, identical on every rank (mc2d_upload keeps the rank's slab)
std::vector<Particle> initial_configuration(long long N, double Lx, double Ly, bool lattice, unsigned seed) {
    std::vector<Particle> p;
    p.reserve((std::size_t)N);
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    if (!lattice) {
        for (long long i = 0; i < N; ++i) p.emplace_back(u(gen) * Lx, u(gen) * Ly);
        return p;
    }
    const long long nx = (long long)std::ceil(std::sqrt((double)N * Lx / Ly));
    const long long ny = (N + nx - 1) / nx;
    const double ax = Lx / (double)nx, ay = Ly / (double)ny, j = 0.2 * std::min(ax, ay);
    for (long long i = 0; i < N; ++i) {
        const long long ix = i % nx, iy = i / nx;
        double x = (ix + 0.5) * ax + (2.0 * u(gen) - 1.0) * j, y = (iy + 0.5) * ay + (2.0 * u(gen) - 1.0) * j;
        p.emplace_back(std::fmin(std::fmax(x, 0.0), std::nextafter(Lx, 0.0)), std::fmin(std::fmax(y, 0.0), std::nextafter(Ly, 0.0)));
    }
    return p;
}

int run_rank(const Ranks &R, const std::string &input_file) {
    Input input(input_file);
    const double Lx = input.getConstant("Lx"), Ly = input.getConstant("Ly");
    const long long Nglob = (long long)input.getConstant("N");
    const double rcut = input.getConstant("rcut"), T = input.getConstant("T");
    const double nSteps = input.getConstant("nSteps"), eSteps = input.getConstant("eSteps");
    const int outputFreq = (int)input.getConstant("outputFreq");
    const std::string potentialName = input.getFilename("potential");
    const std::string out_xyz = input.getFilename("out_xyz"), out_data = input.getFilename("out_data");
    const double delta = input.getConstant("delta");
    const double mu = get_or(input, "mu", 0.0), f = get_or(input, "f", 0.0), alpha = get_or(input, "alpha", 0.0);
    const double A0 = get_or(input, "A_0", 0.0), kappa = get_or(input, "kappa", 0.0);
    const unsigned seed = (unsigned)get_or(input, "seed", 0.0);
    const std::string ensemble = input.getFilename("ensemble"), init = input.getFilename("init");
    if (!ensemble.empty() && ensemble != "NVT" && ensemble != "GCMC")
        throw std::invalid_argument("mc2d_slabs supports NVT and GCMC, got '" + ensemble + "'");
    const bool gcmc = ensemble == "GCMC";
    if (R.rank == 0) {
        std::cout << "\n[Input Summary / GPU slabs]\nRanks = " << R.size << "\nLx = " << Lx << ", Ly = " << Ly
                  << ", N_global = " << Nglob << "\nrcut = " << rcut << ", T = " << T << "\neSteps = " << eSteps
                  << ", nSteps = " << nSteps << "\noutputFreq = " << outputFreq << "\npotential = " << potentialName
                  << "\nmu = " << mu << ", f = " << f << ", alpha = " << alpha << ", A_0 = " << A0 << ", kappa = " << kappa
                  << "\nseed = " << seed << "\nensemble = " << (ensemble.empty() ? "NVT (default)" : ensemble) << "\n\n";
    }
    SimulationBox box(Lx, Ly);
    const PotentialType pot = selectPotentialType(potentialName);
    ThermodynamicCalculator thermo(T, 0.0, pot, rcut, mu, f, alpha, A0, kappa);
    std::vector<Particle> owned = initial_configuration(Nglob, Lx, Ly, init == "lattice", seed);

    MonteCarloNVT_Slabs::Params mp;
    mp.rcut = rcut;
    mp.delta = delta;
    mp.out_every = outputFreq;
    mp.gc_per_cell = gcmc ? 1 : 0;
    MonteCarloNVT_Slabs mc(R.rank, R.size, /*device=*/R.rank, box, thermo, owned, (std::uint64_t)seed, mp,
                           [&](void *buf, std::size_t bytes, int root) { R.bcast(buf, bytes, root); });

    const std::string basename = !out_xyz.empty() ? strip_ext(out_xyz) : (!out_data.empty() ? strip_ext(out_data) : "run");
    std::ofstream csv;
    if (R.rank == 0) {
        csv.open(basename + ".thermo.csv", std::ios::out | std::ios::trunc);
        if (!csv) throw std::runtime_error("cannot open " + basename + ".thermo.csv");
        csv << "timestep,N_global,rho,T,U,W,P,V\n";
    }
    R.barrier();
    const std::string traj_mode = input.getFilename("traj_mode");
    if (!traj_mode.empty() && traj_mode != "perrank" && traj_mode != "gather" && traj_mode != "shared")
        throw std::invalid_argument("traj_mode must be perrank, gather or shared, got '" + traj_mode + "'");
    LoggingTrajSlabs traj(basename,
                          traj_mode == "perrank" ? LoggingTrajSlabs::Mode::PerRank
                                                 : (traj_mode == "gather" ? LoggingTrajSlabs::Mode::GatherRoot : LoggingTrajSlabs::Mode::MPIIO),
                          R, /*append=*/false, /*rank_as_type=*/true);

    int timestep = 0;
    auto log_frame = [&]() {
        // thermo row (logging_data_mpi.cpp:32-67): the reductions are collective
        const double U = mc.totalEnergy(), W = mc.totalVirial(), P = mc.pressure();
        const long long Ng = mc.numParticlesGlobal();
        const double V = box.getV();
        if (R.rank == 0) {
            csv << std::setprecision(17) << timestep << "," << Ng << "," << (V > 0 ? (double)Ng / V : 0.0) << "," << T << ","
                << U << "," << W << "," << P << "," << V << "\n";
            csv.flush();
        }
        // dump frame (MC_parallel.cpp:102)
        mc.download(owned);
        traj.log_dump(owned, box, timestep);
    };
    auto run_phase = [&](std::size_t nsweeps, const char *label) {
        if (nsweeps == 0) return;
        if (R.rank == 0) std::cout << label << "\n";
        std::size_t done = 0;
        double acc = 0.0;
        while (done < nsweeps) {  // MC_parallel.cpp:69-107: log every out_every sweeps
            const std::size_t chunk = outputFreq > 0 ? std::min<std::size_t>(outputFreq - (done % outputFreq), nsweeps - done)
                                                     : nsweeps - done;
            acc = mc.run(chunk, timestep);
            done += chunk;
            timestep += (int)chunk;
            if (outputFreq > 0 && done % outputFreq == 0) log_frame();
        }
        if (R.rank == 0) std::cout << "timestep:\t" << timestep << "  acceptance " << acc << "\n";
    };
    run_phase((std::size_t)eSteps, "Equilibration (GPU slabs)...");
    run_phase((std::size_t)nSteps, "Production (GPU slabs)...");
    if (R.rank == 0) {
        csv.close();
        std::cout << "Simulation completed.\n";
    }
    R.barrier();
    return 0;
}

}  // namespace

int main(int argc, char *argv[]) {
    if (argc < 2) {
        std::cerr << "Usage: " << argv[0] << " <input_file> [nranks = number of GPUs]\n";
        return 1;
    }
    int size = argc > 2 ? std::atoi(argv[2]) : 0;
    if (size <= 0) {
        // count devices in a child so that the parent never creates a CUDA context before forking
        int fds[2];
        if (pipe(fds) != 0) return 2;
        const pid_t c = fork();
        if (c == 0) {
            const int n = mc2d_device_count();
            if (write(fds[1], &n, sizeof n) != (ssize_t)sizeof n) _exit(3);
            _exit(0);
        }
        int n = 0;
        if (read(fds[0], &n, sizeof n) != (ssize_t)sizeof n) n = 0;
        waitpid(c, nullptr, 0);
        size = n;
    }
    if (size <= 0) {
        std::cerr << "Error: no CUDA device visible; mc2d has no CPU path\n";
        return 2;
    }
    if (size > 64) size = 64;
    void *mem = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) {
        std::perror("mmap");
        return 2;
    }
    Shared *sh = new (mem) Shared();
    sh->arrived = 0;
    sh->generation = 0;
    sh->turn = 0;
    sh->failed = 0;
    sh->session = (long long)getpid();
    std::vector<pid_t> kids;
    for (int r = 0; r < size; ++r) {
        const pid_t c = fork();
        if (c < 0) {
            std::perror("fork");
            return 2;
        }
        if (c == 0) {
            int rc = 2;
            try {
                rc = run_rank(Ranks{sh, r, size}, argv[1]);
            } catch (const std::exception &e) {
                sh->failed = 1;
                std::cerr << "Error (rank " << r << "): " << e.what() << "\n";
            }
            std::cout.flush();
            _exit(rc);
        }
        kids.push_back(c);
    }
    int worst = 0;
    for (pid_t c : kids) {
        int st = 0;
        waitpid(c, &st, 0);
        const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 2;
        worst = std::max(worst, rc);
    }
    munmap(mem, sizeof(Shared));
    return worst;
}
