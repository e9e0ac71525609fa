// slab_traj_selftest — writes the frames of traj_cases.hpp through LoggingTrajSlabs in every file mode, with
// forked ranks exactly as mc2d_slabs starts them (no GPU involved): `slab_traj_selftest <nranks> <outdir>`.
// Files: <outdir>/{perrank,gather,mpiio}{,_typed}.{xyz,dump} (+ .rankNNNN for the per-rank mode) and, for the
// append option, <outdir>/append_{gather,mpiio}.dump (two loggers in a row, the second appending).  Also the serial
// Logging class's position files, <outdir>/serial.{xyz,dump} (frames of rank 0, 2, 0).
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdlib>
#include <iostream>
#include <new>
#include <string>
#include <vector>

#include "slab_traj_logger.hpp"
#include "traj_cases.hpp"

using mc2d_slabs::LoggingTrajSlabs;
using mc2d_slabs::Ranks;
using mc2d_slabs::Shared;

static std::vector<Particle> particles(int rank, int frame) {
    const std::vector<double> xy = traj_cases::frame_xy(rank, frame);
    std::vector<Particle> p;
    for (std::size_t i = 0; i + 1 < xy.size(); i += 2) p.emplace_back(xy[i], xy[i + 1]);
    return p;
}

static int run_rank(const Ranks &R, const std::string &dir) {
    SimulationBox box(traj_cases::kLx, traj_cases::kLy);
    const LoggingTrajSlabs::Mode modes[3] = {LoggingTrajSlabs::Mode::PerRank, LoggingTrajSlabs::Mode::GatherRoot,
                                             LoggingTrajSlabs::Mode::MPIIO};
    const char *names[3] = {"perrank", "gather", "mpiio"};
    for (int m = 0; m < 3; ++m)
        for (int typed = 0; typed < 2; ++typed) {
            LoggingTrajSlabs traj(dir + "/" + names[m] + (typed ? "_typed" : ""), modes[m], R, false, typed != 0);
            for (int f = 0; f < traj_cases::kFrames; ++f) {
                const std::vector<Particle> local = particles(R.rank, f);
                traj.log_xyz(local, box);
                traj.log_dump(local, box, traj_cases::timestep(f));
            }
            traj.close();
        }
    // append: frames 0..1 by one logger, frame 2 by a second one opened with append = true
    for (int m = 1; m < 3; ++m) {
        const std::string base = dir + "/append_" + names[m];
        {
            LoggingTrajSlabs first(base, modes[m], R, false, true);
            for (int f = 0; f < 2; ++f) first.log_dump(particles(R.rank, f), box, traj_cases::timestep(f));
        }
        R.barrier();
        LoggingTrajSlabs second(base, modes[m], R, true, true);
        second.log_dump(particles(R.rank, 2), box, traj_cases::timestep(2));
    }
    R.barrier();
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::cerr << "usage: " << argv[0] << " <nranks> <outdir>\n";
        return 1;
    }
    const int size = std::atoi(argv[1]);
    if (size < 1 || size > mc2d_slabs::kMaxRanks) return 1;
    {   // the serial logger's position files (Logging, logging.cpp:30-79): frames of "ranks" 0 and 2
        const std::string dir = argv[2];
        SimulationBox box(traj_cases::kLx, traj_cases::kLy);
        Logging xyz(dir + "/serial.xyz", dir + "/serial_xyz.dat"), dump(dir + "/serial.dump", dir + "/serial_dump.dat");
        for (int f = 0; f < traj_cases::kFrames; ++f) {
            const std::vector<Particle> p = particles(f == 1 ? 2 : 0, f);
            xyz.logPositions_xyz(p, box, 6.25);
            dump.logPositions_dump(p, box, traj_cases::timestep(f));
        }
    }
    void *mem = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) return 2;
    Shared *sh = new (mem) Shared();
    sh->arrived = 0;
    sh->generation = 0;
    sh->turn = 0;
    sh->failed = 0;
    sh->session = (long long)getpid();
    std::vector<pid_t> kids;
    for (int r = 0; r < size; ++r) {
        const pid_t c = fork();
        if (c < 0) return 2;
        if (c == 0) {
            int rc = 2;
            try {
                rc = run_rank(Ranks{sh, r, size}, argv[2]);
            } catch (const std::exception &e) {
                sh->failed = 1;
                std::cerr << "rank " << r << ": " << e.what() << "\n";
            }
            _exit(rc);
        }
        kids.push_back(c);
    }
    int worst = 0;
    for (pid_t c : kids) {
        int st = 0;
        waitpid(c, &st, 0);
        worst = std::max(worst, WIFEXITED(st) ? WEXITSTATUS(st) : 2);
    }
    return worst;
}
