// traj_ref_main.cpp — TEST INFRASTRUCTURE: drives the reference's UNMODIFIED LoggingTrajMPI
// (/root/reference/mpi_src/logging_traj_mpi.cpp) on the threads-as-ranks MPI shim and writes the frames of
// host/traj_cases.hpp in all three file modes.  The launcher of mpi_shim.cpp calls this as every rank's main();
// its "input file" argument is used as the output directory:
//     oracle/_ref/traj_ref <nranks> <outdir>
// writes <outdir>/{perrank,gather,mpiio}{,_typed}.{xyz,dump} (+ .rankNNNN files).  tests/test_traj_logger.py
// compares them byte for byte with what host/slab_traj_selftest writes through LoggingTrajSlabs, and with the
// copies committed under tests/golden/traj/.
#include <mpi.h>

#include <string>
#include <vector>

#include "logging_traj_mpi.h"
#include "particle.h"
#include "simulation_box.h"
#include "traj_cases.hpp"

int mc2d_ref_mpi_main(int argc, char *argv[]) {
    if (argc < 2) return 1;
    MPI_Init(&argc, &argv);
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const std::string dir = argv[1];
    SimulationBox box(traj_cases::kLx, traj_cases::kLy);
    const LoggingTrajMPI::Mode modes[3] = {LoggingTrajMPI::Mode::PerRank, LoggingTrajMPI::Mode::GatherRoot, LoggingTrajMPI::Mode::MPIIO};
    const char *names[3] = {"perrank", "gather", "mpiio"};
    for (int m = 0; m < 3; ++m)
        for (int typed = 0; typed < 2; ++typed) {
            LoggingTrajMPI traj(dir + "/" + names[m] + (typed ? "_typed" : ""), modes[m], MPI_COMM_WORLD, /*append=*/false, typed != 0);
            for (int f = 0; f < traj_cases::kFrames; ++f) {
                const std::vector<double> xy = traj_cases::frame_xy(rank, f);
                std::vector<Particle> local;
                for (std::size_t i = 0; i + 1 < xy.size(); i += 2) local.emplace_back(xy[i], xy[i + 1]);
                traj.log_xyz(local, box);
                traj.log_dump(local, box, traj_cases::timestep(f));
            }
            traj.close();
        }
    MPI_Finalize();
    return 0;
}
