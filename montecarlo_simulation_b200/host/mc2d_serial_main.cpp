// mc2d_serial — the serial_src/main.cpp driver on the GPU hot path.
//
//   mc2d_serial <input_file> [key=value ...]
//
// Reads the reference's input.txt format (serial_src/main.cpp:39-67: Lx Ly N rcut T nSteps eSteps
// outputFreq cellUpdateFreq delta mu f alpha A_0 kappa seed P delta_V; potential out_xyz out_data
// position_file ensemble), with the reference's semantics: a missing numeric key is 0, `mu` is the
// ACTIVITY z, the default ensemble is GCMC.  Runs eSteps equilibration sweeps then nSteps
// production sweeps and writes the same dump / data logs.  Extra optional key: gc_per_cell
// (grand-canonical draws per cell per sweep, default 1).
#include <iostream>

#include "mc2d_host.hpp"

int main(int argc, char *argv[]) {
    if (argc < 2) {
        std::cerr << "Usage: " << argv[0] << " <input_file> [key=value ...]" << std::endl;
        return 1;
    }
    try {
        Input input(argv[1]);
        input.parseCommandLineArgs(argc - 1, argv + 1);
        const double Lx = input.getConstant("Lx"), Ly = input.getConstant("Ly");
        const int N = static_cast<int>(input.getConstant("N"));
        const double rcut = input.getConstant("rcut"), T = input.getConstant("T");
        const double nSteps = input.getConstant("nSteps"), eSteps = input.getConstant("eSteps");
        const double outputFreq = input.getConstant("outputFreq"), cellUpdateFreq = input.getConstant("cellUpdateFreq");
        const double delta = input.getConstant("delta"), mu = input.getConstant("mu");
        const double f = input.getConstant("f"), alpha = input.getConstant("alpha");
        const double A_0 = input.getConstant("A_0"), kappa = input.getConstant("kappa");
        const int seed = static_cast<int>(input.getConstant("seed"));
        const double pressure = input.getConstant("P"), delta_V = input.getConstant("delta_V");
        const std::string potentialName = input.getFilename("potential");
        const std::string out_xyz = input.getFilename("out_xyz"), out_data = input.getFilename("out_data");
        const std::string position_file = input.getFilename("position_file");
        const std::string ensemble_name = input.getFilename("ensemble");
        const int gc_per_cell = input.getConstant("gc_per_cell") > 0 ? (int)input.getConstant("gc_per_cell") : 1;

        Ensemble ensemble = Ensemble::GCMC;
        if (ensemble_name == "NVT") ensemble = Ensemble::NVT;
        else if (ensemble_name == "GCMC" || ensemble_name.empty()) ensemble = Ensemble::GCMC;
        else if (ensemble_name == "NPT") ensemble = Ensemble::NPT;
        else if (ensemble_name == "Gibbs") ensemble = Ensemble::Gibbs;
        else {
            std::cerr << "Error: Unrecognised ensemble" << std::endl;
            return 1;
        }

        std::cout << "\n[Input Summary]\n"
                  << "Lx = " << Lx << ", Ly = " << Ly << ", N = " << N << "\n"
                  << "rcut = " << rcut << ", T = " << T << "\n"
                  << "nSteps = " << nSteps << ", eSteps = " << eSteps << "\n"
                  << "outputFreq = " << outputFreq << ", cellUpdateFreq = " << cellUpdateFreq << "\n"
                  << "potential = " << potentialName << "\n"
                  << "mu = " << mu << ", f = " << f << ", alpha = " << alpha << ", A_0 = " << A_0
                  << ", kappa = " << kappa << "\n"
                  << "seed = " << seed << ", delta = " << delta << ", gc_per_cell = " << gc_per_cell << "\n"
                  << "ensemble = " << (ensemble_name.empty() ? "[default GCMC]" : ensemble_name) << "\n"
                  << "Output XYZ = " << out_xyz << ", Output Data = " << out_data << "\n"
                  << std::endl;

        if (ensemble == Ensemble::Gibbs) {  // serial_src/main.cpp:146-236: a second box from the keys Lx2 Ly2 N2 ...
            const double Lx2 = input.getConstant("Lx2"), Ly2 = input.getConstant("Ly2");
            const int N2 = static_cast<int>(input.getConstant("N2"));
            const std::string out_xyz2 = input.getFilename("out_xyz2"), out_data2 = input.getFilename("out_data2");
            const std::string position_file2 = input.getFilename("position_file2");
            std::cout << "Lx_2 = " << Lx2 << ", Ly_2 = " << Ly2 << ", N_2 = " << N2 << "\nOutput XYZ 2 = " << out_xyz2
                      << ", Output Data 2 = " << out_data2 << "\n" << std::endl;
            SimulationBox box_1(Lx, Ly), box_2(Lx2, Ly2);
            std::vector<Particle> particles_1, particles_2;
            if (!position_file.empty()) initializeParticles_from_file(particles_1, box_1, N, position_file);
            else initializeParticles(particles_1, box_1, N, true, seed * 10);
            if (!position_file2.empty()) initializeParticles_from_file(particles_2, box_2, N2, position_file2);
            else initializeParticles(particles_2, box_2, N2, true, seed * 5);
            Logging logger_1(out_xyz, out_data), logger_2(out_xyz2, out_data2);
            RNG rng(seed);
            const PotentialType pt = selectPotentialType(potentialName);
            ThermodynamicCalculator calc_1(T, pressure, pt, rcut, mu, f, alpha, A_0, kappa);
            ThermodynamicCalculator calc_2(T, pressure, pt, rcut, mu, f, alpha, A_0, kappa);
            GibbsMonteCarlo gmc(box_1, box_2, particles_1, particles_2, calc_1, calc_2, rcut, delta, delta_V, logger_1,
                                logger_2, rng);
            const size_t upd = cellUpdateFreq >= 1 ? static_cast<size_t>(cellUpdateFreq) : 1;
            std::cout << "Equilibration:" << std::endl;
            if (eSteps >= 1) {
                gmc.equlibrateNVT(static_cast<size_t>(eSteps + 1), static_cast<size_t>(eSteps), upd);
                gmc.equlibratemuVT(static_cast<size_t>(eSteps + 1), static_cast<size_t>(eSteps), upd);
            }
            std::cout << "Running:" << std::endl;
            gmc.run(static_cast<size_t>(nSteps), outputFreq >= 1 ? static_cast<size_t>(outputFreq) : 1, upd);
            std::cout << "Simulation completed. N = " << gmc.getNumParticles(0) << " / " << gmc.getNumParticles(1)
                      << ", V = " << gmc.getBox(0).getV() << " / " << gmc.getBox(1).getV() << std::endl;
            return 0;
        }

        SimulationBox box(Lx, Ly);
        std::vector<Particle> particles;
        if (!position_file.empty()) initializeParticles_from_file(particles, box, N, position_file);
        else initializeParticles(particles, box, N, true, seed * 10);

        Logging logger(out_xyz, out_data);
        RNG rng(seed);
        ThermodynamicCalculator calc(T, pressure, selectPotentialType(potentialName), rcut, mu, f, alpha, A_0, kappa);
        MonteCarlo mc(box, particles, calc, rcut, delta, delta_V, ensemble, logger, rng);
        mc.setGcDrawsPerCell(gc_per_cell);

        const size_t upd = cellUpdateFreq >= 1 ? static_cast<size_t>(cellUpdateFreq) : 1;
        if (eSteps >= 1) {
            std::cout << "Equilibration:" << std::endl;
            mc.run(static_cast<size_t>(eSteps), static_cast<size_t>(eSteps * 10), upd);
        }
        std::cout << "Running:" << std::endl;
        mc.run(static_cast<size_t>(nSteps), outputFreq >= 1 ? static_cast<size_t>(outputFreq) : 1, upd);
        const mc2d_stats st = mc.stats();
        std::cout << "Simulation completed. N = " << st.n_particles << ", energy = " << st.energy
                  << ", acceptance = " << mc.acceptanceRatio() << std::endl;
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
