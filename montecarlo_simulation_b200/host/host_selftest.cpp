// host_selftest — the reference's own unit tests, restated against the GPU-backed shims.
// Each block cites the reference test it mirrors (unit_test_serial/*.cpp).  Needs a GPU.
// Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <random>

#include "mc2d_host.hpp"

static int failures = 0, checks = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        ++checks;                                                            \
        if (!(cond)) {                                                       \
            ++failures;                                                      \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
        }                                                                    \
    } while (0)
static bool approx(double a, double b, double eps = 1.2e-5) {
    return std::fabs(a - b) <= eps * std::max(std::fabs(a), std::fabs(b)) + 1e-12;
}

int main() {
    // ---- test_cell_list.cpp:10-31 — neighbour across the periodic boundary ----
    {
        SimulationBox box(3.0, 3.0);
        std::vector<Particle> P = {Particle(0.1, 0.1), Particle(2.9, 2.9), Particle(1.6, 1.5)};
        CellList cl(box, 0.5);
        cl.build(P);
        auto nb0 = cl.getNeighbors(0, P);
        CHECK(nb0.size() == 1);
        CHECK(nb0.size() == 1 && nb0[0].first == 1 && approx(nb0[0].second, 0.08, 1e-6));
        CHECK(cl.getNeighbors(2, P).empty());
    }
    // ---- test_cell_list.cpp:32-53 — exact cutoff inclusion ----
    {
        SimulationBox box(5.0, 5.0);
        std::vector<Particle> P = {Particle(1.0, 1.0), Particle(2.0, 1.0), Particle(3.1, 1.0)};
        CellList cl(box, 1.0);
        cl.build(P);
        auto nb0 = cl.getNeighbors(0, P);
        CHECK(nb0.size() == 1 && nb0[0].first == 1 && nb0[0].second == 1.0);
        CHECK(cl.getNeighbors(2, P).empty());
    }
    // ---- test_cell_list.cpp:54-81 — several neighbours ----
    {
        SimulationBox box(6.0, 6.0);
        std::vector<Particle> P = {Particle(0.5, 0.5), Particle(1.0, 1.0), Particle(2.0, 0.5), Particle(3.5, 3.5)};
        CellList cl(box, 1.5);
        cl.build(P);
        auto nb0 = cl.getNeighbors(0, P);
        CHECK(nb0.size() == 2);
        CHECK(cl.getNeighbors(3, P).empty());
        auto nbp = cl.getNeighbors2(Particle(0.75, 0.75), P);
        CHECK(nbp.size() == 3);
    }
    // ---- test_calc.cpp:16-50 — few-particle closed forms (brute force = GPU O(N^2)) ----
    {
        SimulationBox box(10.0, 10.0);
        std::vector<Particle> P(5, Particle(1.0, 1.0));
        ThermodynamicCalculator calc(1.0, 0.0, PotentialType::Ideal, 1.0);
        CHECK(calc.computeTotalEnergy(P, box) == 0.0);
        CHECK(approx(calc.computePressure(P, box), 0.05));
        const double r0 = std::pow(2.0, 1.0 / 6.0) - 1e-14;
        SimulationBox b2(r0 * 6, r0 * 6);
        std::vector<Particle> D = {Particle(0.0, 0.0), Particle(r0, 0.0)};
        ThermodynamicCalculator lj(1.0, 0.0, PotentialType::LennardJones, 3 * r0);
        CHECK(approx(lj.computeTotalEnergy(D, b2), -1.0));
        CHECK(approx(lj.computeTotalVirial(D, b2) + 1, 1.0));
        SimulationBox b3(r0 * 10, r0 * 10);
        std::vector<Particle> T3 = {Particle(0, 0), Particle(r0, 0), Particle(r0 / 2, r0 * std::sqrt(3.0) / 2)};
        ThermodynamicCalculator lj5(1.0, 0.0, PotentialType::LennardJones, 5 * r0);
        CHECK(approx(lj5.computeTotalEnergy(T3, b3), -3.0));
        SimulationBox b4(3.0, 3.0);
        std::vector<Particle> Y = {Particle(0, 0), Particle(1.0, 0)};
        ThermodynamicCalculator yk(1.0, 0.0, PotentialType::Yukawa, 5.0, 1.0, 1.0);
        CHECK(approx(yk.computeTotalEnergy(Y, b4), std::exp(-1.0)));
        CHECK(approx(yk.computeTotalVirial(Y, b4), 2 * std::exp(-1.0)));
    }
    // ---- test_calc.cpp:84-120 — star-polymer lattices ----
    {
        SimulationBox box(4.0, 8.0);
        std::vector<Particle> P;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 8; ++j) P.emplace_back(i * 1.0, j * 1.0);
        const double f = 8.0, A_0 = 2.0, kappa = 1.2, alpha = 0.3;
        ThermodynamicCalculator calc(1.0, 0.0, PotentialType::ThermalStar, 1.4, 0.0, f, alpha, A_0, kappa);
        const double e2 = A_0 * f * f * std::exp(-kappa) / kappa;
        const double Upair = (f * f * 9 + 2) / 24 * alpha - e2, Wpair = (f * f * 9 + 2) / 24 - A_0 * f * f * std::exp(-kappa);
        const double pairs = P.size() * 4.0 / 2.0;
        CHECK(approx(calc.computeTotalEnergy(P, box), pairs * Upair));
        CHECK(approx(calc.computeTotalVirial(P, box), pairs * Wpair));
        CHECK(approx(calc.computePressure(P, box), 1.0 + pairs * Wpair / box.getV() / 2));
    }
    // ---- test_thermodynamic_celllist.cpp:6-28 — cell list == brute force at 1e-12, N = 5000 ----
    {
        const size_t N = 5000;
        SimulationBox box(100.0, 10.0);
        std::mt19937_64 rng(123);
        std::uniform_real_distribution<double> dx(0.0, 100.0), dy(0.0, 10.0);
        std::vector<Particle> P;
        for (size_t i = 0; i < N; ++i) {
            const double x = dx(rng), y = dy(rng);
            P.emplace_back(x, y);
        }
        ThermodynamicCalculator calc(1.0, 0.0, PotentialType::LennardJones, 2.5);
        const double Ed = calc.computeTotalEnergy(P, box), Wd = calc.computeTotalVirial(P, box);
        const double Ec = calc.computeTotalEnergyCellList(P, box), Wc = calc.computeTotalVirialCellList(P, box);
        CHECK(approx(Ec, Ed, 1e-12));
        CHECK(approx(Wc, Wd, 1e-12));
    }
    // ---- MonteCarlo through the reference's constructor: NVT WCA relaxes, N conserved ----
    // (test_mc_parallel.cpp:466-538 idea: energy drops during equilibration from a random start)
    {
        SimulationBox box(40.0, 40.0);
        std::vector<Particle> P;
        initializeParticles(P, box, 400, true, 7);
        ThermodynamicCalculator calc(1.0, 0.0, PotentialType::WCA, std::pow(2.0, 1.0 / 6.0));
        Logging logger("/dev/null", "/dev/null");
        RNG rng(5);
        const double E0 = calc.computeTotalEnergyCellList(P, box);
        MonteCarlo mc(box, P, calc, std::pow(2.0, 1.0 / 6.0), 0.1, 0.0, Ensemble::NVT, logger, rng);
        CHECK(approx(mc.getEnergy(), E0, 1e-10));
        mc.run(200, 100, 10);
        CHECK(mc.getParticles().size() == 400);
        CHECK(mc.getEnergy() < E0);
        CHECK(approx(mc.getEnergy(), calc.computeTotalEnergyCellList(mc.getParticles(), box), 1e-9));
    }
    // ---- GCMC ideal gas through the reference's constructor: <N> = zV (docs/test_scenarios.md:13-56) ----
    {
        const double Lx = 10 * std::sqrt(std::sqrt(3.0) / 2), Ly = 10 / std::sqrt(std::sqrt(3.0) / 2);
        SimulationBox box(Lx, Ly);
        std::vector<Particle> P;
        initializeParticles(P, box, 50, true, 3);
        ThermodynamicCalculator calc(1.0, 0.0, PotentialType::Ideal, 2.5, 0.5);
        Logging logger("/dev/null", "/dev/null");
        RNG rng(11);
        MonteCarlo mc(box, P, calc, 2.5, 0.1, 0.0, Ensemble::GCMC, logger, rng);
        mc.setGcDrawsPerCell(4);
        mc.run(200, 1000, 10);
        double sum = 0;
        const int samples = 2000;
        for (int s = 0; s < samples; ++s) {
            mc.run(5, 1000000, 1000000);
            sum += double(mc.stats().n_particles);
        }
        const double mean = sum / samples;
        std::printf("ideal gas z=0.5 V=100: <N> = %.3f (theory 50)\n", mean);
        CHECK(std::fabs(mean - 50.0) < 1.0);
    }
    // ---- Gibbs ensemble, ideal gas, exchange moves only (equlibratemuVT): N_1 is binomial(N, V_1 / V) ----
    {
        SimulationBox b1(30.0, 30.0), b2(30.0, 20.0);
        std::vector<Particle> P1, P2;
        initializeParticles(P1, b1, 300, true, 3);
        initializeParticles(P2, b2, 300, true, 4);
        ThermodynamicCalculator c1(1.0, 0.0, PotentialType::Ideal, 2.5), c2(1.0, 0.0, PotentialType::Ideal, 2.5);
        Logging l1("/dev/null", "/dev/null"), l2("/dev/null", "/dev/null");
        RNG rng(21);
        GibbsMonteCarlo g(b1, b2, P1, P2, c1, c2, 2.5, 0.3, 0.05, l1, l2, rng);
        g.equlibratemuVT(4000, 1000000, 1000000);
        double s1 = 0, s2 = 0;
        const int samples = 3000;
        for (int s = 0; s < samples; ++s) {
            g.equlibratemuVT(20, 1000000, 1000000);
            const double n1 = (double)g.getNumParticles(0);
            s1 += n1;
            s2 += n1 * n1;
            CHECK(g.getNumParticles(0) + g.getNumParticles(1) == 600);
        }
        const double mean = s1 / samples, var = s2 / samples - mean * mean;
        const double p = 900.0 / 1500.0;
        std::printf("Gibbs ideal gas: <N_1> = %.2f (theory %.1f), Var N_1 = %.1f (theory %.1f)\n", mean, 600 * p, var,
                    600 * p * (1 - p));
        CHECK(std::fabs(mean - 600 * p) < 5.0);
        CHECK(std::fabs(var / (600 * p * (1 - p)) - 1.0) < 0.25);
    }
    // ---- Gibbs ensemble, LJ: full loop (sweeps, exchanges, coupled volume changes); N and V are conserved, the
    //      running energies track a recomputation on the downloaded configurations ----
    {
        SimulationBox b1(24.0, 24.0), b2(24.0, 24.0);
        std::vector<Particle> P1, P2;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < 16; ++j) {
                P1.emplace_back((i + 0.5) * 1.5, (j + 0.5) * 1.5);
                if ((i + j) % 3 == 0) P2.emplace_back((i + 0.5) * 1.5, (j + 0.5) * 1.5);
            }
        const long long Ntot = (long long)(P1.size() + P2.size());
        ThermodynamicCalculator c1(1.5, 0.0, PotentialType::LennardJones, 2.5), c2(1.5, 0.0, PotentialType::LennardJones, 2.5);
        Logging l1("/dev/null", "/dev/null"), l2("/dev/null", "/dev/null");
        RNG rng(33);
        GibbsMonteCarlo g(b1, b2, P1, P2, c1, c2, 2.5, 0.15, 0.05, l1, l2, rng);
        g.equlibrateNVT(50, 1000000, 10);
        g.run(3000, 1000000, 1000000);
        CHECK(g.getNumParticles(0) + g.getNumParticles(1) == Ntot);
        CHECK(approx(g.getBox(0).getV() + g.getBox(1).getV(), 2 * 576.0, 1e-9));
        for (int w = 0; w < 2; ++w) {
            ThermodynamicCalculator &c = w ? c2 : c1;
            SimulationBox bw = g.getBox(w);
            const double U = c.computeTotalEnergyCellList(g.getParticles(w), bw);
            CHECK(approx(g.getEnergy(w), U, 1e-8));
            CHECK((long long)g.getParticles(w).size() == g.getNumParticles(w));
        }
        std::printf("Gibbs LJ T=1.5: N = %lld / %lld, V = %.1f / %.1f, rho = %.3f / %.3f\n", g.getNumParticles(0),
                    g.getNumParticles(1), g.getBox(0).getV(), g.getBox(1).getV(), g.getNumParticles(0) / g.getBox(0).getV(),
                    g.getNumParticles(1) / g.getBox(1).getV());
    }
    std::printf("%d checks, %d failures\n", checks, failures);
    return failures ? 1 : 0;
}
