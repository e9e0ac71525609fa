// step_harness — the bench step (5 GCMC sweeps + 1 energy/virial reduction, config #4) called straight through the
// C ABI from C++, to separate the library's own host overhead from the Python binding's:
//   g++ -O2 -std=c++17 -Iinclude tools/step_harness.cpp -Lmontecarlo_simulation_b200 -lmc2d -Wl,-rpath,$PWD/montecarlo_simulation_b200 -o /tmp/step_harness
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "mc2d.h"

int main(int argc, char **argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 1000000;
    const double L = std::sqrt(N / 0.5);
    mc2d_params p;
    std::memset(&p, 0, sizeof p);
    p.Lx = p.Ly = L;
    p.rcut = 2.5;
    p.potential = MC2D_LENNARD_JONES;
    p.temperature = 1.0;
    p.activity = 0.497;
    p.delta = 0.1;
    p.seed = 42;
    p.nranks = 1;
    mc2d_ctx *ctx = nullptr;
    if (mc2d_create(&p, &ctx)) return 1;
    const long long side = (long long)std::ceil(std::sqrt((double)N));
    std::vector<double> xy;
    std::mt19937_64 gen(1);
    std::uniform_real_distribution<double> u(-0.1, 0.1);
    for (long long i = 0; i < N; ++i) {
        xy.push_back(((i % side) + 0.5 + u(gen)) * L / side);
        xy.push_back(((i / side) + 0.5 + u(gen)) * L / side);
    }
    if (mc2d_upload(ctx, xy.data(), nullptr, N)) { std::fprintf(stderr, "%s\n", mc2d_last_error(ctx)); return 1; }
    mc2d_sweep_plan plan;
    std::memset(&plan, 0, sizeof plan);
    plan.disp_per_particle = 1;
    plan.gc_per_cell = 1;
    plan.shift_grid = 1;
    mc2d_stats st;
    if (mc2d_run_sweeps(ctx, 200, &plan, &st)) { std::fprintf(stderr, "%s\n", mc2d_last_error(ctx)); return 1; }
    double tr = 0, te = 0;
    const int reps = 40;
    for (int i = 0; i < reps + 5; ++i) {
        const auto a = std::chrono::steady_clock::now();
        mc2d_run_sweeps(ctx, 5, &plan, &st);
        const auto b = std::chrono::steady_clock::now();
        double U, W;
        long long n;
        mc2d_total_energy_virial(ctx, 1, 0, &U, &W, &n);
        const auto c = std::chrono::steady_clock::now();
        if (i >= 5) {
            tr += std::chrono::duration<double, std::milli>(b - a).count();
            te += std::chrono::duration<double, std::milli>(c - b).count();
        }
    }
    mc2d_info info;
    mc2d_get_info(ctx, &info);
    std::printf("C ABI: run_sweeps wall %.3f ms (device %.3f)  energy wall %.3f ms (device %.3f)  step %.3f ms\n", tr / reps,
                info.last_sweep_ms, te / reps, info.last_energy_ms, (tr + te) / reps);
    mc2d_destroy(ctx);
    return 0;
}
