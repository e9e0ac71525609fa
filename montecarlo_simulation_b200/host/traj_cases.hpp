// traj_cases.hpp — the frames both trajectory-logger drivers write (TEST SCAFFOLDING shared by
// host/slab_traj_selftest.cpp and oracle/traj_ref_main.cpp, so that the two sides cannot drift apart).
// Rank r of n holds 2 r + (r odd ? 0 : 3) particles in frame 0 — an EMPTY rank 1 — and one more per later frame.
#ifndef MC2D_TRAJ_CASES_HPP
#define MC2D_TRAJ_CASES_HPP

#include <vector>

namespace traj_cases {

constexpr int kFrames = 3;
constexpr double kLx = 20.0, kLy = 12.345678901234567;

inline int count(int rank, int frame) { return (rank == 1 ? 0 : 2 * rank + 3) + frame; }

// flat {x0, y0, x1, y1, ...}: values with more than 6 and more than 10 significant digits, a tiny and a large one
inline std::vector<double> frame_xy(int rank, int frame) {
    std::vector<double> xy;
    const int n = count(rank, frame);
    for (int i = 0; i < n; ++i) {
        double x = 0.123456789012345678 + 1.7320508075688772 * i + 3.0 * rank + 0.001 * frame;
        double y = 12.345678901234567 / (1.0 + i + rank) + 1e-7 * frame;
        if (i == 1) x = 1.0e-9 * (rank + 1);
        if (i == 2) y = 12.0;
        xy.push_back(x);
        xy.push_back(y);
    }
    return xy;
}
inline int timestep(int frame) { return 100 * frame; }

}  // namespace traj_cases

#endif
