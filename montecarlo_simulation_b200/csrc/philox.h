// philox.h — Philox4x32-10 counter-based RNG (Salmon et al., SC'11), host + device.
//
// Replaces the reference's sequential std::mt19937 (serial_src/rng.cpp) and per-rank
// std::mt19937_64 (mpi_src/rng_parallel.h): every cell owns the stream
// key = seed, counter = (global cell id, trial tag, sweep), so a trajectory does not depend on how
// cells are mapped to warps, CTAs or GPUs, and no seed broadcast (MC_parallel.cpp:218-230) is needed.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MC2D_HD __host__ __device__ __forceinline__
#else
#define MC2D_HD inline
#endif

struct Philox4 {
    uint32_t x, y, z, w;
};

MC2D_HD uint32_t mc2d_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MC2D_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mc2d_mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mc2d_mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o = {c0, c1, c2, c3};
    return o;
}

// uniform in (0,1) from one 32-bit word: (w + 0.5) / 2^32 — symmetric about 1/2, never 0 or 1
MC2D_HD double mc2d_u01(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }

// trial tags (counter word 1): bits 28..31 = stream kind, bits 0..27 = trial index in the cell visit
#define MC2D_TAG_DISP 0x00000000u
#define MC2D_TAG_GC 0x10000000u
#define MC2D_TAG_SCHEDULE 0xF0000000u

// Per-sweep schedule: colour order (Fisher-Yates over the 4 checkerboard colours) and the grid
// shift in [0,dx) x [0,dy).  Pure function of (seed, sweep): identical on every rank.
inline void mc2d_schedule(uint64_t seed, uint64_t sweep, double dx, double dy, int shift_grid, int order[4],
                          double *sx, double *sy) {
    Philox4 r = philox4x32_10(0xFFFFFFFFu, MC2D_TAG_SCHEDULE, (uint32_t)sweep, (uint32_t)(sweep >> 32), (uint32_t)seed,
                              (uint32_t)(seed >> 32));
    order[0] = 0; order[1] = 1; order[2] = 2; order[3] = 3;
    uint32_t w[3] = {r.x, r.y, r.z};
    for (int i = 3; i >= 1; --i) {
        int j = (int)mc2d_mulhi32(w[3 - i], (uint32_t)(i + 1));
        int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    Philox4 s = philox4x32_10(0xFFFFFFFFu, MC2D_TAG_SCHEDULE | 1u, (uint32_t)sweep, (uint32_t)(sweep >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    if (shift_grid) {
        *sx = mc2d_u01(s.x) * dx;
        *sy = mc2d_u01(s.y) * dy;
        if (*sx >= dx) *sx = 0.0;
        if (*sy >= dy) *sy = 0.0;
    } else {
        *sx = 0.0;
        *sy = 0.0;
    }
}
