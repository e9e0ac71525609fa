// sweep_kernel.cuh — K2 + K3: one colour pass of the checkerboard sweep (trial displacements and
// GCMC insert/delete trials for every cell of one colour).  Included by mc2d_kernels.cuh.
//
//   replaces displacementMove_cell_list_dE (MC.cpp:96-123), grandCanonicalMove (MC.cpp:189-249),
//   getNeighbors/getNeighbors2 (cell_list.cpp:43-101) and, on the MPI side, do_one_parity_step_ /
//   try_displacement_ / local_energy_of_* (MC_parallel.cpp:117-216).
//
// One warp per active cell.  The 8 surrounding cells (frozen during the pass: same-colour cells are
// >= one cell >= rcut apart) are staged once into shared memory as coordinates LOCAL to the active
// cell (periodic image already applied: the inner loop has no minimum-image code), followed in the
// same array by the cell's own particles, which are updated in place:
//       cand[0 .. M)            neighbours, frozen
//       cand[M .. M + n_own)    the cell's own particles (insertions append, deletions swap-with-last)
// Each trial: the random words come from a per-cell Philox stream (lanes generate the words of 32
// consecutive trials at once and broadcast them by shuffle); every lane evaluates the pair energy of
// ITS candidates for the old and the new position; a warp-shuffle reduction gives dE; every lane
// takes the same Metropolis decision.  Detailed balance: a displacement that leaves the cell is
// rejected; insertions are uniform in the cell with acceptance z*A_cell/(n+1)*exp(-beta dU),
// deletions n/(z*A_cell)*exp(-beta dU) — the per-cell form of MC.cpp:201,226.
// MINIMG = true adds a per-pair minimum image (grids with fewer than 4 cells across).
#pragma once

struct SweepArgs {
    double2 *pos;
    int *cnt;
    int *ids;
    const int *order;  // k_sweep_p: perm[colour][column pair][rank] (k_order_local)
    int *work;         // k_sweep_p: work-item ticket counter of this pass
    int nitems;        // k_sweep_p: work items of this launch = items per column pair x pair_n
    int pair_lo, pair_n;  // k_sweep_p: the column pairs this launch covers (a slab's boundary pair runs apart)
    int npass;            // k_sweep_p: colour passes done by this launch: 1, or 4 with a grid barrier in between
    int colours[4];       //            their colours (npass > 1); pass p uses work[p] and partial_dU + p * dU_stride
    int dU_stride;
    int pipeline;         // k_sweep_p, npass = 4, single rank: no grid barrier; items of all colours in one list
    int *done;            //            (colour-major, then column pair), done[pass][pair] finished items gate the next colour
    // k_sweep_p, slab mode with all passes in one launch: the halo is part of the kernel (sweep_kernel_p.cuh)
    int halo_fused;
    double2 *hpos[2];                // [0] left neighbour's right-ghost column, [1] right neighbour's left-ghost column
    int *hcnt[2];
    int *hids[2];
    unsigned long long *hflag[2];    // the neighbours' flag words this rank signals
    unsigned long long *myflag;      // [0] from left, [1] from right, [3] finished boundary items of the pass
    unsigned long long seq_base;     // sync sequence number reached before this launch
    Grid g;
    PairParams pp;
    double r2cut, beta, zA, delta;
    int colour, pass;
    uint32_t sweep_lo, sweep_hi, seed_lo, seed_hi;
    int disp_mult, disp_cell, n_gc;
    double p_ins, p_del;
    double *partial_dU;            // one accumulator per CTA, private to it, summed at the end of the batch
    unsigned long long *counters;  // [0..2] attempted, [3..5] accepted, [6] overflow
    int *err;                      // ctx->d_err: [4] a wait on a flag word gave up, [5] insertions refused in THIS call
    unsigned long long spin_budget_ns;  // budget of one wait on a flag word (mc2d_tuning.spin_timeout_ms)
    mc2d_trial *trace;
    unsigned long long *trace_n;
    long long trace_cap;
};

// 1/x to ~1 ulp without the IEEE-division slow path: hardware seed (rcp.approx.ftz.f64, ~2^-23)
// plus two Newton steps.  x = 0 gives NaN, which fails every acceptance test (a trial position
// exactly on top of another particle is rejected).  Only the sweep kernel uses it; the parity
// hooks (K4/K6) keep the IEEE division of the reference.
#ifndef MC2D_RCP_NEWTON_STEPS
#define MC2D_RCP_NEWTON_STEPS 2
#endif
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
    for (int it = 0; it < MC2D_RCP_NEWTON_STEPS; ++it) {
        const double e = fma(-x, r, 1.0);
        r = fma(r, e, r);
    }
    return r;
}

// Metropolis test in log form.  "u < exp(-x)" (MC.cpp:110-115,201,226) is "-ln u > x": the transcendental then
// depends on the random word only, not on the energy difference, so k_sweep_p evaluates it off the critical
// path (once per batch of Philox words, one lane per trial) and the decision itself is one compare.  Both sweep
// kernels use exactly these expressions, so they take the same decisions.
__device__ __forceinline__ double mc2d_exp_variate(uint32_t w) { return -log(mc2d_u01(w)); }

// pair energy with the inclusive cutoff; LJ / WCA take the fast reciprocal
template <int POT>
__device__ __forceinline__ double pair_e_cut(double dx, double dy, double r2cut, const PairParams &pp) {
    const double r2 = fma(dx, dx, dy * dy);
    if (!(r2 <= r2cut)) return 0.0;
    if (POT == POT_LJ) {
        const double r2inv = fast_rcp(r2);
        const double r6inv = r2inv * r2inv * r2inv;
        return 4.0 * r6inv * (r6inv - 1.0);
    } else if (POT == POT_WCA) {
        if (r2 < 1.2599) {
            const double r2inv = fast_rcp(r2);
            const double r6inv = r2inv * r2inv * r2inv;
            return 4.0 * r6inv * (r6inv - 1.0) + 1.0;
        }
        return 0.0;
    } else {
        return pair_u<POT>(r2, pp);
    }
}

__device__ __forceinline__ void wrap_pair(double &dx, double &dy, const Grid &g, double hLx, double hLy) {
    if (dx >= hLx) dx -= g.Lx; else if (dx < -hLx) dx += g.Lx;
    if (dy >= hLy) dy -= g.Ly; else if (dy < -hLy) dy += g.Ly;
}

// 4 CTAs of 8 warps per SM (64 registers, no spills): the kernel is latency bound (RNG -> distances
// -> reciprocal -> shuffle reduction -> exp is one dependent chain per trial), occupancy pays:
// measured 13.6 ms vs 15.9 ms per 20 sweeps of config #4 against 3 CTAs/SM (profiles/r01_notes.md).
#ifndef MC2D_SWEEP_MIN_BLOCKS
#define MC2D_SWEEP_MIN_BLOCKS 4
#endif

template <int POT, bool MINIMG, bool TRACE>
__global__ void __launch_bounds__(256, MC2D_SWEEP_MIN_BLOCKS) k_sweep(SweepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    __shared__ unsigned int cred[8][7];
    const Grid &g = a.g;
    const int K = g.K;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    // per-warp shared memory: cand[(8K or 0) + K] double2, own_id[K] int
    const int nb_cap = (POT == POT_IDEAL) ? 0 : 8 * K;
    const size_t per_warp = (size_t)(K + nb_cap) * sizeof(double2) + (size_t)K * sizeof(int);
    double2 *cand = reinterpret_cast<double2 *>(smem_raw + wib * per_warp);
    int *own_id = reinterpret_cast<int *>(cand + nb_cap + K);

    const int nay = g.ny >> 1;
    const int nact = (g.ncol >> 1) * nay;
    double dU_total = 0.0;
    unsigned int c_att[3] = {0, 0, 0}, c_acc[3] = {0, 0, 0}, c_ovf = 0;

    if (w < nact) {
        const int px = a.colour & 1, py = a.colour >> 1;
        const int acx = w / nay, acy = w - acx * nay;
        const int lcx = g.ghost + 2 * acx + px;  // stored column (col0 is even, so parity is global)
        const int cy = 2 * acy + py;
        const int gcx = g.col0 + 2 * acx + px;
        const int cell = lcx * g.ny + cy;
        const uint32_t cell_gid = (uint32_t)(gcx * g.ny + cy);
        const double x_lo = g.ox + gcx * g.dx, y_lo = g.oy + cy * g.dy;
        const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;
        int n_own = a.cnt[cell];
        const int n_disp = n_own > 0 ? (a.disp_cell > 0 ? a.disp_cell : a.disp_mult * n_own) : 0;
        if (n_disp > 0 || a.n_gc > 0) {
            // ---- stage the 8 neighbour cells, flattened, into cand[0..M) ----
            int M = 0;
            if (POT != POT_IDEAL) {
                int nb_cell = -1, nb_n = 0;
                if (lane < 8) {
                    const int o = lane + (lane >= 4);  // skip (0,0)
                    const int ddx = o / 3 - 1, ddy = o % 3 - 1;
                    const bool dup = (g.nx == 2 && ddx < 0) || (g.ny == 2 && ddy < 0);
                    if (!dup) {
                        const int ncx = g.ghost ? lcx + ddx : (lcx + ddx + g.nx) % g.nx;
                        const int ncy = (cy + ddy + g.ny) % g.ny;
                        nb_cell = ncx * g.ny + ncy;
                        if (nb_cell != cell) nb_n = a.cnt[nb_cell];
                    }
                }
                int incl = nb_n;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    int t = __shfl_up_sync(FULL_MASK, incl, d);
                    if (lane >= d) incl += t;
                }
                M = __shfl_sync(FULL_MASK, incl, 7);
                const int excl = incl - nb_n;
                for (int k0 = 0; k0 < M; k0 += 32) {  // uniform trip count: the shuffles need all lanes
                    const int k = k0 + lane;
                    // owner of candidate k: the last of the 8 cells whose exclusive prefix is <= k
                    int cc = __shfl_sync(FULL_MASK, nb_cell, 0), cs = 0;
#pragma unroll
                    for (int q = 1; q < 8; ++q) {
                        const int tq = __shfl_sync(FULL_MASK, excl, q), cq = __shfl_sync(FULL_MASK, nb_cell, q);
                        if (k >= tq) { cc = cq; cs = tq; }
                    }
                    if (k < M) {
                        const double2 p = a.pos[(size_t)cc * K + (k - cs)];
                        double ux = p.x - x_lo, uy = p.y - y_lo;
                        if (ux < -hLx) ux += g.Lx; else if (ux >= hLx) ux -= g.Lx;
                        if (uy < -hLy) uy += g.Ly; else if (uy >= hLy) uy -= g.Ly;
                        cand[k] = make_double2(ux, uy);
                    }
                }
            }
            double2 *own = cand + M;  // own particles follow the neighbours
            for (int s = lane; s < n_own; s += 32) {
                const double2 p = a.pos[(size_t)cell * K + s];
                double ux = p.x - x_lo, uy = p.y - y_lo;
                if (ux < -hLx) ux += g.Lx; else if (ux >= hLx) ux -= g.Lx;
                if (uy < -hLy) uy += g.Ly; else if (uy >= hLy) uy -= g.Ly;
                own[s] = make_double2(ux, uy);
                if (a.ids) own_id[s] = a.ids[(size_t)cell * K + s];
            }
            __syncwarp();
            bool dirty = false;
            Philox4 batch = {0u, 0u, 0u, 0u};

            // ---- K2: displacement trials (MC.cpp:96-123) ----
            for (int t = 0; t < n_disp; ++t) {
                if ((t & 31) == 0)  // lane l draws the words of trial t + l
                    batch = philox4x32_10(cell_gid, MC2D_TAG_DISP | (uint32_t)(t + lane), a.sweep_lo, a.sweep_hi,
                                          a.seed_lo, a.seed_hi);
                Philox4 r;
                r.x = __shfl_sync(FULL_MASK, batch.x, t & 31);
                r.y = __shfl_sync(FULL_MASK, batch.y, t & 31);
                r.z = __shfl_sync(FULL_MASK, batch.z, t & 31);
                r.w = __shfl_sync(FULL_MASK, batch.w, t & 31);
                const int i = (int)__umulhi(r.x, (uint32_t)n_own);
                const double2 uo = own[i];
                double2 un;
                un.x = uo.x + a.delta * (2.0 * mc2d_u01(r.y) - 1.0);
                un.y = uo.y + a.delta * (2.0 * mc2d_u01(r.z) - 1.0);
                ++c_att[0];
                const bool inside = (un.x >= 0.0 && un.x < g.dx && un.y >= 0.0 && un.y < g.dy);
                double dE = 0.0;
                bool accept = false;
                if (inside) {
                    if (POT != POT_IDEAL) {
                        double acc = 0.0;
                        const int ntot = M + n_own, self = M + i;
                        for (int k = lane; k < ntot; k += 32) {
                            if (k == self) continue;
                            const double2 c = cand[k];
                            double dxn = c.x - un.x, dyn = c.y - un.y, dxo = c.x - uo.x, dyo = c.y - uo.y;
                            if (MINIMG) {
                                wrap_pair(dxn, dyn, g, hLx, hLy);
                                wrap_pair(dxo, dyo, g, hLx, hLy);
                            }
                            acc += pair_e_cut<POT>(dxn, dyn, a.r2cut, a.pp) - pair_e_cut<POT>(dxo, dyo, a.r2cut, a.pp);
                        }
                        dE = warp_sum(acc);
                    }
                    accept = a.beta * dE < mc2d_exp_variate(r.w);  // MC.cpp:110-115 in log form (below)
                    if (accept) {
                        __syncwarp();
                        if (lane == 0) own[i] = un;
                        dU_total += dE;
                        ++c_acc[0];
                        dirty = true;
                    }
                }
                if (TRACE && lane == 0) {
                    unsigned long long slot = atomicAdd(a.trace_n, 1ull);
                    if ((long long)slot < a.trace_cap) {
                        mc2d_trial tr;
                        tr.pass = a.pass; tr.cell = (int)cell_gid; tr.seq = t; tr.kind = 0;
                        tr.accepted = accept ? 1 : 0; tr.reserved = inside ? 1 : 0;
                        double ox_ = x_lo + uo.x, oy_ = y_lo + uo.y, nx_ = x_lo + un.x, ny_ = y_lo + un.y;
                        if (ox_ >= g.Lx) ox_ -= g.Lx; if (oy_ >= g.Ly) oy_ -= g.Ly;
                        if (nx_ >= g.Lx) nx_ -= g.Lx; if (ny_ >= g.Ly) ny_ -= g.Ly;
                        tr.old_x = ox_; tr.old_y = oy_; tr.new_x = nx_; tr.new_y = ny_;
                        tr.dE = inside ? dE : nan("");
                        a.trace[slot] = tr;
                    }
                }
                __syncwarp();
            }

            // ---- K3: grand-canonical trials (MC.cpp:189-249), restricted to the active cell ----
            for (int t = 0; t < a.n_gc; ++t) {
                if ((t & 31) == 0)
                    batch = philox4x32_10(cell_gid, MC2D_TAG_GC | (uint32_t)(t + lane), a.sweep_lo, a.sweep_hi,
                                          a.seed_lo, a.seed_hi);
                Philox4 r;
                r.x = __shfl_sync(FULL_MASK, batch.x, t & 31);
                r.y = __shfl_sync(FULL_MASK, batch.y, t & 31);
                r.z = __shfl_sync(FULL_MASK, batch.z, t & 31);
                r.w = __shfl_sync(FULL_MASK, batch.w, t & 31);
                const double sel = mc2d_u01(r.x);
                int kind = -1;
                bool accept = false, evaluated = false;
                double dE = 0.0;
                double2 pt = make_double2(0.0, 0.0), old = make_double2(0.0, 0.0);
                if (sel < a.p_ins) {  // insertion
                    kind = 1;
                    ++c_att[1];
                    pt.x = mc2d_u01(r.y) * g.dx;
                    pt.y = mc2d_u01(r.z) * g.dy;
                    if (pt.x >= g.dx) pt.x = 0.0;
                    if (pt.y >= g.dy) pt.y = 0.0;
                    if (n_own >= K) {
                        ++c_ovf;
                    } else {
                        if (POT != POT_IDEAL) {
                            double acc = 0.0;
                            const int ntot = M + n_own;
                            for (int k = lane; k < ntot; k += 32) {
                                const double2 c = cand[k];
                                double dxn = c.x - pt.x, dyn = c.y - pt.y;
                                if (MINIMG) wrap_pair(dxn, dyn, g, hLx, hLy);
                                acc += pair_e_cut<POT>(dxn, dyn, a.r2cut, a.pp);
                            }
                            dE = warp_sum(acc);
                        }
                        evaluated = true;
                        // MC.cpp:201: u < z A / (n + 1) exp(-beta dU), in log form
                        accept = a.beta * dE - (log(a.zA) - log((double)(n_own + 1))) < mc2d_exp_variate(r.w);
                        if (accept) {
                            __syncwarp();
                            if (lane == 0) {
                                own[n_own] = pt;
                                if (a.ids) own_id[n_own] = -1;
                            }
                            ++n_own;
                            dU_total += dE;
                            ++c_acc[1];
                            dirty = true;
                        }
                    }
                } else if (sel < a.p_ins + a.p_del) {  // deletion
                    kind = 2;
                    ++c_att[2];
                    if (n_own > 0) {
                        const int i = (int)__umulhi(r.y, (uint32_t)n_own);
                        old = own[i];
                        if (POT != POT_IDEAL) {
                            double acc = 0.0;
                            const int ntot = M + n_own, self = M + i;
                            for (int k = lane; k < ntot; k += 32) {
                                if (k == self) continue;
                                const double2 c = cand[k];
                                double dxo = c.x - old.x, dyo = c.y - old.y;
                                if (MINIMG) wrap_pair(dxo, dyo, g, hLx, hLy);
                                acc += pair_e_cut<POT>(dxo, dyo, a.r2cut, a.pp);
                            }
                            dE = -warp_sum(acc);  // MC.cpp:225
                        }
                        evaluated = true;
                        // MC.cpp:226: u < n / (z A) exp(-beta dU), in log form
                        accept = a.beta * dE - (log((double)n_own) - log(a.zA)) < mc2d_exp_variate(r.w);
                        if (accept) {
                            __syncwarp();
                            if (lane == 0) {
                                own[i] = own[n_own - 1];
                                if (a.ids) own_id[i] = own_id[n_own - 1];
                            }
                            --n_own;
                            dU_total += dE;
                            ++c_acc[2];
                            dirty = true;
                        }
                    }
                }
                if (TRACE && lane == 0 && kind > 0) {
                    unsigned long long slot = atomicAdd(a.trace_n, 1ull);
                    if ((long long)slot < a.trace_cap) {
                        mc2d_trial tr;
                        tr.pass = a.pass; tr.cell = (int)cell_gid; tr.seq = n_disp + t; tr.kind = kind;
                        tr.accepted = accept ? 1 : 0; tr.reserved = evaluated ? 1 : 0;
                        double ox_ = x_lo + old.x, oy_ = y_lo + old.y, nx_ = x_lo + pt.x, ny_ = y_lo + pt.y;
                        if (ox_ >= g.Lx) ox_ -= g.Lx; if (oy_ >= g.Ly) oy_ -= g.Ly;
                        if (nx_ >= g.Lx) nx_ -= g.Lx; if (ny_ >= g.Ly) ny_ -= g.Ly;
                        tr.old_x = ox_; tr.old_y = oy_; tr.new_x = nx_; tr.new_y = ny_;
                        tr.dE = evaluated ? dE : nan("");
                        a.trace[slot] = tr;
                    }
                }
                __syncwarp();
            }

            // ---- write the cell back (absolute coordinates) ----
            if (dirty) {
                for (int s = lane; s < n_own; s += 32) {
                    const double2 u = own[s];
                    double x = x_lo + u.x, y = y_lo + u.y;
                    if (x >= g.Lx) x -= g.Lx; else if (x < 0.0) x += g.Lx;
                    if (y >= g.Ly) y -= g.Ly; else if (y < 0.0) y += g.Ly;
                    a.pos[(size_t)cell * K + s] = make_double2(x, y);
                    if (a.ids) a.ids[(size_t)cell * K + s] = own_id[s];
                }
                if (lane == 0) a.cnt[cell] = n_own;
            }
        }
    }

    // ---- CTA epilogue: deterministic dU partial, integer counters ----
    if (lane == 0) {
        red[wib] = dU_total;
        cred[wib][0] = c_att[0]; cred[wib][1] = c_att[1]; cred[wib][2] = c_att[2];
        cred[wib][3] = c_acc[0]; cred[wib][4] = c_acc[1]; cred[wib][5] = c_acc[2];
        cred[wib][6] = c_ovf;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < wpb; ++i) s += red[i];
        a.partial_dU[blockIdx.x] += s;  // this CTA slot is private; launches are stream-ordered
    }
    if (threadIdx.x < 7) {
        unsigned long long s = 0;
        for (int i = 0; i < wpb; ++i) s += cred[i][threadIdx.x];
        if (s) atomicAdd(&a.counters[threadIdx.x], s);
        if (s && threadIdx.x == 6) atomicAdd(&a.err[5], (int)s);
    }
}
