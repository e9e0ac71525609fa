// sweep_kernel_g.cuh — K2 + K3, production variant: G lanes per active cell, 32/G cells per warp.
//
// Why: profiles/r01_notes.md.  With one warp per cell every trial pays ~300 warp instructions of
// fixed cost (Philox, index arithmetic, Metropolis test, a 5-step shuffle reduction, branch
// convergence) replicated over 32 lanes, while the ~28 candidates of a 2.5-sigma LJ cell at
// rho = 0.5 fill the lanes once.  Here a warp advances 32/G cells in lock step: the fixed cost is
// shared by 32/G trials, the candidate loop runs ceil(28/G) iterations with the same lane
// efficiency, the reduction needs log2(G) steps.  The warp never diverges in the trial loop: a
// group that has run out of trials (or whose trial left the cell) keeps executing with its lanes
// masked, so all shuffles use the full mask; the cutoff test is a select on r^2, not a branch.
//
// Same algorithm, same Philox streams (global cell id, trial index) and the same arithmetic as
// k_sweep (sweep_kernel.cuh): a run gives bit-identical configurations with either kernel.
// Per group, shared memory holds cand[0..Ms) = the 8 neighbour cells in cell-local coordinates,
// followed by cand[Ms..Ms+n_own) = the cell's own particles.  A cell with more than nb_cap
// neighbours keeps only its own particles in shared memory (Ms = 0) and reads the neighbours from
// global memory in a second, cold loop.  Grids narrower than 4 cells take k_sweep<MINIMG>.
#pragma once

// 3 CTAs/SM: 80 registers, no spills (at 64 registers the kernel spills and re-reads the thread id
// inside the candidate loop); shared memory allows 3 CTAs at G = 4 anyway.
// Measured on config #4, 20 sweeps: 7.8 ms (3 CTAs) vs 8.4 ms (4 CTAs).
#ifndef MC2D_SWEEPG_MIN_BLOCKS
#define MC2D_SWEEPG_MIN_BLOCKS 3
#endif

// neighbour q (0..7, skipping the centre) of stored cell (lcx, cy): stored cell index
__device__ __forceinline__ int nbr_cell_index(const Grid &g, int lcx, int cy, int q) {
    const int o = q + (q >= 4);
    const int ddx = o / 3 - 1, ddy = o % 3 - 1;
    int ncx = lcx + ddx, ncy = cy + ddy;
    if (!g.ghost) {
        if (ncx < 0) ncx += g.nx; else if (ncx >= g.nx) ncx -= g.nx;
    }
    if (ncy < 0) ncy += g.ny; else if (ncy >= g.ny) ncy -= g.ny;
    return ncx * g.ny + ncy;
}

__device__ __forceinline__ double2 to_local(double2 p, double x_lo, double y_lo, const Grid &g, double hLx,
                                            double hLy) {
    double ux = p.x - x_lo, uy = p.y - y_lo;
    if (ux < -hLx) ux += g.Lx; else if (ux >= hLx) ux -= g.Lx;
    if (uy < -hLy) uy += g.Ly; else if (uy >= hLy) uy -= g.Ly;
    return make_double2(ux, uy);
}

// candidate k of a cell whose neighbours did not fit in shared memory: walk the 8 cells.
// Scalar arguments only: taking the address of the kernel parameter struct would force a copy of it
// to the local stack in every thread.
__device__ __noinline__ double2 slow_candidate(const double2 *pos, const int *cnt, int ghost, int nx, int ny, int K,
                                               int lcx, int cy, int k, double x_lo, double y_lo, double Lx,
                                               double Ly) {
    for (int q = 0; q < 8; ++q) {
        const int o = q + (q >= 4);
        int ncx = lcx + o / 3 - 1, ncy = cy + o % 3 - 1;
        if (!ghost) {
            if (ncx < 0) ncx += nx; else if (ncx >= nx) ncx -= nx;
        }
        if (ncy < 0) ncy += ny; else if (ncy >= ny) ncy -= ny;
        const int c = ncx * ny + ncy;
        const int n = cnt[c];
        if (k < n) {
            const double2 p = pos[(size_t)c * K + k];
            double ux = p.x - x_lo, uy = p.y - y_lo;
            if (ux < -0.5 * Lx) ux += Lx; else if (ux >= 0.5 * Lx) ux -= Lx;
            if (uy < -0.5 * Ly) uy += Ly; else if (uy >= 0.5 * Ly) uy -= Ly;
            return make_double2(ux, uy);
        }
        k -= n;
    }
    return make_double2(1e300, 1e300);
}

// Pair energy, cutoff and lane mask without a branch.  For LJ the mask is applied to r^2: a masked
// or out-of-range pair gets r^2 = 1e300, for which 4 r^-6 (r^-6 - 1) underflows to -0.0 — no select
// on the result.  The other potentials select the result.
template <int POT>
__device__ __forceinline__ double pair_e_sel(double dx, double dy, double r2cut, const PairParams &pp, bool valid) {
    const double r2 = fma(dx, dx, dy * dy);
    const bool in = valid && r2 <= r2cut;  // inclusive cutoff, cell_list.cpp:62
    if (POT == POT_LJ) {
        const double r2inv = fast_rcp(in ? r2 : 1e300);
        const double r6inv = r2inv * r2inv * r2inv;
        return 4.0 * r6inv * (r6inv - 1.0);
    } else if (POT == POT_WCA) {
        const double r2inv = fast_rcp(r2);
        const double r6inv = r2inv * r2inv * r2inv;
        const double e = 4.0 * r6inv * (r6inv - 1.0) + 1.0;
        return (in && r2 < 1.2599) ? e : 0.0;
    } else {
        return in ? pair_u<POT>(r2, pp) : 0.0;
    }
}

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, d));
    return v;
}

template <int POT, int G, bool TRACE>
__global__ void __launch_bounds__(256, MC2D_SWEEPG_MIN_BLOCKS) k_sweep_g(SweepArgs a, int nb_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    __shared__ unsigned int cred[8][7];
    constexpr int GPW = 32 / G;
    const Grid &g = a.g;
    const int K = g.K;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gi = lane / G, sl = lane % G;
    const int gid = (blockIdx.x * wpb + wib) * GPW + gi;  // index of this group's active cell
    const size_t per_group = (size_t)(nb_cap + K) * sizeof(double2) + (a.ids ? (size_t)K * sizeof(int) : 0);
    double2 *cand = reinterpret_cast<double2 *>(smem_raw + (size_t)(wib * GPW + gi) * per_group);
    int *own_id = reinterpret_cast<int *>(cand + nb_cap + K);

    const int nay = g.ny >> 1;
    const int nact = (g.ncol >> 1) * nay;
    const bool has_cell = gid < nact;
    const int px = a.colour & 1, py = a.colour >> 1;
    const int acx = has_cell ? gid / nay : 0, acy = has_cell ? gid - acx * nay : 0;
    const int lcx = g.ghost + 2 * acx + px, cy = 2 * acy + py, gcx = g.col0 + 2 * acx + px;
    const int cell = lcx * g.ny + cy;
    const uint32_t cell_gid = (uint32_t)(gcx * g.ny + cy);
    const double x_lo = g.ox + gcx * g.dx, y_lo = g.oy + cy * g.dy;
    const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;

    int n_own = has_cell ? a.cnt[cell] : 0;
    const int n_disp = n_own > 0 ? (a.disp_cell > 0 ? a.disp_cell : a.disp_mult * n_own) : 0;
    const int n_gc = has_cell ? a.n_gc : 0;
    const bool work = (n_disp > 0 || n_gc > 0);

    // ---- stage: 8 neighbour cells -> cand[0..Ms), own cell -> cand[Ms..Ms+n_own) (cell-local) ----
    int M = 0, Ms = 0;
    if (POT != POT_IDEAL && work) {
        int cidx[8], pre[9];
        pre[0] = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            cidx[q] = nbr_cell_index(g, lcx, cy, q);
            pre[q + 1] = pre[q] + a.cnt[cidx[q]];
        }
        M = pre[8];
        if (M <= nb_cap) {
            Ms = M;
            for (int k = sl; k < M; k += G) {
                int cc = cidx[0], cs = 0;
#pragma unroll
                for (int q = 1; q < 8; ++q)
                    if (k >= pre[q]) { cc = cidx[q]; cs = pre[q]; }
                cand[k] = to_local(a.pos[(size_t)cc * K + (k - cs)], x_lo, y_lo, g, hLx, hLy);
            }
        }
    }
    double2 *own = cand + Ms;
    if (work) {
        for (int s = sl; s < n_own; s += G) {
            own[s] = to_local(a.pos[(size_t)cell * K + s], x_lo, y_lo, g, hLx, hLy);
            if (a.ids) own_id[s] = a.ids[(size_t)cell * K + s];
        }
    }
    __syncwarp();
    const int cap_idx = nb_cap + K - 1;  // last valid slot of this group's block (clamp for masked lanes)
    const bool cold = (POT != POT_IDEAL) && __any_sync(FULL_MASK, M > Ms);  // some group reads neighbours from global

    double dU_total = 0.0;
    unsigned int c_att[3] = {0, 0, 0}, c_acc[3] = {0, 0, 0}, c_ovf = 0;
    bool dirty = false;
    Philox4 batch = {0u, 0u, 0u, 0u};
    const int src_base = gi * G;

    // ---- K2: displacement rounds (MC.cpp:96-123); one trial per group per round ----
    {
        const int T = warp_max_int(n_disp);
        const int KM = warp_max_int(Ms + n_own);
        const int KC = cold ? warp_max_int(M - Ms) : 0;
        for (int t = 0; t < T; ++t) {
            if ((t & (G - 1)) == 0)  // sub-lane s draws the words of trial t + s of its cell's stream
                batch = philox4x32_10(cell_gid, MC2D_TAG_DISP | (uint32_t)(t + sl), a.sweep_lo, a.sweep_hi, a.seed_lo,
                                      a.seed_hi);
            const int src = src_base + (t & (G - 1));
            const uint32_t rx = __shfl_sync(FULL_MASK, batch.x, src), ry = __shfl_sync(FULL_MASK, batch.y, src);
            const uint32_t rz = __shfl_sync(FULL_MASK, batch.z, src), rw = __shfl_sync(FULL_MASK, batch.w, src);
            const bool act = t < n_disp;
            const int i = act ? (int)__umulhi(rx, (uint32_t)n_own) : 0;
            const double2 uo = own[i];
            double2 un;
            un.x = uo.x + a.delta * (2.0 * mc2d_u01(ry) - 1.0);
            un.y = uo.y + a.delta * (2.0 * mc2d_u01(rz) - 1.0);
            const bool inside = act && (un.x >= 0.0 && un.x < g.dx && un.y >= 0.0 && un.y < g.dy);
            double dE = 0.0;
            if (POT != POT_IDEAL) {
                double acc = 0.0;
                const int ntot = Ms + n_own, self = Ms + i;
                for (int k = sl; k < KM; k += G) {  // hot loop: shared memory only
                    const bool valid = inside && k < ntot && k != self;
                    const double2 c = cand[min(k, cap_idx)];
                    acc += pair_e_sel<POT>(c.x - un.x, c.y - un.y, a.r2cut, a.pp, valid) -
                           pair_e_sel<POT>(c.x - uo.x, c.y - uo.y, a.r2cut, a.pp, valid);
                }
                if (cold) {
                    for (int k = sl; k < KC; k += G) {
                        const bool valid = inside && k < M - Ms;
                        double2 c = make_double2(0.0, 0.0);
                        if (valid)
                            c = slow_candidate(a.pos, a.cnt, g.ghost, g.nx, g.ny, K, lcx, cy, k, x_lo, y_lo, g.Lx, g.Ly);
                        acc += pair_e_sel<POT>(c.x - un.x, c.y - un.y, a.r2cut, a.pp, valid) -
                               pair_e_sel<POT>(c.x - uo.x, c.y - uo.y, a.r2cut, a.pp, valid);
                    }
                }
                dE = group_sum<G>(acc);
            }
            bool accept = inside && (dE <= 0.0);
            if (__any_sync(FULL_MASK, inside && dE > 0.0))
                accept = accept || (inside && mc2d_u01(rw) < exp(-a.beta * dE));  // MC.cpp:110-115
            if (accept && sl == 0) own[i] = un;
            if (sl == 0) {
                c_att[0] += act;
                c_acc[0] += accept;
                if (accept) dU_total += dE;
            }
            dirty = dirty || accept;
            if (TRACE && sl == 0 && act) {
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
    }

    // ---- K3: grand-canonical rounds (MC.cpp:189-249); insertion and deletion share one probe loop ----
    {
        const int T = warp_max_int(n_gc);
        const int KM0 = warp_max_int(Ms + n_own);
        const int KC = cold ? warp_max_int(M - Ms) : 0;
        for (int t = 0; t < T; ++t) {
            if ((t & (G - 1)) == 0)
                batch = philox4x32_10(cell_gid, MC2D_TAG_GC | (uint32_t)(t + sl), a.sweep_lo, a.sweep_hi, a.seed_lo,
                                      a.seed_hi);
            const int src = src_base + (t & (G - 1));
            const uint32_t rx = __shfl_sync(FULL_MASK, batch.x, src), ry = __shfl_sync(FULL_MASK, batch.y, src);
            const uint32_t rz = __shfl_sync(FULL_MASK, batch.z, src), rw = __shfl_sync(FULL_MASK, batch.w, src);
            const bool act = t < n_gc;
            const double sel = mc2d_u01(rx);
            const int kind = !act ? 0 : (sel < a.p_ins ? 1 : (sel < a.p_ins + a.p_del ? 2 : 0));
            const bool full = (kind == 1) && (n_own >= K);
            const bool evaluate = (kind == 1 && !full) || (kind == 2 && n_own > 0);
            const int i = (kind == 2 && n_own > 0) ? (int)__umulhi(ry, (uint32_t)n_own) : 0;
            double2 pt;
            pt.x = mc2d_u01(ry) * g.dx;
            pt.y = mc2d_u01(rz) * g.dy;
            if (pt.x >= g.dx) pt.x = 0.0;
            if (pt.y >= g.dy) pt.y = 0.0;
            const double2 old = own[i];
            const double2 probe = (kind == 1) ? pt : old;
            double dE = 0.0;
            if (POT != POT_IDEAL) {
                double acc = 0.0;
                const int ntot = Ms + n_own, self = (kind == 2) ? Ms + i : -1;
                const int KM = KM0 + t;  // at most t insertions so far
                for (int k = sl; k < KM; k += G) {
                    const bool valid = evaluate && k < ntot && k != self;
                    const double2 c = cand[min(k, cap_idx)];
                    acc += pair_e_sel<POT>(c.x - probe.x, c.y - probe.y, a.r2cut, a.pp, valid);
                }
                if (cold) {
                    for (int k = sl; k < KC; k += G) {
                        const bool valid = evaluate && k < M - Ms;
                        double2 c = make_double2(0.0, 0.0);
                        if (valid)
                            c = slow_candidate(a.pos, a.cnt, g.ghost, g.nx, g.ny, K, lcx, cy, k, x_lo, y_lo, g.Lx, g.Ly);
                        acc += pair_e_sel<POT>(c.x - probe.x, c.y - probe.y, a.r2cut, a.pp, valid);
                    }
                }
                acc = group_sum<G>(acc);
                dE = (kind == 2) ? -acc : acc;  // MC.cpp:200,225
            }
            // MC.cpp:201 and :226, per-cell form
            const double pref = (kind == 1) ? a.zA / (double)(n_own + 1) : (double)n_own / a.zA;
            const double av = pref * exp(-a.beta * dE);
            const bool accept = evaluate && ((av >= 1.0) || (mc2d_u01(rw) < av));
            if (sl == 0) {
                if (accept && kind == 1) {
                    own[n_own] = pt;
                    if (a.ids) own_id[n_own] = -1;
                }
                if (accept && kind == 2) {
                    own[i] = own[n_own - 1];
                    if (a.ids) own_id[i] = own_id[n_own - 1];
                }
                c_att[1] += (kind == 1);
                c_att[2] += (kind == 2);
                c_acc[1] += (accept && kind == 1);
                c_acc[2] += (accept && kind == 2);
                c_ovf += full;
                if (accept) dU_total += dE;
            }
            if (accept) n_own += (kind == 1) ? 1 : -1;
            dirty = dirty || accept;
            if (TRACE && sl == 0 && kind > 0) {
                unsigned long long slot = atomicAdd(a.trace_n, 1ull);
                if ((long long)slot < a.trace_cap) {
                    mc2d_trial tr;
                    tr.pass = a.pass; tr.cell = (int)cell_gid; tr.seq = n_disp + t; tr.kind = kind;
                    tr.accepted = accept ? 1 : 0; tr.reserved = evaluate ? 1 : 0;
                    const double2 o_ = (kind == 2) ? old : make_double2(0.0, 0.0);
                    const double2 p_ = (kind == 1) ? pt : make_double2(0.0, 0.0);
                    double ox_ = x_lo + o_.x, oy_ = y_lo + o_.y, nx_ = x_lo + p_.x, ny_ = y_lo + p_.y;
                    if (ox_ >= g.Lx) ox_ -= g.Lx; if (oy_ >= g.Ly) oy_ -= g.Ly;
                    if (nx_ >= g.Lx) nx_ -= g.Lx; if (ny_ >= g.Ly) ny_ -= g.Ly;
                    tr.old_x = ox_; tr.old_y = oy_; tr.new_x = nx_; tr.new_y = ny_;
                    tr.dE = evaluate ? dE : nan("");
                    a.trace[slot] = tr;
                }
            }
            __syncwarp();
        }
    }

    // ---- write the cell back (absolute coordinates) ----
    if (dirty) {
        for (int s = sl; s < n_own; s += G) {
            const double2 u = own[s];
            double x = x_lo + u.x, y = y_lo + u.y;
            if (x >= g.Lx) x -= g.Lx; else if (x < 0.0) x += g.Lx;
            if (y >= g.Ly) y -= g.Ly; else if (y < 0.0) y += g.Ly;
            a.pos[(size_t)cell * K + s] = make_double2(x, y);
            if (a.ids) a.ids[(size_t)cell * K + s] = own_id[s];
        }
        if (sl == 0) a.cnt[cell] = n_own;
    }

    // ---- epilogue: fixed-order dU (lane order inside the warp, warp order inside the CTA), counters ----
    __syncwarp();
    const double dU_w = warp_sum(dU_total);
    unsigned int cw[7] = {c_att[0], c_att[1], c_att[2], c_acc[0], c_acc[1], c_acc[2], c_ovf};
#pragma unroll
    for (int q = 0; q < 7; ++q) cw[q] = (unsigned int)warp_sum_int((int)cw[q]);
    if (lane == 0) {
        red[wib] = dU_w;
#pragma unroll
        for (int q = 0; q < 7; ++q) cred[wib][q] = cw[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < wpb; ++i) s += red[i];
        a.partial_dU[blockIdx.x] += s;
    }
    if (threadIdx.x < 7) {
        unsigned long long s = 0;
        for (int i = 0; i < wpb; ++i) s += cred[i][threadIdx.x];
        if (s) atomicAdd(&a.counters[threadIdx.x], s);
    }
}
