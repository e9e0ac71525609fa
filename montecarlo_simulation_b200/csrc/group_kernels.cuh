// group_kernels.cuh — production variants of K4 (energy + virial) and K1' (re-bin) with G lanes per
// cell and 32/G cells per warp (same reasoning as sweep_kernel_p.cuh: a 2.5-sigma LJ cell at
// rho = 0.5 has ~16 half-shell / ~28 full-shell candidates, which does not fill a warp; the per-cell
// fixed cost — nine count loads, prefix, candidate decode — is shared by 32/G cells).
// The warp-per-cell versions in mc2d_kernels.cuh remain for the per-particle parity hooks (PERP).
#pragma once

template <int POT>
__device__ __forceinline__ void pair_uw_fast(double r2, const PairParams &pp, double &u, double &w) {
    if (POT == POT_LJ) {
        const double r2inv = fast_rcp(r2), r6inv = r2inv * r2inv * r2inv;
        u = 4.0 * r6inv * (r6inv - 1.0);
        w = 48.0 * r6inv * (r6inv - 0.5);
    } else if (POT == POT_WCA) {
        const double r2inv = fast_rcp(r2), r6inv = r2inv * r2inv * r2inv;
        const bool in = r2 < 1.2599;
        u = in ? 4.0 * r6inv * (r6inv - 1.0) + 1.0 : 0.0;
        w = in ? 48.0 * r6inv * (r6inv - 0.5) : 0.0;
    } else {
        pair_uw<POT>(r2, pp, u, w);
    }
}

// K4, G lanes per cell.  half = 1: own cell (j > i) + the 4 forward neighbours; half = 0: the
// reference's directed 3x3 double count (caller halves).  Distances use the exact reference
// arithmetic (exact_r2); LJ/WCA take the Newton reciprocal (<= 1 ulp from the IEEE quotient).
template <class V, int POT, int G>
__global__ void __launch_bounds__(256) k_energy_virial_g(V v, PairParams pp, double r2cut, int half, double *partU,
                                                         double *partW, unsigned long long *pair_stats) {
    __shared__ double red[32];
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int gi = lane / G, sl = lane % G;
    const int w = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + gi;
    const double Lx = v_Lx(v), Ly = v_Ly(v), hLx = 0.5 * Lx, hLy = 0.5 * Ly;
    double accU = 0.0, accW = 0.0;
    unsigned int tests = 0, evals = 0;
    int cx = 0, cy = 0, b0 = 0, n0 = 0;
    const bool has = w < v.ncells();
    if (has) {
        v.cell(w, cx, cy);
        v.range(cx, cy, 0, 0, b0, n0);
    }
    int tb[9], ts[10];
    ts[0] = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const int ddx = c / 3 - 1, ddy = c % 3 - 1;  // ddx outer, ddy inner (cell_list.cpp:49-50)
        int b = 0, n = 0;
        bool use = has && n0 > 0 && v.range(cx, cy, ddx, ddy, b, n);
        if (half) use = use && ((ddx == 0 && ddy == 0) || ddx == 1 || (ddx == 0 && ddy == 1));
        tb[c] = b;
        ts[c + 1] = ts[c] + (use ? n : 0);
    }
    const int M = ts[9];
    const int KM = warp_max_int(M), N0 = warp_max_int(n0);
    for (int k = sl; k < KM; k += G) {
        const bool vk = k < M;
        int cb = tb[0], cs = 0;
#pragma unroll
        for (int q = 1; q < 9; ++q)
            if (k >= ts[q]) { cb = tb[q]; cs = ts[q]; }
        const int j = cb + (k - cs);
        const bool same_cell = (cb == b0);
        const int jrel = j - b0;
        double2 pj = make_double2(0.0, 0.0);
        if (vk) pj = v.pos[j];
        for (int i = 0; i < N0; ++i) {
            const bool ok = vk && i < n0 && !(same_cell && (half ? (jrel <= i) : (jrel == i)));
            double2 pi = make_double2(1.0, 1.0);
            if (ok) pi = v.pos[b0 + i];
            const double r2 = exact_r2(pi.x, pi.y, pj.x, pj.y, Lx, Ly, hLx, hLy);
            const bool hit = ok && r2 <= r2cut;  // inclusive, cell_list.cpp:62
            double u, wv;
            pair_uw_fast<POT>(hit ? r2 : 1.0, pp, u, wv);
            accU += hit ? u : 0.0;
            accW += hit ? wv : 0.0;
            tests += ok;
            evals += hit;
        }
    }
    const double sU = block_sum(accU, red);
    const double sW = block_sum(accW, red);
    if (threadIdx.x == 0) {
        partU[blockIdx.x] = sU;
        partW[blockIdx.x] = sW;
    }
    const unsigned int t = (unsigned int)warp_sum_int((int)tests), e = (unsigned int)warp_sum_int((int)evals);
    if (lane == 0 && pair_stats && (t | e)) {
        atomicAdd(&pair_stats[0], (unsigned long long)t);
        atomicAdd(&pair_stats[1], (unsigned long long)e);
    }
}

// K1', G lanes per new cell: gather from the 3x3 old cells by ballot compaction inside the group.
// Same result (cell contents AND slot order) as k_rebin.
//
// The origin moves by (sdx, sdy) = new - old, each inside (-d, d): a new cell overlaps only the old
// cell of the same index and its neighbour on the side the origin moved to, so 4 of the 9 old
// cells can contribute.  The other 5 are skipped (their count is taken as 0) unless the shift along
// that axis is smaller than the rounding slack `eps` of the cell assignment; skipping them does not
// change the slot order because they contribute nothing.
template <int G>
__global__ void __launch_bounds__(256) k_rebin_g(const double2 *pos_o, const int *cnt_o, const int *ids_o,
                                                 double2 *pos_n, int *cnt_n, int *ids_n, Grid gn, double sdx,
                                                 double sdy, int w0, int n0, int w1, int n1, int *err) {
    // this launch covers the owned cells [w0, w0 + n0) and [w1, w1 + n1) (column-major): all of them, or a
    // slab's two boundary columns, or its interior
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int gi = lane / G, sl = lane % G;
    const int wl = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + gi;
    const bool has = wl < n0 + n1;
    const int w = wl < n0 ? w0 + wl : w1 + (wl - n0);
    const int lcx = gn.ghost + (has ? w / gn.ny : 0), cy = has ? w % gn.ny : 0;
    const int gcx = gn.col0 + (lcx - gn.ghost);
    const int K = gn.K;
    const size_t obase = (size_t)(lcx * gn.ny + cy) * K;
    int tc[9], ts[10];
    ts[0] = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const int ddx = c / 3 - 1, ddy = c % 3 - 1;
        const bool dup = (gn.nx == 2 && ddx < 0) || (gn.ny == 2 && ddy < 0);
        const double ex = 1e-9 * gn.dx, ey = 1e-9 * gn.dy;
        const bool reach = (ddx == 0 || (sdx > ex ? ddx > 0 : (sdx < -ex ? ddx < 0 : true)) || gn.nx == 2) &&
                           (ddy == 0 || (sdy > ey ? ddy > 0 : (sdy < -ey ? ddy < 0 : true)) || gn.ny == 2);
        int ocx = lcx + ddx, ocy = cy + ddy;
        if (!gn.ghost) {
            if (ocx < 0) ocx += gn.nx; else if (ocx >= gn.nx) ocx -= gn.nx;
        }
        if (ocy < 0) ocy += gn.ny; else if (ocy >= gn.ny) ocy -= gn.ny;
        tc[c] = ocx * gn.ny + ocy;
        ts[c + 1] = ts[c] + ((has && !dup && reach) ? cnt_o[tc[c]] : 0);
    }
    const int M = ts[9];
    const int KM = warp_max_int(M);
    const unsigned gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (gi * G);
    int nout = 0;
    for (int k0 = 0; k0 < KM; k0 += G) {
        const int k = k0 + sl;
        bool take = false;
        double2 p = make_double2(0.0, 0.0);
        size_t src = 0;
        if (k < M) {
            int cc = tc[0], cs = 0;
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (k >= ts[q]) { cc = tc[q]; cs = ts[q]; }
            src = (size_t)cc * K + (k - cs);
            p = pos_o[src];
            int ngx, ncy;
            sweep_cell_of(gn, p.x, p.y, ngx, ncy);
            take = (ngx == gcx && ncy == cy);
        }
        const unsigned m = __ballot_sync(FULL_MASK, take) & gmask;
        if (take) {
            const int off = nout + __popc(m & ((1u << lane) - 1u));
            if (off < K) {
                pos_n[obase + off] = p;
                if (ids_n) ids_n[obase + off] = ids_o[src];
            }
        }
        nout += __popc(m);
    }
    if (has && sl == 0) {
        cnt_n[lcx * gn.ny + cy] = min(nout, K);
        if (nout > K) atomicMax(&err[0], nout);
    }
}

// K1', common case (rebin4_cells, sweep_kernel_p.cuh): exactly 4 old cells can hold particles of a new cell.
#ifndef MC2D_REBIN_HALO_BLOCKS
#define MC2D_REBIN_HALO_BLOCKS 6
#endif
template <int G, bool HALO>
__global__ void __launch_bounds__(256, HALO ? MC2D_REBIN_HALO_BLOCKS : 6) k_rebin4(const double2 *pos_o, const int *cnt_o, const int *ids_o,
                                                double2 *pos_n, int *cnt_n, int *ids_n, Grid gn, int sx, int sy,
                                                int w0, int n0, int w1, int n1, int *err, SlabHalo sh) {
    constexpr int GPW = 32 / G;
    const int wl0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW;
    rebin4_cells<G, HALO>(pos_o, cnt_o, ids_o, pos_n, cnt_n, ids_n, gn, sx, sy, wl0, w0, n0, w1, n1, err, sh);
}

// ------------------------------------------------------------------------------------------------
// K5 — halo over peer memory (replaces ParticleExchange::refreshGhosts / flushGhostUpdates,
// particle_exchange.cpp:153-258, and the ncclSend/ncclRecv version of this library).
//
// A neighbour's arena is mapped into this process with CUDA IPC.  One launch per sync point:
//   1. push: the live slots and the counts of the boundary column(s) that changed are stored
//      directly into the neighbour's ghost column (NVLink stores, no staging buffer, no pack);
//   2. signal: when the last CTA has finished (ticket counter + system fences) it stores the sync
//      sequence number into the flag word of BOTH neighbours;
//   3. wait: the same CTA spins until the neighbour(s) whose column the NEXT kernel reads have signalled
//      the same sequence number: a pass over even columns reads only the left ghost, a pass over odd
//      columns only the right one; re-bin, energy and download read both.
// The kernels that follow on the stream therefore start after the ghosts they read are complete.  No
// handshake is needed in the opposite direction: a ghost column has one writer (the neighbour, after
// its passes of one x-parity) and one reader (this rank, in its passes of the other x-parity, which
// wait for that writer first), the colour order is the same on every rank, and a re-bin writes the
// other ping-pong buffer.
// ------------------------------------------------------------------------------------------------
struct HaloArgs {
    const double2 *pos;  // local buffer
    const int *cnt;
    const int *ids;
    double2 *dpos[2];  // [0] left neighbour's right-ghost column, [1] right neighbour's left-ghost column
    int *dcnt[2];
    int *dids[2];
    unsigned long long *dflag[2];  // [0] left neighbour's "from right" word, [1] right neighbour's "from left" word
    unsigned long long *myflag;    // [0] from left, [1] from right, [2] CTA ticket counter
    int col[2];                    // local column pushed to the left / right neighbour
    int push[2];
    int signal;                    // store seq into both neighbours' flag words when the pushes are complete
    int wait[2];                   // wait for the left / right neighbour's signal
    int ny, K;
    unsigned long long seq;
    int *err;                           // ctx->d_err: [4] is set when a wait gives up
    unsigned long long spin_budget_ns;  // budget of the wait (mc2d_tuning.spin_timeout_ms)
};

__global__ void __launch_bounds__(256) k_halo_sync(HaloArgs h) {
    const int per_col = h.ny * h.K;
    const int total = (h.push[0] ? per_col : 0) + (h.push[1] ? per_col : 0);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int side = (h.push[0] && idx < per_col) ? 0 : 1;
        const int j = idx - ((side == 1 && h.push[0]) ? per_col : 0);
        const int cell = j / h.K, sl = j - cell * h.K;
        const int n = h.cnt[h.col[side] * h.ny + cell];
        if (sl < n) {
            h.dpos[side][j] = h.pos[(size_t)h.col[side] * per_col + j];
            if (h.ids) h.dids[side][j] = h.ids[(size_t)h.col[side] * per_col + j];
        }
        if (sl == 0) h.dcnt[side][cell] = n;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long ticket = atomicAdd(&h.myflag[2], 1ull);
        if (ticket == gridDim.x - 1) {
            __threadfence_system();
            h.myflag[2] = 0;  // the next launch on this stream starts from zero
            if (h.signal) {
                *(volatile unsigned long long *)h.dflag[0] = h.seq;
                *(volatile unsigned long long *)h.dflag[1] = h.seq;
                __threadfence_system();
            }
            if (h.wait[0]) spin_until_ge(h.myflag + 0, h.seq, h.err, h.spin_budget_ns, 4000);
            if (h.wait[1]) spin_until_ge(h.myflag + 1, h.seq, h.err, h.spin_budget_ns, 4001);
            __threadfence_system();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4 on the slot layout, production variant (computeTotalEnergyCellList + computeTotalVirialCellList,
// thermodynamic_calculator.cpp:194-242; ThermodynamicCalculatorParallel::totalEnergy/totalVirial,
// thermodynamic_calculator_parallel.cpp:56-145): one pass for U and W.
//
// Same decomposition as k_sweep_p: G lanes per cell, 32/G cells per warp, the work items of
// k_order_local (equally full cells share a warp), candidates staged by cp.async into shared memory
// as cell-local coordinates.  Half shell: a cell pairs its own particles (j > i) and those of its 4
// forward neighbours (0,+1), (+1,-1), (+1,0), (+1,+1), so every unordered pair is evaluated once; in a
// slab the last owned column pairs with the right ghost column and the first one leaves its left
// pairs to the left rank.  LJ / WCA accumulate a = sum r^-12 and b = sum r^-6 (one hardware reciprocal + Newton
// steps per pair, rcp_newton1): U = 4 (a - b) [+ 1 per WCA pair], W = 48 a - 24 b.
// Items whose candidates do not fit in the group's shared memory raise err[3]; the host then repeats
// the reduction with k_energy_virial_g.
// ------------------------------------------------------------------------------------------------
template <int POT, int G>
__global__ void __launch_bounds__(256) k_energy_p(const double2 *pos, const int *cnt, const int *perm, Grid g,
                                                  PairParams pp, double r2cut, int nitems, int cap, int stride,
                                                  double *partU, double *partW, unsigned long long *pair_stats,
                                                  int *ovf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GPW = 32 / G;
    const int K = g.K;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gi = lane / G, sl = lane % G;
    double2 *cand = reinterpret_cast<double2 *>(smem_raw + (size_t)(wib * GPW + gi) * stride);
    const int nay = g.ny >> 1, npairs = g.ncol >> 1;
    const int idx = blockIdx.x * wpb + wib;  // heavy-first over (item, colour); also the index of the warp's partial sums
    const int colour = idx & 3, item = idx >> 2;
    const int px = colour & 1, py = colour >> 1;
    const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;

    int acy = -1;
    const int p = item % npairs, j = (item / npairs) * GPW + gi;
    if (item < nitems && j < nay) acy = perm[((size_t)colour * npairs + p) * nay + j];
    const int lcx = g.ghost + 2 * p + px, cy = 2 * max(acy, 0) + py, gcx = g.col0 + 2 * p + px;
    int col3[3], row3[3];
    nbr_rows_cols(g, lcx, cy, col3, row3);
    int n_own = 0, nf[4] = {0, 0, 0, 0}, cf[4];
    cf[0] = col3[1] + row3[2];  // ( 0, +1)
    cf[1] = col3[2] + row3[0];  // (+1, -1)
    cf[2] = col3[2] + row3[1];  // (+1,  0)
    cf[3] = col3[2] + row3[2];  // (+1, +1)
    if (acy >= 0) {
        n_own = cnt[col3[1] + row3[1]];
        if (n_own > 0 && POT != POT_IDEAL) {
#pragma unroll
            for (int q = 0; q < 4; ++q) nf[q] = cnt[cf[q]];
        }
    }
    int Mf = nf[0] + nf[1] + nf[2] + nf[3];
    if (Mf + n_own > cap) {  // does not fit: the host falls back to the unstaged kernel
        if (sl == 0) atomicExch(ovf, 1);
        Mf = 0;
        n_own = 0;
    }
    const bool work = n_own > 0 && POT != POT_IDEAL;
    const double x_lo = g.ox + gcx * g.dx, y_lo = g.oy + cy * g.dy;
    const int ntot = work ? Mf + n_own : 0;
    {
        // candidate k of the group: forward neighbours in the order (0,+1), (+1,-1), (+1,0), (+1,+1), then the own
        // cell; one cp.async per lane and round
        // (slot index of candidate k) = k + o_q for the cell q it lies in: 32-bit offsets, picked by running compares
        const int t1 = nf[0], t2 = t1 + nf[1], t3 = t2 + nf[2];
        const int o0 = cf[0] * K, o1 = cf[1] * K - t1, o2 = cf[2] * K - t2, o3 = cf[3] * K - t3,
                  o4 = (col3[1] + row3[1]) * K - Mf;
        for (int k = sl; k < ntot; k += G) {
            int o = o0;
            o = k >= t1 ? o1 : o;
            o = k >= t2 ? o2 : o;
            o = k >= t3 ? o3 : o;
            o = k >= Mf ? o4 : o;
            cp_async16(&cand[k], &pos[o + k]);
        }
    }
    cp_async_wait_all();
    __syncwarp();
    // to cell-local coordinates, in place; only a cell on the rim of the periodic box (or of the shifted grid) can see a
    // periodic image, everywhere else the conversion is two exact subtractions (as in k_sweep_p)
    const unsigned cand_s = (unsigned)__cvta_generic_to_shared(cand);
    const unsigned cand_e = cand_s + 16u * (unsigned)ntot, nb_e = cand_s + 16u * (unsigned)(work ? Mf : 0);
    {
        const bool rim = work && (gcx == 0 || gcx >= g.nx - 2 || cy == 0 || cy >= g.ny - 2);
        if (__any_sync(FULL_MASK, rim)) {
            for (unsigned q = cand_s + 16u * sl; q < cand_e; q += 16u * G) sts_d2(q, to_local(lds_d2(q), x_lo, y_lo, g, hLx, hLy));
        } else {
            for (unsigned q = cand_s + 16u * sl; q < cand_e; q += 16u * G) {
                const double2 v = lds_d2(q);
                sts_d2(q, make_double2(v.x - x_lo, v.y - y_lo));
            }
        }
    }
    __syncwarp();

    // Per-group loop bounds: the groups of a warp diverge on the trip counts only (equally full cells share a warp,
    // k_order_local), there is no shuffle inside the loops.  Shared-window addresses, one address register per loop.
    double a = 0.0, b = 0.0, accU = 0.0, accW = 0.0;
    unsigned int evals = 0;
    for (int i = 0; i < (work ? n_own : 0); ++i) {
        const unsigned self_s = nb_e + 16u * (unsigned)i;  // own candidates count from j > i
        const double2 ui = lds_d2(self_s);
        for (unsigned q = cand_s + 16u * sl; q < cand_e; q += 16u * G) {
            const double2 c = lds_d2(q);
            const double dx = c.x - ui.x, dy = c.y - ui.y;
            const double r2 = fma(dx, dx, dy * dy);
            bool ok = (q < nb_e || q > self_s) && r2 <= r2cut;  // inclusive cutoff, cell_list.cpp:62
            if (POT == POT_LJ || POT == POT_WCA) {
                if (POT == POT_WCA) ok = ok && r2 < 1.2599;
                const double r2m = __hiloint2double(ok ? __double2hiint(r2) : 0x7E37E43C, __double2loint(r2));
                const double ri = rcp_newton1(r2m);
                const double r6 = ri * ri * ri;
                a = fma(r6, r6, a);
                b += r6;
            } else if (ok) {
                double u, w;
                pair_uw<POT>(r2, pp, u, w);
                accU += u;
                accW += w;
            }
            evals += ok;
        }
    }
    __syncwarp();
    if (POT == POT_LJ || POT == POT_WCA) {
        accU = 4.0 * (a - b) + (POT == POT_WCA ? (double)evals : 0.0);
        accW = 48.0 * a - 24.0 * b;
    }
    // one partial per warp (fixed order inside the warp, fixed-order final reduction): no block barrier
    const double sU = warp_sum(accU), sW = warp_sum(accW);
    if (lane == 0) {
        partU[idx] = sU;
        partW[idx] = sW;
    }
    // distance tests of this cell: n * Mf forward pairs + n (n - 1) / 2 own pairs
    const unsigned int tests_g = (sl == 0 && work) ? (unsigned int)(n_own * Mf + n_own * (n_own - 1) / 2) : 0u;
    const unsigned int t = (unsigned int)warp_sum_int((int)tests_g), e = (unsigned int)warp_sum_int((int)evals);
    if (lane == 0 && pair_stats && (t | e)) {
        atomicAdd(&pair_stats[0], (unsigned long long)t);
        atomicAdd(&pair_stats[1], (unsigned long long)e);
    }
}

// NPT volume move: rescaled copy of the slot layout (SimulationBox::recenter, initial.cpp:27-30: p *= L_new / L_old).
// The cell of a particle does not change (x / dx is invariant), so counts and slot order are copied.
__global__ void __launch_bounds__(256) k_scale_slots(const double2 *pos_o, const int *cnt_o, const int *ids_o,
                                                     double2 *pos_n, int *cnt_n, int *ids_n, int ncell, int K, double rx,
                                                     double ry, double Lx_new, double Ly_new) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)ncell * K) return;
    const int cell = (int)(i / K), s = (int)(i - (long long)cell * K);
    const int n = cnt_o[cell];
    if (s == 0) cnt_n[cell] = n;
    if (s < n) {
        double2 p = pos_o[i];
        p.x *= rx;
        p.y *= ry;
        if (p.x >= Lx_new) p.x = nextafter(Lx_new, 0.0);  // rounding must not push a particle onto the far wall
        if (p.y >= Ly_new) p.y = nextafter(Ly_new, 0.0);
        pos_n[i] = p;
        if (ids_n) ids_n[i] = ids_o[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Single-particle probes on the slot layout (Gibbs-ensemble particle exchange, GibbsMC.cpp:340-394).
// One warp.  mode 0: energy of a test particle at pt (getNeighbors2, cell_list.cpp:72-101); mode 1: the
// index-th particle in cell-major order (offsets = exclusive scan of the owned cells' counts), its energy with
// its neighbours, negated (cell_list.cpp:43-70).  Exact reference distance arithmetic, IEEE division.
// out: [0] dU, [1] x, [2] y; sel: [0] stored cell, [1] slot.
// ------------------------------------------------------------------------------------------------
template <int POT>
__global__ void __launch_bounds__(32) k_probe_slots(const double2 *pos, const int *cnt, Grid g, PairParams pp,
                                                    double r2cut, int mode, double2 pt, long long index,
                                                    const int *offsets, double *out, int *sel) {
    const int lane = threadIdx.x;
    const int ncell = g.ncol * g.ny, cell0 = g.ghost * g.ny;
    int cell = -1, slot = -1;
    if (mode == 1) {
        int lo = 0, hi = ncell - 1;  // last owned cell whose offset is <= index
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)offsets[mid] <= index) lo = mid; else hi = mid - 1;
        }
        cell = cell0 + lo;
        slot = (int)(index - offsets[lo]);
        pt = pos[(size_t)cell * g.K + slot];
    } else {
        int gcx, cy;
        sweep_cell_of(g, pt.x, pt.y, gcx, cy);
        cell = (gcx - g.col0 + g.ghost) * g.ny + cy;
    }
    const int lcx = cell / g.ny, cy = cell - lcx * g.ny;
    const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;
    double acc = 0.0;
    for (int c = 0; c < 9; ++c) {
        const int ddx = c / 3 - 1, ddy = c % 3 - 1;
        if ((g.nx == 2 && ddx < 0) || (g.ny == 2 && ddy < 0)) continue;  // 2 cells across: the same cell twice
        int ocx = lcx + ddx, ocy = cy + ddy;
        if (!g.ghost) {
            if (ocx < 0) ocx += g.nx; else if (ocx >= g.nx) ocx -= g.nx;
        }
        if (ocy < 0) ocy += g.ny; else if (ocy >= g.ny) ocy -= g.ny;
        const int oc = ocx * g.ny + ocy;
        const int n = cnt[oc];
        for (int s = lane; s < n; s += 32) {
            if (mode == 1 && oc == cell && s == slot) continue;
            const double2 q = pos[(size_t)oc * g.K + s];
            const double r2 = exact_r2(pt.x, pt.y, q.x, q.y, g.Lx, g.Ly, hLx, hLy);
            if (r2 <= r2cut) acc += pair_u<POT>(r2, pp);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        out[0] = mode == 1 ? -acc : acc;
        out[1] = pt.x;
        out[2] = pt.y;
        sel[0] = cell;
        sel[1] = slot;
    }
}

// carries out a probed insertion (what = 1) or deletion (what = 2); err[0] = 1 when the cell is full
__global__ void k_apply_probe(double2 *pos, int *cnt, int *ids, int K, int what, const double *probe, const int *sel,
                              double *energy, int *err) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int cell = sel[0];
    const int n = cnt[cell];
    if (what == 1) {
        if (n >= K) {
            err[0] = 1;
            return;
        }
        pos[(size_t)cell * K + n] = make_double2(probe[1], probe[2]);
        if (ids) ids[(size_t)cell * K + n] = -1;
        cnt[cell] = n + 1;
    } else {
        const int slot = sel[1];
        pos[(size_t)cell * K + slot] = pos[(size_t)cell * K + n - 1];
        if (ids) ids[(size_t)cell * K + slot] = ids[(size_t)cell * K + n - 1];
        cnt[cell] = n - 1;
    }
    energy[0] += probe[0];
}

// fp64 roofline probe: 8 independent DFMA chains per thread, enough warps to fill every SM
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}
