// sweep_kernel_p.cuh — K2 + K3, production variant: persistent warps over a sorted work list.
//
// Same trial algorithm and Philox streams (keyed by global cell id and trial index) as k_sweep
// (sweep_kernel.cuh, one warp per cell): the two give the same configurations (tested).  What
// changes is the work decomposition — G lanes per cell, 32/G cells per warp — and the scheduling
// (profiles/r01_notes.md):
//
//  * k_order_local sorts, per colour and per pair of cell columns, the active cells by falling
//    (occupancy, candidate count).  A work item is 32/G consecutive cells of that order, so the
//    32/G groups of a warp run the same number of displacement rounds and candidate iterations
//    (in grid order a warp ran max-of-8 Poisson rounds: half of the lanes idle).
//  * Items are handed out through an atomic ticket counter to 148 x 3 resident CTAs: no CTA-launch /
//    CTA-tail cost per item, and the warps of an SM drift out of phase, so the global-memory latency
//    of one warp's staging hides behind the arithmetic of the others.
//  * One launch does the four colour passes of a sweep.  Single rank: the items of all colours are ONE
//    list (colour-major, then column pair, heavy-first inside a pair); an item of colour p+1 in
//    column pair P waits — between items, never inside one — until the items of colour p in the
//    pairs P-1, P, P+1 are finished (per-pair completion counters).  No grid barrier:
//    the tail of a colour overlaps the head of the next.  Slab / single-pass launches: items of one
//    colour, heavy-first over all pairs, grid barrier between colours.
//  * The index chain of an item (work ticket -> sorted cell -> 9 cell counts) is prefetched into
//    registers while the previous items run: staging starts with the position loads.
//
// Slab mode (halo_fused): the halo exchange is part of this kernel.  The items of the boundary column pair of a
// pass come first in the item list; their warps wait for the neighbour's flag word (the ghost column they read is
// complete), read that column past the L1, and after the write-back store their cells straight into the
// neighbour's ghost column (peer memory over NVLink).  The warp that finishes the last boundary item of the pass
// fences and stores the next sync sequence number into both neighbours' flag words.  The interior items never
// touch a ghost, so the NVLink traffic and the neighbour's latency hide behind them; the grid barrier between
// colours is local.  (Protocol and hazards: k_halo_sync in group_kernels.cuh.)
//
// The energy change of a pass is accumulated per ITEM (fixed composition, fixed order inside the
// warp), so the running energy stays reproducible although items are scheduled dynamically.
#pragma once
#include <cooperative_groups.h>

// ------------------------------------------------------------------------------------------------
// k_order_local: grid (column pairs, 4 colours), 128 threads.  perm[colour][pair][j] = row index acy
// of the j-th cell of that colour in that column pair, by falling (n_own, n_own + neighbours).
// Stable counting sort: every rank depends only on the counts, never on timing.
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Bounded waits.  Every in-kernel wait on a word that another rank (or another warp) writes has a time
// budget (mc2d_tuning.spin_timeout_ms): a neighbour that failed before its launch, returned early or
// died must not hang this GPU inside a cooperative kernel.  When the budget runs out the waiter sets
// err[4]; every other wait sees that word and gives up at once, the kernel drains (on garbage), and
// the host returns MC2D_ERR_PEER_TIMEOUT and drops the configuration.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mc2d_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one look inside a spin loop: true when the wait must be abandoned (t0 = 0 on the first call)
// (code, detail: which wait gave up first — err[6], err[7], reported with the failure)
__device__ __forceinline__ bool spin_expired(int *err, unsigned long long budget_ns, unsigned long long &t0, int code = 0,
                                             int detail = 0) {
    if (*(volatile int *)&err[4]) return true;
    const unsigned long long now = mc2d_globaltimer();
    if (t0 == 0) t0 = now;
    if (now - t0 > budget_ns) {
        if (atomicExch(&err[4], 1) == 0) {
            err[6] = code;
            err[7] = detail;
        }
        return true;
    }
    return false;
}
__device__ __forceinline__ bool spin_until_ge(const unsigned long long *flag, unsigned long long want, int *err,
                                              unsigned long long budget_ns, int code = 0) {
    const volatile unsigned long long *f = flag;
    unsigned long long t0 = 0;
    while (*f < want) {
        if (spin_expired(err, budget_ns, t0, code, (int)(want - *f))) return false;
        __nanosleep(32);
    }
    return true;
}

__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, d));
    return v;
}

// Slab mode, halo inside the re-bin and work-order kernels (no second stream, no wait / push launches of their own):
//  * k_rebin4: a warp that re-bins cells of the first or last owned column first waits until both neighbours have
//    completed the ghost columns of the OLD buffer (flag words >= wait_seq), stores every particle it places also
//    into the neighbour's ghost column of the NEW buffer, and the last such warp to finish signals signal_seq.
//  * k_order_local: the CTAs of the slab-edge column pairs wait for that signal from the neighbour (they read the
//    new ghost counts).
struct SlabHalo {
    int on;
    const unsigned long long *myflag;  // [0] from left, [1] from right
    unsigned long long wait_seq, signal_seq;
    double2 *hpos[2];                  // [0] left neighbour's right-ghost column, [1] right neighbour's left-ghost column
    int *hcnt[2];
    int *hids[2];
    unsigned long long *hflag[2];
    int *bcount;                       // finished boundary warps of this launch
    int nbwarps;                       // ... out of this many
    int *err;
    unsigned long long spin_budget_ns;
};

#ifndef MC2D_ORDL_SHIFT
#define MC2D_ORDL_SHIFT 3
#endif
constexpr int ORDL_NB = 16;  // occupancy bins (15 and above share the first)
constexpr int ORDL_MB = 64 >> MC2D_ORDL_SHIFT;  // candidate-count bins of width 2^MC2D_ORDL_SHIFT (the last ones share the first)
constexpr int ORDL_BINS = ORDL_NB * ORDL_MB;
constexpr int ORDL_WARPS = 4;

// One (column pair p, colour) task by the NW warps of a CTA (all of its threads call this).
// ol_smem: 3 columns x ny counts, then nay packed (bin << 16 | rank in warp); wc: NW x ORDL_BINS; bin_base: ORDL_BINS.
template <int NW>
__device__ __forceinline__ void order_local_task(const Grid &g, const int *cnt, int *perm, int p, int colour,
                                                 int *ol_smem, int *wc, int *bin_base, const SlabHalo &sh) {
    constexpr int NT = NW * 32;
    const int px = colour & 1, py = colour >> 1;
    const int nay = g.ny >> 1, npairs = g.ncol >> 1;
    const int lcx = g.ghost + 2 * p + px;
    if (sh.on && ((p == 0 && px == 0) || (p == npairs - 1 && px == 1))) {  // reads a ghost column's counts
        if (threadIdx.x == 0) {
            spin_until_ge(sh.myflag + px, sh.signal_seq, sh.err, sh.spin_budget_ns, 5000 + px);
            __threadfence();
        }
        __syncthreads();
    }
    int *ol_cnt = ol_smem, *ol_key = ol_smem + 3 * g.ny;
    for (int c = 0; c < 3; ++c) {
        int col = lcx + c - 1;
        if (!g.ghost) {
            if (col < 0) col += g.nx; else if (col >= g.nx) col -= g.nx;
        }
        for (int cy = threadIdx.x; cy < g.ny; cy += NT) ol_cnt[c * g.ny + cy] = __ldcg(&cnt[col * g.ny + cy]);
    }
    for (int i = threadIdx.x; i < NW * ORDL_BINS; i += NT) wc[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int per_warp = (nay + NW - 1) / NW;  // each warp owns a contiguous range of rows (stable order)
    for (int k0 = 0; k0 < per_warp; k0 += 32) {
        const int k = k0 + lane;
        const int acy = wib * per_warp + k;
        const bool has = k < per_warp && acy < nay;
        int b = ORDL_BINS;  // lanes without a cell match among themselves and are ignored
        if (has) {
            const int cy = 2 * acy + py;
            const int ym = (cy == 0) ? g.ny - 1 : cy - 1, yp = (cy == g.ny - 1) ? 0 : cy + 1;
            const int n = ol_cnt[g.ny + cy];
            const int m = ol_cnt[ym] + ol_cnt[cy] + ol_cnt[yp] + ol_cnt[g.ny + ym] + ol_cnt[g.ny + yp] +
                          ol_cnt[2 * g.ny + ym] + ol_cnt[2 * g.ny + cy] + ol_cnt[2 * g.ny + yp] + n;
            b = (ORDL_NB - 1 - min(n, ORDL_NB - 1)) * ORDL_MB + (ORDL_MB - 1 - min(m >> MC2D_ORDL_SHIFT, ORDL_MB - 1));
        }
        const unsigned peers = __match_any_sync(FULL_MASK, b);
        const int r = __popc(peers & lt);
        int prev = 0;
        if (has) prev = wc[wib * ORDL_BINS + b];
        __syncwarp();
        if (has && r == 0) wc[wib * ORDL_BINS + b] = prev + __popc(peers);
        __syncwarp();
        if (has) ol_key[acy] = (b << 16) | (prev + r);
    }
    __syncthreads();
    // per bin: exclusive prefix over the warps, bin total -> bin_base
    for (int b = threadIdx.x; b < ORDL_BINS; b += NT) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int v = wc[w * ORDL_BINS + b];
            wc[w * ORDL_BINS + b] = run;
            run += v;
        }
        bin_base[b] = run;
    }
    __syncthreads();
    if (wib == 0) {  // exclusive scan of the bin totals by one warp
        constexpr int PER = ORDL_BINS / 32;
        int s = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) s += bin_base[lane * PER + i];
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        int run = incl - s;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int v = bin_base[lane * PER + i];
            bin_base[lane * PER + i] = run;
            run += v;
        }
    }
    __syncthreads();
    int *out = perm + ((size_t)colour * npairs + p) * nay;
    for (int k0 = 0; k0 < per_warp; k0 += 32) {
        const int k = k0 + lane;
        const int acy = wib * per_warp + k;
        if (k < per_warp && acy < nay) {
            const int key = ol_key[acy], b = key >> 16;
            out[bin_base[b] + wc[wib * ORDL_BINS + b] + (key & 0xffff)] = acy;
        }
    }
    __syncthreads();  // the next task of this CTA reuses the scratch
}

// dynamic shared memory: 3 columns x ny counts, then nay keys
// zero_work / zero_done (optional): the sweep kernel's ticket counters and completion counters, which the launch that
// follows this one on the stream expects to be zero — cleared here instead of by two memsets between the kernels.
__global__ void __launch_bounds__(ORDL_WARPS * 32) k_order_local(Grid g, const int *cnt, int *perm, SlabHalo sh,
                                                                 int *zero_work, int n_work, int *zero_done, int n_done) {
    extern __shared__ int ol_smem[];
    __shared__ int wc[ORDL_WARPS * ORDL_BINS];
    __shared__ int bin_base[ORDL_BINS];
    {
        const int cta = blockIdx.y * gridDim.x + blockIdx.x, nthreads = gridDim.x * gridDim.y * blockDim.x;
        for (int i = cta * blockDim.x + threadIdx.x; i < n_done; i += nthreads) zero_done[i] = 0;
        if (cta == 0 && (int)threadIdx.x < n_work) zero_work[threadIdx.x] = 0;
    }
    order_local_task<ORDL_WARPS>(g, cnt, perm, blockIdx.x, blockIdx.y, ol_smem, wc, bin_base, sh);
}

// ------------------------------------------------------------------------------------------------
// K1', common case (body of k_rebin4, group_kernels.cuh): both components of
// the origin shift exceed the rounding slack and the grid is at least 3 cells wide, so exactly the 4 old cells
// {0, sx} x {0, sy} (sx, sy = sign of the shift) can hold particles of a new cell.  They are scanned in the order
// k_rebin_g scans them (dx outer, dy inner, ascending): same slots.  A warp handles 32/G consecutive owned cells
// starting at wl0 (all 32 lanes call this).
// ------------------------------------------------------------------------------------------------
template <int G, bool HALO>
__device__ __forceinline__ void rebin4_cells(const double2 *__restrict__ pos_o, const int *__restrict__ cnt_o,
                                             const int *__restrict__ ids_o, double2 *pos_n, int *cnt_n, int *ids_n,
                                             const Grid &gn, int sx, int sy, int wl0, int w0, int n0, int w1, int n1,
                                             int *err, const SlabHalo &sh) {
    const int lane = threadIdx.x & 31;
    const int gi = lane / G, sl = lane % G;
    const int wl = wl0 + gi;
    const bool has = wl < n0 + n1;
    const int w = wl < n0 ? w0 + wl : w1 + (wl - n0);
    const int lcx = gn.ghost + (has ? w / gn.ny : 0), cy = has ? w % gn.ny : 0;
    const int gcx = gn.col0 + (lcx - gn.ghost);
    const int K = gn.K;
    const size_t obase = (size_t)(lcx * gn.ny + cy) * K;
    // slab edge: 0 = first owned column (goes to the left neighbour), 1 = last owned column, -1 = interior
    const int hs = (HALO && has) ? (lcx == gn.ghost ? 0 : (lcx == gn.ghost + gn.ncol - 1 ? 1 : -1)) : -1;
    const bool bwarp = HALO && __any_sync(FULL_MASK, hs >= 0);
    if (bwarp) {  // the old ghost columns must be complete before this warp reads them
        if (lane == 0) {
            spin_until_ge(sh.myflag + 0, sh.wait_seq, sh.err, sh.spin_budget_ns, 6000);
            spin_until_ge(sh.myflag + 1, sh.wait_seq, sh.err, sh.spin_budget_ns, 6001);
            __threadfence();
        }
        __syncwarp();
    }
    int tc[4], ts[5];
    ts[0] = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int ddx = (c >> 1) ? max(sx, 0) : min(sx, 0), ddy = (c & 1) ? max(sy, 0) : min(sy, 0);
        int ocx = lcx + ddx, ocy = cy + ddy;
        if (!gn.ghost) {
            if (ocx < 0) ocx += gn.nx; else if (ocx >= gn.nx) ocx -= gn.nx;
        }
        if (ocy < 0) ocy += gn.ny; else if (ocy >= gn.ny) ocy -= gn.ny;
        tc[c] = ocx * gn.ny + ocy;
        // (explicit ld.global.nc for the others: with an ld.global.cg on the same pointer in sight the compiler no
        // longer proves the buffer read-only, and plain loads made the whole kernel 10 us slower)
        ts[c + 1] = ts[c] + (has ? (!HALO ? cnt_o[tc[c]] : (bwarp ? __ldcg(&cnt_o[tc[c]]) : __ldg(&cnt_o[tc[c]]))) : 0);
    }
    const int M = ts[4];
    const int KM = warp_max_int(M);
    const unsigned gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (gi * G);
    const unsigned lt = (1u << lane) - 1u;
    int nout = 0;
    // One round = G candidates per new cell: fetch (index -> old slot -> position), then test / vote / place.  TWO rounds
    // are fetched before the first is placed, so two loads per lane are in flight (the compiler interleaves like this
    // by itself only for the plain-load instantiation, where it is worth 4 us of 35; with the .cg / .nc intrinsics and
    // the peer stores in the loop it has to be written out).
    auto fetch = [&](int k, size_t &src, double2 &p) {
        if (k >= M) return false;
        const int q = (k >= ts[1]) + (k >= ts[2]) + (k >= ts[3]);
        const int cc = q == 0 ? tc[0] : (q == 1 ? tc[1] : (q == 2 ? tc[2] : tc[3]));
        const int cs = q == 0 ? 0 : (q == 1 ? ts[1] : (q == 2 ? ts[2] : ts[3]));
        src = (size_t)cc * K + (k - cs);
        // (a ghost cell: written by the neighbour, read past the L1)
        p = !HALO ? pos_o[src] : (bwarp ? __ldcg(&pos_o[src]) : __ldg(&pos_o[src]));
        return true;
    };
    auto place = [&](bool valid, size_t src, double2 p) {
        bool take = false;
        if (valid) {
            int ngx, ncy;
            sweep_cell_of(gn, p.x, p.y, ngx, ncy);
            take = (ngx == gcx && ncy == cy);
        }
        const unsigned m = __ballot_sync(FULL_MASK, take) & gmask;
        if (take) {
            const int off = nout + __popc(m & lt);
            if (off < K) {
                pos_n[obase + off] = p;
                const int id = ids_n ? (!HALO ? ids_o[src] : (bwarp ? __ldcg(&ids_o[src]) : __ldg(&ids_o[src]))) : 0;
                if (ids_n) ids_n[obase + off] = id;
                if (hs >= 0) {  // ... and into the neighbour's ghost column of the new buffer
                    // (no dynamic index into the kernel parameters: that would copy the struct to local memory)
                    double2 *hp = hs == 0 ? sh.hpos[0] : sh.hpos[1];
                    hp[(size_t)cy * K + off] = p;
                    if (ids_n) {
                        int *hi = hs == 0 ? sh.hids[0] : sh.hids[1];
                        hi[(size_t)cy * K + off] = id;
                    }
                }
            }
        }
        nout += __popc(m);
    };
    int k0 = 0;
    for (; k0 + G < KM; k0 += 2 * G) {
        size_t src_a = 0, src_b = 0;
        double2 pa = make_double2(0.0, 0.0), pb = pa;
        const bool va = fetch(k0 + sl, src_a, pa), vb = fetch(k0 + G + sl, src_b, pb);
        place(va, src_a, pa);
        place(vb, src_b, pb);
    }
    if (k0 < KM) {
        size_t src_a = 0;
        double2 pa = make_double2(0.0, 0.0);
        const bool va = fetch(k0 + sl, src_a, pa);
        place(va, src_a, pa);
    }
    if (has && sl == 0) {
        cnt_n[lcx * gn.ny + cy] = min(nout, K);
        if (hs >= 0) {
            int *hc = hs == 0 ? sh.hcnt[0] : sh.hcnt[1];
            hc[cy] = min(nout, K);
        }
        if (nout > K) atomicMax(&err[0], nout);
    }
    if (bwarp) {  // one more boundary warp done; the last one signals both neighbours
        __threadfence_system();
        __syncwarp();
        if (lane == 0 && atomicAdd(sh.bcount, 1) == sh.nbwarps - 1) {
            __threadfence_system();
            *(volatile unsigned long long *)sh.hflag[0] = sh.signal_seq;
            *(volatile unsigned long long *)sh.hflag[1] = sh.signal_seq;
            __threadfence_system();
        }
    }
}

// absolute position -> coordinates relative to the cell origin (x_lo, y_lo), nearest periodic image
__device__ __forceinline__ double2 to_local(double2 p, double x_lo, double y_lo, const Grid &g, double hLx,
                                            double hLy) {
    double ux = p.x - x_lo, uy = p.y - y_lo;
    if (ux < -hLx) ux += g.Lx; else if (ux >= hLx) ux -= g.Lx;
    if (uy < -hLy) uy += g.Ly; else if (uy >= hLy) uy -= g.Ly;
    return make_double2(ux, uy);
}

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}

// (unrolling the candidate loops does not pay: 49.2 us per colour pass of config #4 without, 50.3 us x2, 54.0 us x4)
#ifndef MC2D_SHARE_GC_BATCH
#define MC2D_SHARE_GC_BATCH 1
#endif
#ifndef MC2D_HOT_UNROLL
#define MC2D_HOT_UNROLL 1
#endif
#define MC2D_PRAGMA_(x) _Pragma(#x)
#define MC2D_UNROLL(n) MC2D_PRAGMA_(unroll n)
#ifndef MC2D_SWEEPP_THREADS
#define MC2D_SWEEPP_THREADS 256
#endif
#ifndef MC2D_SWEEPP_MIN_BLOCKS
#define MC2D_SWEEPP_MIN_BLOCKS 3
#endif

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// L2-only variant: for cells a neighbour GPU writes while this kernel runs (the L1 of this SM may hold an older copy)
__device__ __forceinline__ void cp_async16_cg(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// 1/x from the hardware seed (rcp.approx.ftz.f64, ~20 good bits) and MC2D_SWEEPP_NEWTON Newton steps.  One step
// leaves ~1e-12 per reciprocal (worst per-trial dE error measured against the oracle: 7e-12 of max(1, |dE|),
// tools/check_dE_accuracy.py); two steps reach the last bit (worst 1e-15) and cost two more DFMA per pair.
// The parity hooks K4/K6 keep the IEEE division of the reference.
// MC2D_SWEEPP_NEWTON = 3 (default): ONE second-order step, r = r0 (1 + e + e^2) with e = 1 - x r0: the error is e^3
// (~2^-60 for the ~20-bit seed), the same last-bit accuracy as two Newton steps for three DFMA instead of four.
#ifndef MC2D_SWEEPP_NEWTON
#define MC2D_SWEEPP_NEWTON 3
#endif
__device__ __forceinline__ double rcp_newton1(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if MC2D_SWEEPP_NEWTON == 3
    const double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
#else
#pragma unroll
    for (int it = 0; it < MC2D_SWEEPP_NEWTON; ++it) r = fma(r, fma(-x, r, 1.0), r);
#endif
    return r;
}

// shared-memory accesses by 32-bit shared-window address: the hot loops keep ONE address register per array instead of
// re-deriving generic pointers from the kernel arguments (what the compiler did under register pressure)
__device__ __forceinline__ double2 lds_d2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_d2(unsigned addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ unsigned pin_u32(unsigned v) {  // opaque to the optimiser: the value stays in a register
    asm volatile("" : "+r"(v));
    return v;
}

// Empty candidate slots hold this point: its distance to anything in a cell is ~1e150, beyond every
// cutoff, so the hot loop needs no bounds test.
#define MC2D_FAR 1e150

// One pair term t with u(r) = pair_scale<POT>() * t, zero beyond the inclusive cutoff (cell_list.cpp:62)
// and for the excluded (self) slot.  LJ: t = r^-12 - r^-6 with the mask applied to the HIGH WORD of r^2
// (masked pairs get r^2 ~ 1e300, whose r^-6 underflows to 0): one predicate, one select, no branch.
template <int POT>
__device__ __forceinline__ double pair_term(double dx, double dy, double r2cut, const PairParams &pp, bool notself) {
    const double r2 = fma(dx, dx, dy * dy);
    if (POT == POT_LJ || POT == POT_WCA) {
        bool ok = notself && r2 <= r2cut;
        if (POT == POT_WCA) ok = ok && r2 < 1.2599;  // potential.cpp:62
        const double r2m = __hiloint2double(ok ? __double2hiint(r2) : 0x7E37E43C, __double2loint(r2));
        const double ri = rcp_newton1(r2m);
        const double r6 = ri * ri * ri;
        const double t = fma(r6, r6, -r6);
        return (POT == POT_WCA) ? t + (ok ? 0.25 : 0.0) : t;
    } else {
        return (notself && r2 <= r2cut) ? pair_u<POT>(r2, pp) : 0.0;
    }
}
template <int POT>
__device__ __forceinline__ double pair_scale() {
    return (POT == POT_LJ || POT == POT_WCA) ? 4.0 : 1.0;
}

// Everything another SM (or the neighbour GPU) may have written while this kernel runs — positions, counts, ids — is read
// PAST THE L1 (cp.async.cg, ld.global.cg: the L2 is the point of coherence), and the completion counters that order those
// reads are read with ld.relaxed.gpu: the data loads are issued only after the branch on the counter's value, so they
// see at least what the writer had made visible (__threadfence) before it counted.  The alternative, ld.acquire.gpu
// (= the same load + CCTL.IVALL, the SM's whole L1 dropped at every look) with L1-cached data loads, cost 49.2 against
// 47.7 us per colour pass; -DMC2D_SWEEPP_IVALL builds it for A-B.
#ifdef MC2D_SWEEPP_IVALL
#define MC2D_LD_DONE "ld.acquire.gpu.global.s32"
#else
#define MC2D_LD_DONE "ld.relaxed.gpu.global.s32"
#endif

struct ItemMeta {
    int acy;              // row index of the cell inside its colour (from perm), -1: no cell
    int n_own;
    uint32_t c_lo, c_hi;  // the 8 neighbour counts, one byte each (K <= 255)
};

// stored index of the 3x3 neighbourhood of stored cell (lcx, cy): col3[0..2] + row3[0..2]
__device__ __forceinline__ void nbr_rows_cols(const Grid &g, int lcx, int cy, int col3[3], int row3[3]) {
    int cm = lcx - 1, cp = lcx + 1;
    if (!g.ghost) {
        if (cm < 0) cm += g.nx;
        if (cp >= g.nx) cp -= g.nx;
    }
    col3[0] = cm * g.ny; col3[1] = lcx * g.ny; col3[2] = cp * g.ny;
    row3[0] = (cy == 0) ? g.ny - 1 : cy - 1; row3[1] = cy; row3[2] = (cy == g.ny - 1) ? 0 : cy + 1;
}

template <int POT, int G, bool TRACE>
__global__ void __launch_bounds__(MC2D_SWEEPP_THREADS, MC2D_SWEEPP_MIN_BLOCKS) k_sweep_p(SweepArgs a, int R) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int cred[8][7];
    __shared__ double lntab[258];  // ln k for the insertion / deletion prefactors (K <= 255)
    constexpr int GPW = 32 / G;
    const Grid &g = a.g;
    const int K = g.K;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gi = lane / G, sl = lane % G;
    if (a.n_gc > 0) {
        for (int k = threadIdx.x; k < 258; k += blockDim.x) lntab[k] = log((double)max(k, 1));
        __syncthreads();
    }
    const double ln_zA = log(a.zA);
    // Shared memory: every warp owns R candidate slots (+ R ints when ids are carried).  The groups of an
    // item take what they need from them by a prefix sum (neighbours + own + room for insertions), so the
    // mean need, not the worst case, sets how many cells are in flight per SM.  An item that does not fit
    // is run in batches of consecutive groups; the host guarantees that one cell always fits (R >= 9 K).
    double2 *wslots = reinterpret_cast<double2 *>(smem_raw + (size_t)wib * R * (sizeof(double2) + (a.ids ? sizeof(int) : 0)));
    int *wids = reinterpret_cast<int *>(wslots + R);

    const int nay = g.ny >> 1, npairs = g.ncol >> 1;
    // per colour pass (a launch does one pass, or all four of a sweep with a grid barrier in between: a launch and
    // its ramp-up cost ~8 us, a sixth of a pass of config #4)
    int px = 0, py = 0, pass_no = a.pass;
    const int *perm = a.order;
    int *work = a.work;
    double *pdU = a.partial_dU;
    const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;
    const int src_base = gi * G;
    const double escale = pair_scale<POT>();

    unsigned int c_att[3] = {0, 0, 0}, c_acc[3] = {0, 0, 0}, c_ovf = 0;
#ifdef MC2D_PROFILE_WAIT  // -DMC2D_PROFILE_WAIT: a few warps print where their cycles went (profiles/r01_notes.md)
    long long prof_wait = 0, prof_items = 0, prof_bar = 0;
    const long long prof_t0 = clock64();
#endif

    // ---- work distribution and prefetch ----
    // Warp w starts with item w of the heavy-first list (a fixed first item: thousands of warps asking one address
    // for a ticket at the same instant would queue for tens of microseconds); every further item comes from the
    // ticket counter, asked for only when the previous item starts (late binding: a warp that drew a heavy item
    // simply takes fewer items, the tail is one light item long).  The index chain of the next item is fetched in
    // the shadow of the current one: ticket at the top, sorted cell after the staging wait, the 9 counts after
    // the displacement rounds.
    const int W = gridDim.x * wpb;
    auto ticket = [&]() {
        int v = 0;
        if (lane == 0) v = W + atomicAdd(work, 1);
        return v;
    };
    const int ipp = (nay + GPW - 1) / GPW;  // items per column pair
    int bpair = 0;                          // slab mode: the boundary pair of this pass, its items are 0 .. ipp-1
    // item -> column pair, rank inside the pair; the division is a multiplication by the reciprocal + one fix-up
    const unsigned div_n = (unsigned)(a.halo_fused ? npairs - 1 : a.pair_n), div_inv = 0xFFFFFFFFu / div_n;
    auto divmod = [&](unsigned x, unsigned &q, unsigned &rem) {
        q = __umulhi(x, div_inv);  // floor(x / n) or one less
        rem = x - q * div_n;
        if (rem >= div_n) {
            ++q;
            rem -= div_n;
        }
    };
    const unsigned div_i = (unsigned)ipp, div_i_inv = 0xFFFFFFFFu / div_i;
    auto colour_of = [&](int psv) { return a.npass > 1 ? a.colours[psv] : a.colour; };
    // pass of an item: the loop variable, or (pipeline) the colour-major position in the list of all four colours
    auto pass_of = [&](int item, int ps_loop) {
        return a.pipeline ? (int)(item >= a.nitems) + (int)(item >= 2 * a.nitems) + (int)(item >= 3 * a.nitems) : ps_loop;
    };
    auto split = [&](int item, int psv, int &p, int &r) {
        unsigned q, rem;
        if (a.pipeline) {  // pair-major inside a colour: early pairs finish early and release the next colour
            const unsigned idx = (unsigned)(item - psv * a.nitems);
            q = __umulhi(idx, div_i_inv);
            rem = idx - q * div_i;
            if (rem >= div_i) {
                ++q;
                rem -= div_i;
            }
            p = (int)q;
            r = (int)rem;
            // slab: the slab-edge pair of the pass goes first (pair 0 is there already; for odd active columns the
            // last pair is rotated to the front).  Its items then finish about one pass before the neighbour's next
            // pass asks for them instead of just in time; their own local dependencies (the last pairs of the
            // previous pass) hold them up for one item at most, because tickets are handed out in order.
            if (a.halo_fused && (colour_of(psv) & 1)) p = (q == 0) ? npairs - 1 : (int)q - 1;
        } else if (a.halo_fused) {
            if (item < ipp) {
                p = bpair;
                r = item;
            } else {
                divmod((unsigned)(item - ipp), q, rem);
                r = (int)q;
                p = (int)rem + ((int)rem >= bpair ? 1 : 0);
            }
        } else {
            divmod((unsigned)item, q, rem);
            p = a.pair_lo + (int)rem;
            r = (int)q;
        }
    };
    const int n_all = a.pipeline ? a.npass * a.nitems : a.nitems;
    // -> acy of this group's cell (or -1); the item's pass, pair and rank
    auto load_perm = [&](int item, int ps_loop, int &psv, int &p, int &r) {
        psv = item < n_all ? pass_of(item, ps_loop) : ps_loop;
        if (item >= n_all) {
            p = 0;
            r = 0;
            return -1;
        }
        split(item, psv, p, r);
        const int j = r * GPW + gi;
        const int *pm = a.order + (size_t)colour_of(psv) * npairs * nay;
        return (j < nay) ? pm[(size_t)p * nay + j] : -1;
    };
    auto load_counts = [&](int psv, int p, int acy, bool past_l1) {
#ifndef MC2D_SWEEPP_IVALL
        past_l1 = true;
#endif
        ItemMeta m;
        m.acy = acy;
        m.n_own = 0;
        m.c_lo = m.c_hi = 0;
        if (acy >= 0) {
            const int c = colour_of(psv);
            int col3[3], row3[3];
            nbr_rows_cols(g, g.ghost + 2 * p + (c & 1), 2 * acy + (c >> 1), col3, row3);
            m.n_own = past_l1 ? __ldcg(&a.cnt[col3[1] + row3[1]]) : a.cnt[col3[1] + row3[1]];
            if (POT != POT_IDEAL) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int o = q + (q >= 4);
                    const int *pc = &a.cnt[col3[o / 3] + row3[o % 3]];
                    const uint32_t n = (uint32_t)(past_l1 ? __ldcg(pc) : *pc);
                    if (q < 4) m.c_lo |= n << (8 * q);
                    else m.c_hi |= n << (8 * (q - 4));
                }
            }
        }
        return m;
    };
    // pipeline: an item of colour pass psv > 0 in column pair p reads cells that the previous colour pass wrote in the
    // pairs p-1, p, p+1: wait until all their items are finished (the cells are then read past the L1, see MC2D_LD_DONE).
    // block = false: one look (used while the current item is still open: it may itself be one of the items the next
    // one waits for, so blocking there would deadlock); block = true: spin (between items only).
    // Slab (pipeline + halo_fused): the column pairs do not wrap — beyond the slab edge lies the neighbour rank.  The item
    // of a pass's slab-edge pair (pair 0 when even columns are active, the last pair otherwise) reads the ghost column
    // that the neighbour's earlier passes wrote and will overwrite the neighbour's ghost copy of its own column, so it
    // also waits for the neighbour's flag word: seq_base + psv there means "all slab-edge items of my passes < psv are
    // finished" (the neighbour signals in pass order, see the end of the item loop).
    auto wait_deps = [&](int psv, int p, bool block) {
        if (!a.pipeline) return true;
        const int cpx = colour_of(psv) & 1;
        const bool edge = a.halo_fused && p == (cpx == 0 ? 0 : npairs - 1);
        if (psv == 0 && !edge) return true;
        const int *d = a.done + (psv > 0 ? psv - 1 : 0) * npairs;
        int pm = p == 0 ? npairs - 1 : p - 1, pp = p == npairs - 1 ? 0 : p + 1;
        if (a.halo_fused) {
            pm = max(p - 1, 0);
            pp = min(p + 1, npairs - 1);
        }
        unsigned long long t0 = 0;
        for (;;) {
            bool ok = true, lok = true;
            if (psv > 0) {
                int v0, v1, v2;
                asm volatile(MC2D_LD_DONE " %0, [%1];" : "=r"(v0) : "l"(d + pm) : "memory");
                asm volatile(MC2D_LD_DONE " %0, [%1];" : "=r"(v1) : "l"(d + p) : "memory");
                asm volatile(MC2D_LD_DONE " %0, [%1];" : "=r"(v2) : "l"(d + pp) : "memory");
                ok = lok = min(v0, min(v1, v2)) >= ipp;
            }
            if (ok && edge) {
                unsigned long long f;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(a.myflag + cpx) : "memory");
                ok = f >= a.seq_base + (unsigned long long)psv;
            }
            if (ok) return true;
            if (!block) return false;
            // bounded: these counters are written by this kernel only, so the budget can only run out on a bug or
            // after another wait has already given up; the host then reports the failure instead of hanging
            // (code 1000 + 100 pass + 10 [local counters were complete] + [slab-edge item]; detail = column pair)
            if (spin_expired(a.err, a.spin_budget_ns, t0, 1000 + 100 * psv + (lok ? 10 : 0) + (edge ? 1 : 0), p)) return true;
            __nanosleep(64);
        }
    };
    auto set_pass = [&](int psv) {
        const int colour = colour_of(psv);
        px = colour & 1;
        py = colour >> 1;
        perm = a.order + (size_t)colour * npairs * nay;
        pdU = a.partial_dU + (size_t)psv * a.dU_stride;
        pass_no = a.pass + psv;
    };
  for (int ps = 0; ps < (a.pipeline ? 1 : a.npass); ++ps) {
    set_pass(ps);
    work = a.work + ps;
    bpair = (px == 0) ? 0 : npairs - 1;   // even columns active: the first owned column is the one at the slab edge
    int hside = (px == 0) ? 0 : 1;        // ... it reads the left ghost and is pushed to the left neighbour
    int it0 = blockIdx.x * wpb + wib;
    int acx = 0, it0_rank = 0, ps0 = ps;  // column pair of the current item, its rank inside the pair, its colour pass
    const int acy0 = load_perm(it0, ps, ps0, acx, it0_rank);  // the order does not change during a sweep: fetched before the barrier
#ifdef MC2D_PROFILE_WAIT
    const long long pb0 = clock64();
#endif
    if (ps > 0) cooperative_groups::this_grid().sync();  // the cells of the previous colour are written
#ifdef MC2D_PROFILE_WAIT
    prof_bar += clock64() - pb0;
#endif
    if (a.halo_fused && !a.pipeline && it0 < ipp) {  // boundary item: wait until the neighbour has completed the ghost column it reads
        if (lane == 0) spin_until_ge(a.myflag + hside, a.seq_base + (unsigned long long)ps, a.err, a.spin_budget_ns, 2000 + ps);
        __syncwarp();
        __threadfence();
    }
    wait_deps(ps0, acx, true);  // (pipeline: the first item of a warp belongs to the first colour unless the grid is tiny)
    ItemMeta m0 = load_counts(ps0, acx, acy0, a.halo_fused && !a.pipeline && it0 < ipp);

    while (it0 < n_all) {
        if (a.pipeline) set_pass(ps0);
        const int tick = ticket();  // in flight during the staging of this item
        int it1 = 0, acy1 = -1, acx1 = 0, rank1 = 0, ps1 = ps;
        ItemMeta m1;
        bool have_it1 = false, have_m1 = false;

        const bool has_cell = m0.acy >= 0;
        if (a.pipeline) hside = (px == 0) ? 0 : 1;
        // item of the slab-edge column pair (warp uniform): barrier mode lists them first, the pipeline by pair
        const bool bitem = a.halo_fused && (a.pipeline ? acx == (px == 0 ? 0 : npairs - 1) : it0 < ipp);
        const int lcx = g.ghost + 2 * acx + px, cy = 2 * max(m0.acy, 0) + py, gcx = g.col0 + 2 * acx + px;
        const int cell = lcx * g.ny + cy;
        const uint32_t cell_gid = (uint32_t)(gcx * g.ny + cy);
        const double x_lo = g.ox + gcx * g.dx, y_lo = g.oy + cy * g.dy;
        const int n_own0 = m0.n_own;
        const int n_disp0 = n_own0 > 0 ? (a.disp_cell > 0 ? a.disp_cell : a.disp_mult * n_own0) : 0;
        const bool work0 = has_cell && (n_disp0 > 0 || a.n_gc > 0);
        const int M = (POT != POT_IDEAL && work0)
                          ? (int)__dp4a(m0.c_lo, 0x01010101u, __dp4a(m0.c_hi, 0x01010101u, 0u)) : 0;
        // slots this group needs: neighbours, own particles, room for the insertions of this visit
        const int need = work0 ? M + min(n_own0 + a.n_gc, K) : 0;
        int incl = (sl == 0) ? need : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += v;
        }
        // (leaders only carry a value: every lane now holds the prefix sum over the groups up to and including its own)
        double dU_total = 0.0;
        int start = 0;  // slots taken by the groups of earlier batches
        bool pending = work0;
        do {
            const bool work = pending && (incl - start <= R);  // a run of consecutive groups that fits
            double2 *cand = wslots + (work ? incl - need - start : 0);
            int *cand_id = wids + (work ? incl - need - start : 0);
            int n_own = work ? n_own0 : 0;
            const int n_disp = work ? n_disp0 : 0;
            const int n_gc = work ? a.n_gc : 0;
            const int Ms = work ? M : 0;
            double2 *own = cand + Ms;
            int *own_id = cand_id + Ms;

            // ---- stage: 8 neighbour cells -> cand[0..Ms), own cell -> cand[Ms..Ms+n_own): raw copies by
            //      cp.async (no registers, all in flight together: one exposed memory latency per item) ----
            if (work) {
                for (int s = sl; s < n_own; s += G) {
#ifndef MC2D_SWEEPP_IVALL
                    cp_async16_cg(&own[s], &a.pos[(size_t)cell * K + s]);
                    if (a.ids) own_id[s] = __ldcg(&a.ids[(size_t)cell * K + s]);
#else
                    cp_async16(&own[s], &a.pos[(size_t)cell * K + s]);
                    if (a.ids) own_id[s] = a.ids[(size_t)cell * K + s];
#endif
                }
                // (own cells are written by this GPU only; the neighbours' cells of a boundary item include the ghost)
                if (Ms > 0) {  // lane sl copies neighbour cells sl, sl + G, ...
                    int col3[3], row3[3];
                    nbr_rows_cols(g, lcx, cy, col3, row3);
                    for (int q = sl; q < 8; q += G) {
                        const int sh = 8 * (q & 3);
                        const uint32_t below = (1u << sh) - 1u;  // bytes of the cells before q in their word
                        const int n = (int)(((q < 4 ? m0.c_lo : m0.c_hi) >> sh) & 255u);
                        const int pre = (q < 4) ? (int)__dp4a(m0.c_lo & below, 0x01010101u, 0u)
                                                : (int)__dp4a(m0.c_hi & below, 0x01010101u,
                                                              __dp4a(m0.c_lo, 0x01010101u, 0u));
                        const int o = q + (q >= 4), oc = (o * 11) >> 5, orow = o - 3 * oc;
                        const int cb = oc == 0 ? col3[0] : (oc == 1 ? col3[1] : col3[2]);
                        const int rb = orow == 0 ? row3[0] : (orow == 1 ? row3[1] : row3[2]);
                        const double2 *src = a.pos + (size_t)(cb + rb) * K;
#ifndef MC2D_SWEEPP_IVALL
                        for (int s = 0; s < n; ++s) cp_async16_cg(&cand[pre + s], &src[s]);
#else
                        if (bitem)
                            for (int s = 0; s < n; ++s) cp_async16_cg(&cand[pre + s], &src[s]);
                        else
                            for (int s = 0; s < n; ++s) cp_async16(&cand[pre + s], &src[s]);
#endif
                    }
                }
            }
#ifdef MC2D_PROFILE_WAIT
            const long long pw0 = clock64();
#endif
            cp_async_wait_all();
            __syncwarp();
#ifdef MC2D_PROFILE_WAIT
            prof_wait += clock64() - pw0;
            prof_items += 1;
#endif
            if (!have_it1) {  // the ticket has arrived by now; the sorted cell of the next item travels during the rounds
                it1 = __shfl_sync(FULL_MASK, tick, 0);
                acy1 = load_perm(it1, ps, ps1, acx1, rank1);
                have_it1 = true;
            }
            // to cell-local coordinates, in place.  Only a cell on the rim of the periodic box (or of the shifted grid) can
            // see a periodic image: everywhere else the conversion is two exact subtractions (the same arithmetic).
            const unsigned cand_s = pin_u32((unsigned)__cvta_generic_to_shared(cand));
            const unsigned own_s = cand_s + 16u * (unsigned)Ms;
            {
                // (rim: the cell or one of its neighbours wraps — column 0 looks back at column nx-1, columns nx-2 and
                // nx-1 see the cell that the grid shift pushes across the box edge; likewise for rows)
                const bool rim = has_cell && (gcx == 0 || gcx >= g.nx - 2 || cy == 0 || cy >= g.ny - 2);
                const unsigned pe = own_s + 16u * (unsigned)n_own;
                if (__any_sync(FULL_MASK, rim)) {
                    for (unsigned q = cand_s + 16u * sl; q < pe; q += 16u * G) sts_d2(q, to_local(lds_d2(q), x_lo, y_lo, g, hLx, hLy));
                } else {
                    for (unsigned q = cand_s + 16u * sl; q < pe; q += 16u * G) {
                        const double2 v = lds_d2(q);
                        sts_d2(q, make_double2(v.x - x_lo, v.y - y_lo));
                    }
                }
            }
            __syncwarp();

            bool dirty = false;
            Philox4 batch = {0u, 0u, 0u, 0u};
            double batchE = 0.0;

            // ---- K2: displacement rounds (MC.cpp:96-123); one trial per group per round ----
            const int T_disp = warp_max_int(n_disp);
            // A batch of Philox words covers G trials.  When the displacement rounds do not fill their last batch and
            // there is one grand-canonical draw per visit, the first spare sub-lane draws the words of THAT trial
            // (same stream, same counter as the batch of its own it would otherwise get): one Philox block per item
            // instead of two for three items in four.
            const bool share_gc = MC2D_SHARE_GC_BATCH && a.n_gc == 1 && (T_disp & (G - 1)) != 0;
            {
                const int T = T_disp;
                for (int t = 0; t < T; ++t) {
                    if ((t & (G - 1)) == 0) {  // sub-lane s draws the words of trial t + s of its cell's stream
                        const uint32_t tag = (share_gc && t + sl == T) ? (MC2D_TAG_GC | 0u) : (MC2D_TAG_DISP | (uint32_t)(t + sl));
                        batch = philox4x32_10(cell_gid, tag, a.sweep_lo, a.sweep_hi, a.seed_lo, a.seed_hi);
                        batchE = mc2d_exp_variate(batch.w);  // -ln u of that trial: off the critical path
                    }
                    const int src = src_base + (t & (G - 1));
                    const uint32_t rx = __shfl_sync(FULL_MASK, batch.x, src), ry = __shfl_sync(FULL_MASK, batch.y, src);
                    const uint32_t rz = __shfl_sync(FULL_MASK, batch.z, src);
                    const double Et = __shfl_sync(FULL_MASK, batchE, src);
                    const bool act = t < n_disp;
                    const int i = act ? (int)__umulhi(rx, (uint32_t)n_own) : 0;
                    const unsigned self_s = own_s + 16u * (unsigned)i;
                    const double2 uo = act ? lds_d2(self_s) : make_double2(0.0, 0.0);
                    double2 un;
                    un.x = uo.x + a.delta * (2.0 * mc2d_u01(ry) - 1.0);
                    un.y = uo.y + a.delta * (2.0 * mc2d_u01(rz) - 1.0);
                    const bool inside = act && (un.x >= 0.0 && un.x < g.dx && un.y >= 0.0 && un.y < g.dy);
                    double dE = 0.0;
                    if (POT != POT_IDEAL) {
                        double acc = 0.0;
                        // hot loop: shared memory only; every group runs over its own list, lane sl takes candidates
                        // sl, sl + G, ...: first the frozen neighbours (no self test), then the cell's own particles.
                        // A trial that left the cell is rejected unseen (empty ranges).
                        unsigned q = cand_s + 16u * sl;
                        const unsigned qn = inside ? own_s : q, qe = inside ? own_s + 16u * (unsigned)n_own : q;
                        MC2D_UNROLL(MC2D_HOT_UNROLL)
                        for (; q < qn; q += 16u * G) {
                            const double2 c = lds_d2(q);
                            acc += pair_term<POT>(c.x - un.x, c.y - un.y, a.r2cut, a.pp, true) -
                                   pair_term<POT>(c.x - uo.x, c.y - uo.y, a.r2cut, a.pp, true);
                        }
                        for (; q < qe; q += 16u * G) {
                            const double2 c = lds_d2(q);
                            const bool notself = q != self_s;
                            acc += pair_term<POT>(c.x - un.x, c.y - un.y, a.r2cut, a.pp, notself) -
                                   pair_term<POT>(c.x - uo.x, c.y - uo.y, a.r2cut, a.pp, notself);
                        }
                        dE = escale * group_sum<G>(acc);
                    }
                    const bool accept = inside && (a.beta * dE < Et);  // MC.cpp:110-115, log form (sweep_kernel.cuh)
                    if (accept && sl == 0) sts_d2(self_s, un);
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
                            tr.pass = pass_no; tr.cell = (int)cell_gid; tr.seq = t; tr.kind = 0;
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

            // the counts of the next item travel during the GC rounds and the write-back (pipeline: only if the cells
            // it reads are final already; otherwise after this item is finished)
            if (!have_m1 && wait_deps(ps1, acx1, false)) {
                m1 = load_counts(ps1, acx1, acy1, false);  // tickets start behind the boundary items
                have_m1 = true;
            }

            // ---- K3: grand-canonical rounds (MC.cpp:189-249); insertion and deletion share one probe loop ----
            {
                const int T = warp_max_int(n_gc);
                for (int t = 0; t < T; ++t) {
                    if ((t & (G - 1)) == 0 && !share_gc) {
                        batch = philox4x32_10(cell_gid, MC2D_TAG_GC | (uint32_t)(t + sl), a.sweep_lo, a.sweep_hi,
                                              a.seed_lo, a.seed_hi);
                        batchE = mc2d_exp_variate(batch.w);
                    }
                    // (share_gc: T = 1 and the words sit in the first spare sub-lane of the last displacement batch)
                    const int src = src_base + (share_gc ? (T_disp & (G - 1)) : (t & (G - 1)));
                    const uint32_t rx = __shfl_sync(FULL_MASK, batch.x, src), ry = __shfl_sync(FULL_MASK, batch.y, src);
                    const uint32_t rz = __shfl_sync(FULL_MASK, batch.z, src);
                    const double Et = __shfl_sync(FULL_MASK, batchE, src);
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
                    const unsigned self_s = (kind == 2) ? own_s + 16u * (unsigned)i : 0u;
                    const double2 old = (kind == 2 && n_own > 0) ? lds_d2(self_s) : make_double2(0.0, 0.0);
                    const double2 probe = (kind == 1) ? pt : old;
                    double dE = 0.0;
                    if (POT != POT_IDEAL) {
                        double acc = 0.0;
                        unsigned q = cand_s + 16u * sl;
                        const unsigned qn = evaluate ? own_s : q, qe = evaluate ? own_s + 16u * (unsigned)n_own : q;
                        MC2D_UNROLL(MC2D_HOT_UNROLL)
                        for (; q < qn; q += 16u * G) {
                            const double2 c = lds_d2(q);
                            acc += pair_term<POT>(c.x - probe.x, c.y - probe.y, a.r2cut, a.pp, true);
                        }
                        for (; q < qe; q += 16u * G) {
                            const double2 c = lds_d2(q);
                            acc += pair_term<POT>(c.x - probe.x, c.y - probe.y, a.r2cut, a.pp, q != self_s);
                        }
                        acc = escale * group_sum<G>(acc);
                        dE = (kind == 2) ? -acc : acc;  // MC.cpp:200,225
                    }
                    // MC.cpp:201 and :226, per-cell form, in log form: ln of z A / (n + 1) or of n / (z A)
                    const double ln_pref = (kind == 1) ? ln_zA - lntab[n_own + 1] : lntab[n_own] - ln_zA;
                    const bool accept = evaluate && (a.beta * dE - ln_pref < Et);
                    if (sl == 0) {
                        if (accept && kind == 1) {
                            sts_d2(own_s + 16u * (unsigned)n_own, pt);
                            if (a.ids) own_id[n_own] = -1;
                        }
                        if (accept && kind == 2) {
                            sts_d2(self_s, lds_d2(own_s + 16u * (unsigned)(n_own - 1)));
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
                            tr.pass = pass_no; tr.cell = (int)cell_gid; tr.seq = n_disp + t; tr.kind = kind;
                            tr.accepted = accept ? 1 : 0; tr.reserved = evaluate ? 1 : 0;
                            const double2 o_ = (kind == 2 && evaluate) ? old : make_double2(0.0, 0.0);
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
                    const double2 u = lds_d2(own_s + 16u * (unsigned)s);
                    double x = x_lo + u.x, y = y_lo + u.y;
                    if (x >= g.Lx) x -= g.Lx; else if (x < 0.0) x += g.Lx;
                    if (y >= g.Ly) y -= g.Ly; else if (y < 0.0) y += g.Ly;
                    a.pos[(size_t)cell * K + s] = make_double2(x, y);
                    if (a.ids) a.ids[(size_t)cell * K + s] = own_id[s];
                    if (bitem) {  // ... and into the neighbour's ghost column
                        a.hpos[hside][(size_t)cy * K + s] = make_double2(x, y);
                        if (a.ids) a.hids[hside][(size_t)cy * K + s] = own_id[s];
                    }
                }
                if (sl == 0) {
                    a.cnt[cell] = n_own;
                    if (bitem) a.hcnt[hside][cy] = n_own;
                }
            }
            __syncwarp();  // the next batch / item reuses this warp's slots
            start = warp_max_int(work ? incl : start);
            pending = pending && !work;
        } while (__any_sync(FULL_MASK, pending));
        // energy change of this item, summed in lane order; the slot is private to the item
        const double dU_w = warp_sum(dU_total);
        if (lane == 0) pdU[it0_rank * npairs + acx] += dU_w;  // the item's slot in the full list
        if (a.pipeline) {  // this item's cells are written: one more finished item of (colour pass, column pair)
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(&a.done[ps0 * npairs + acx], 1);
        }
        if (bitem && a.pipeline) {
            // One more slab-edge item of pass ps0 done.  Signals go out in PASS ORDER: the flag value seq_base + f + 1
            // promises that the slab-edge items of ALL passes <= f are finished — they have read the ghost column (the
            // neighbour may overwrite it) and pushed their cells (the neighbour may read them).  Without a grid barrier
            // the edge items of pass f + 1 can finish before those of pass f, so whoever completes a pass takes a
            // (local) lock and moves the frontier over every pass that is complete by then; under the lock the plain
            // stores into the neighbours' flag words stay monotonic (no peer atomics needed).  This happens BEFORE the
            // warp may block on its next item: that item can depend, through the neighbour, on this very signal.
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                int *bd = a.done + 4 * npairs;  // [0..3] finished slab-edge items per pass, [4] passes signalled, [5] lock
                if (atomicAdd(&bd[ps0], 1) == ipp - 1) {
                    __threadfence();
                    unsigned long long t0 = 0;
                    while (atomicCAS(&bd[5], 0, 1) != 0) {
                        if (spin_expired(a.err, a.spin_budget_ns, t0, 3000 + ps0)) break;
                        __nanosleep(64);
                    }
                    for (;;) {
                        int f, n = 0;
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(bd + 4) : "memory");
                        if (f < a.npass) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(n) : "l"(bd + f) : "memory");
                        if (f >= a.npass || n < ipp) break;  // (whoever completes pass f continues from there)
                        const unsigned long long seq = a.seq_base + (unsigned long long)f + 1ull;
                        __threadfence_system();
                        *(volatile unsigned long long *)a.hflag[0] = seq;
                        *(volatile unsigned long long *)a.hflag[1] = seq;
                        __threadfence_system();
                        atomicExch(&bd[4], f + 1);
                    }
                    __threadfence();
                    atomicExch(&bd[5], 0);
                }
            }
            __syncwarp();
        }
        if (a.pipeline && !have_m1) {
#ifdef MC2D_PROFILE_WAIT
            const long long pd0 = clock64();
#endif
            wait_deps(ps1, acx1, true);
#ifdef MC2D_PROFILE_WAIT
            prof_bar += clock64() - pd0;
#endif
            m1 = load_counts(ps1, acx1, acy1, false);
            have_m1 = true;
        }
        if (bitem && !a.pipeline) {  // grid-barrier variant: one more boundary item done; the last one signals both neighbours
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                const unsigned long long done = atomicAdd(&a.myflag[3], 1ull);
                if (done == (unsigned long long)(ipp - 1)) {
                    __threadfence_system();
                    a.myflag[3] = 0;  // the next pass starts counting after the grid barrier
                    const unsigned long long seq = a.seq_base + (unsigned long long)ps + 1ull;
                    *(volatile unsigned long long *)a.hflag[0] = seq;
                    *(volatile unsigned long long *)a.hflag[1] = seq;
                    __threadfence_system();
                }
            }
        }
        __syncwarp();  // the next item's staging reuses this warp's shared memory

        it0 = it1;
        acx = acx1;
        it0_rank = rank1;
        ps0 = ps1;
        m0 = m1;
    }
  }

#ifdef MC2D_PROFILE_WAIT
    if (lane == 0 && (blockIdx.x * wpb + wib) % 601 == 0)
        printf("warp %d: total %lld cycles, staging wait %lld (%lld items), grid barrier %lld\n", blockIdx.x * wpb + wib,
               clock64() - prof_t0, prof_wait, prof_items, prof_bar);
#endif
    // ---- epilogue: move counters ----
    unsigned int cw[7] = {c_att[0], c_att[1], c_att[2], c_acc[0], c_acc[1], c_acc[2], c_ovf};
#pragma unroll
    for (int q = 0; q < 7; ++q) cw[q] = (unsigned int)warp_sum_int((int)cw[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 7; ++q) cred[wib][q] = cw[q];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        unsigned long long s = 0;
        for (int i = 0; i < wpb; ++i) s += cred[i][threadIdx.x];
        if (s) atomicAdd(&a.counters[threadIdx.x], s);
        if (s && threadIdx.x == 6) atomicAdd(&a.err[5], (int)s);
    }
}
