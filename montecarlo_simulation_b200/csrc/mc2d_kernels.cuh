// mc2d_kernels.cuh — hand-written sm_100a kernels of the 2D Monte Carlo hot path.
//
// Data layout in HBM:
//   slot layout  — the checkerboard grid the sweeps run on: cells are stored column-major
//                  (cell = column*ny + row), every cell owns K fixed particle slots of packed
//                  {double x,y} plus one int count.  A cell column is one contiguous block, so a
//                  halo exchange is a plain copy of a column.  Insert/delete are O(1) in-cell.
//   CSR layout   — the reference grid (nx = int(Lx/rcut), row-major like cell_list.cpp:103-109):
//                  positions sorted by cell key + cell_start offsets; built by key sort + boundary
//                  pass from caller-owned vectors (the ThermodynamicCalculator / CellList hooks).
// No tensor cores anywhere: there is no dense contraction on this path; the kernels are bound by
// the fp64 pipe and by latency, not by HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mc2d.h"
#include "pair_potentials.cuh"
#include "philox.h"

#define FULL_MASK 0xffffffffu

struct Grid {
    int nx, ny;      // global cell counts of the checkerboard grid (even)
    int col0, ncol;  // owned columns [col0, col0+ncol)
    int ghost;       // 1 when a ghost column is stored on each side (nranks > 1)
    int ncs;         // stored columns = ncol + 2*ghost
    int K;           // slots per cell
    double dx, dy, Lx, Ly;
    double ox, oy;   // grid origin shift of the current sweep, in [0,dx) x [0,dy)
};

// ---------------------------------------------------------------------------------------------
// exact reference distance: cell_list.cpp:57-61
//   dx = xj - xi; dx -= Lx*round(dx/Lx); r2 = dx*dx + dy*dy
// For |dx| <= Lx (positions inside the box) round(dx/Lx) is 1 iff dx >= Lx/2 and -1 iff
// dx <= -Lx/2 (exactly, including the rounding of the division — DESIGN.md §5), so the division
// is replaced by two compares.  __d*_rn intrinsics forbid FMA contraction: r2 is bit-identical to
// the CPU's, which makes the inclusive cutoff test and the neighbour counts bit-exact.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double exact_r2(double xi, double yi, double xj, double yj, double Lx, double Ly,
                                           double hLx, double hLy) {
    double dx = __dsub_rn(xj, xi);
    double dy = __dsub_rn(yj, yi);
    if (dx >= hLx) dx = __dsub_rn(dx, Lx);
    else if (dx <= -hLx) dx = __dadd_rn(dx, Lx);
    if (dy >= hLy) dy = __dsub_rn(dy, Ly);
    else if (dy <= -hLy) dy = __dadd_rn(dy, Ly);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

__device__ __forceinline__ int wrap_index(int i, int n) { return ((i % n) + n) % n; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}

// ---------------------------------------------------------------------------------------------
// views: how a kernel finds the particles of a cell
// ---------------------------------------------------------------------------------------------
struct CsrView {  // reference grid, row-major cells, sorted particles
    const double2 *pos;
    const int *perm;   // sorted position -> caller's particle index
    const int *start;  // ncell + 1
    int nx, ny;
    double cdx, cdy, Lx, Ly;
    __device__ __forceinline__ int ncells() const { return nx * ny; }
    __device__ __forceinline__ void cell(int w, int &cx, int &cy) const {
        cx = w % nx;
        cy = w / nx;
    }
    // the reference's literal 3x3 scan (cell_list.cpp:49-53): never skips, revisits cells on grids < 3 wide
    __device__ __forceinline__ bool range(int cx, int cy, int ddx, int ddy, int &b, int &n) const {
        int c = (cx + ddx + nx) % nx + ((cy + ddy + ny) % ny) * nx;
        b = start[c];
        n = start[c + 1] - b;
        return true;
    }
    __device__ __forceinline__ bool distinct3() const { return nx >= 3 && ny >= 3; }
    __device__ __forceinline__ int out_index(int j) const { return perm[j]; }
};

struct SlotView {  // checkerboard grid, column-major cells, fixed slots (+ ghost columns)
    const double2 *pos;
    const int *cnt;
    Grid g;
    __device__ __forceinline__ int ncells() const { return g.ncol * g.ny; }
    __device__ __forceinline__ void cell(int w, int &cx, int &cy) const {
        cx = g.ghost + w / g.ny;  // stored column
        cy = w % g.ny;
    }
    // physical neighbourhood: every distinct neighbour cell once
    __device__ __forceinline__ bool range(int cx, int cy, int ddx, int ddy, int &b, int &n) const {
        if ((g.nx == 2 && ddx < 0) || (g.ny == 2 && ddy < 0)) return false;
        int ncx = g.ghost ? cx + ddx : (cx + ddx + g.nx) % g.nx;
        int ncy = (cy + ddy + g.ny) % g.ny;
        int c = ncx * g.ny + ncy;
        b = c * g.K;
        n = cnt[c];
        return true;
    }
    __device__ __forceinline__ bool distinct3() const { return g.nx >= 3 && g.ny >= 3; }
    __device__ __forceinline__ int out_index(int j) const { return j; }
};

__device__ __forceinline__ double v_Lx(const CsrView &v) { return v.Lx; }
__device__ __forceinline__ double v_Ly(const CsrView &v) { return v.Ly; }
__device__ __forceinline__ double v_Lx(const SlotView &v) { return v.g.Lx; }
__device__ __forceinline__ double v_Ly(const SlotView &v) { return v.g.Ly; }

// deterministic block reduction of one double per thread; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *smem /* >= 32 doubles */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) s += smem[i];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// K4 — full-system energy + virial in ONE pass
//   replaces computeTotalEnergyCellList / computeTotalVirialCellList (thermodynamic_calculator.cpp:194-232)
//   and ThermodynamicCalculatorParallel::totalEnergy/totalVirial (thermodynamic_calculator_parallel.cpp:56-134).
// One warp per cell.  The 3x3 neighbourhood is flattened into one candidate list (lane k <-> k-th
// candidate), the cell's own particles are the outer loop.  half = 1 visits every unordered cell
// pair once (own cell j>i + 4 forward neighbours); half = 0 is the reference's directed double
// count (summed, then x0.5 by the caller) and also yields per-particle neighbour counts / local
// energies (PERP) = getNeighbors(i).size() and computeLocalEnergy.
// ---------------------------------------------------------------------------------------------
template <class V, int POT, bool PERP>
__global__ void __launch_bounds__(256) k_energy_virial(V v, PairParams pp, double r2cut, int half, double *partU,
                                                       double *partW, unsigned long long *pair_stats,
                                                       int *count_out, double *eloc_out) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const double hLx = 0.5 * v_Lx(v), hLy = 0.5 * v_Ly(v);
    double accU = 0.0, accW = 0.0;
    unsigned int tests = 0, evals = 0;
    if (w < v.ncells()) {
        int cx, cy, b0, n0;
        v.cell(w, cx, cy);
        v.range(cx, cy, 0, 0, b0, n0);
        if (n0 > 0) {
            int nb_b = 0, nb_n = 0;
            if (lane < 9) {
                int ddx = lane / 3 - 1, ddy = lane % 3 - 1;  // ddx outer, ddy inner (cell_list.cpp:49-50)
                bool use = v.range(cx, cy, ddx, ddy, nb_b, nb_n);
                if (half) use = use && ((ddx == 0 && ddy == 0) || ddx == 1 || (ddx == 0 && ddy == 1));
                if (!use) nb_n = 0;
            }
            int incl = nb_n;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                int t = __shfl_up_sync(FULL_MASK, incl, d);
                if (lane >= d) incl += t;
            }
            const int M = __shfl_sync(FULL_MASK, incl, 8);
            const int excl = incl - nb_n;
            int tb[9], ts[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                tb[c] = __shfl_sync(FULL_MASK, nb_b, c);
                ts[c] = __shfl_sync(FULL_MASK, excl, c);
            }
            for (int i = 0; i < n0; ++i) {
                const double2 pi = v.pos[b0 + i];
                double ui = 0.0, wi = 0.0;
                int ci = 0;
                for (int k = lane; k < M; k += 32) {
                    int c = 0;
#pragma unroll
                    for (int q = 1; q < 9; ++q)
                        if (k >= ts[q]) c = q;
                    int cb = tb[0], cs = ts[0];
#pragma unroll
                    for (int q = 1; q < 9; ++q)
                        if (c == q) { cb = tb[q]; cs = ts[q]; }
                    const int j = cb + (k - cs);
                    const bool same_cell = (cb == b0);
                    const bool skip = same_cell && (half ? (j <= b0 + i) : (j == b0 + i));
                    if (!skip) {
                        const double2 pj = v.pos[j];
                        const double r2 = exact_r2(pi.x, pi.y, pj.x, pj.y, v_Lx(v), v_Ly(v), hLx, hLy);
                        ++tests;
                        if (r2 <= r2cut) {  // inclusive, cell_list.cpp:62
                            double u, wv;
                            pair_uw<POT>(r2, pp, u, wv);
                            ui += u;
                            wi += wv;
                            ++ci;
                        }
                    }
                }
                accU += ui;
                accW += wi;
                evals += ci;
                if (PERP) {
                    double es = warp_sum(ui);
                    int cs = warp_sum_int(ci);
                    if (lane == 0) {
                        int o = v.out_index(b0 + i);
                        if (count_out) count_out[o] = cs;
                        if (eloc_out) eloc_out[o] = es;
                    }
                }
            }
        }
    }
    double sU = block_sum(accU, red);
    double sW = block_sum(accW, red);
    if (threadIdx.x == 0) {
        partU[blockIdx.x] = sU;
        partW[blockIdx.x] = sW;
    }
    unsigned int t = (unsigned int)warp_sum_int((int)tests), e = (unsigned int)warp_sum_int((int)evals);
    if (lane == 0 && pair_stats && (t | e)) {
        atomicAdd(&pair_stats[0], (unsigned long long)t);
        atomicAdd(&pair_stats[1], (unsigned long long)e);
    }
}

// fixed-order final reduction of per-CTA partials: out[0] += / = scale * sum(part[0..n))
__global__ void __launch_bounds__(256) k_reduce_partials(const double *part, int n, double scale, double *out,
                                                         int accumulate) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.0) + scale * s;
}

// ---------------------------------------------------------------------------------------------
// K6 — energy of arbitrary trial points against a CSR configuration
//   computeLocalEnergy over getNeighbors2(point) (cell_list.cpp:72-101), optional excluded particle.
// One warp per query.
// ---------------------------------------------------------------------------------------------
template <int POT>
__global__ void __launch_bounds__(256) k_point_energy(CsrView v, PairParams pp, double r2cut, const double2 *pts,
                                                      const int *exclude, int nq, double *e_out, int *c_out) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const double2 p = pts[q];
    const int ex = exclude ? exclude[q] : -1;
    int cx = (int)floor(p.x / v.cdx), cy = (int)floor(p.y / v.cdy);  // cell_list.cpp:76-79
    cx = wrap_index(cx, v.nx);
    cy = wrap_index(cy, v.ny);
    const double hLx = 0.5 * v.Lx, hLy = 0.5 * v.Ly;
    double e = 0.0;
    int cnt = 0;
    for (int ddx = -1; ddx <= 1; ++ddx)
        for (int ddy = -1; ddy <= 1; ++ddy) {
            int b, n;
            v.range(cx, cy, ddx, ddy, b, n);
            for (int s = lane; s < n; s += 32) {
                const int j = b + s;
                if (v.perm[j] == ex) continue;
                const double2 pj = v.pos[j];
                const double r2 = exact_r2(p.x, p.y, pj.x, pj.y, v.Lx, v.Ly, hLx, hLy);
                if (r2 <= r2cut) {
                    e += pair_u<POT>(r2, pp);
                    ++cnt;
                }
            }
        }
    e = warp_sum(e);
    cnt = warp_sum_int(cnt);
    if (lane == 0) {
        e_out[q] = e;
        if (c_out) c_out[q] = cnt;
    }
}

// ---------------------------------------------------------------------------------------------
// K1 — cell keys.  mode 0: reference grid, key = cx + cy*nx with the arithmetic of
// cell_list.cpp:103-109 (IEEE division, floor, modular wrap): bit-exact cell ids.
// mode 1: checkerboard grid (shifted origin, column-major stored index); particles outside this
// rank's slab get the sentinel key `sentinel` and sort to the end.
// flags[0] |= 1 when a coordinate is NaN or outside [0,L].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sweep_cell_of(const Grid &g, double x, double y, int &gcx, int &cy) {
    int ix = (int)floor((x - g.ox) / g.dx), iy = (int)floor((y - g.oy) / g.dy);
    gcx = wrap_index(ix, g.nx);
    cy = wrap_index(iy, g.ny);
}

__global__ void k_cell_keys(const double2 *pos, int n, int mode, int nx, int ny, double cdx, double cdy, double Lx,
                            double Ly, Grid g, int sentinel, int *keys, int *index, int *flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 p = pos[i];
    if (!(p.x >= 0.0 && p.x <= Lx && p.y >= 0.0 && p.y <= Ly)) atomicOr(&flags[0], 1);
    int key;
    if (mode == 0) {
        int cx = wrap_index((int)floor(p.x / cdx), nx);
        int cy = wrap_index((int)floor(p.y / cdy), ny);
        key = cx + cy * nx;
    } else {
        int gcx, cy;
        sweep_cell_of(g, p.x, p.y, gcx, cy);
        int lc = wrap_index(gcx - g.col0, g.nx);
        key = (lc < g.ncol) ? (lc + g.ghost) * g.ny + cy : sentinel;
    }
    keys[i] = key;
    index[i] = i;
}

// after the key sort: gather positions (and ids) into sorted order
__global__ void k_gather_sorted(const double2 *pos, const int *ids, const int *sorted_index, int n, double2 *pos_out,
                                int *id_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = sorted_index[k];
    pos_out[k] = pos[i];
    if (id_out) id_out[k] = ids ? ids[i] : i;
}

// cell-offset pass over the sorted keys: start[c] = first sorted position whose key >= c
__global__ void k_cell_bounds(const int *sorted_keys, int n, int ncell, int *start) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    int prev = (k == 0) ? -1 : min(sorted_keys[k - 1], ncell);
    int cur = (k == n) ? ncell : min(sorted_keys[k], ncell);
    for (int c = prev + 1; c <= cur; ++c) start[c] = k;
}

__global__ void k_max_cell_count(const int *start, int ncell, int *out_max) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (c < ncell) ? start[c + 1] - start[c] : 0;
    n = max(n, __shfl_xor_sync(FULL_MASK, n, 16));
    n = max(n, __shfl_xor_sync(FULL_MASK, n, 8));
    n = max(n, __shfl_xor_sync(FULL_MASK, n, 4));
    n = max(n, __shfl_xor_sync(FULL_MASK, n, 2));
    n = max(n, __shfl_xor_sync(FULL_MASK, n, 1));
    if ((threadIdx.x & 31) == 0 && n > 0) atomicMax(out_max, n);
}

// CSR (keys = stored cell index) -> slot layout.  One warp per stored cell.
__global__ void k_csr_to_slots(const double2 *spos, const int *sids, const int *start, int ncell_store, int K,
                               double2 *pos, int *cnt, int *ids, int *err) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncell_store) return;
    const int b = start[c], n = start[c + 1] - b;
    if (n > K) {
        if (lane == 0) atomicMax(&err[0], n);
        return;
    }
    for (int s = lane; s < n; s += 32) {
        pos[(size_t)c * K + s] = spos[b + s];
        if (ids) ids[(size_t)c * K + s] = sids[b + s];
    }
    if (lane == 0) cnt[c] = n;
}

// ---------------------------------------------------------------------------------------------
// K1' — steady-state re-bin after a grid shift: slot layout -> slot layout, a GATHER.
// One warp per owned NEW cell scans the 3x3 OLD cells around the same (column,row) index — a new
// cell overlaps at most those because both shifts lie in [0,dx) — keeps the particles whose
// canonical new cell is this one (ballot compaction, deterministic order, no atomics, no sort).
// With ghost columns this also performs the ownership migration of mpi_src
// (particle_exchange.cpp:305-337): a particle that crossed the slab boundary is simply picked up
// from the ghost copy by its new owner.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rebin(const double2 *pos_o, const int *cnt_o, const int *ids_o,
                                               double2 *pos_n, int *cnt_n, int *ids_n, Grid go, Grid gn, int *err) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= gn.ncol * gn.ny) return;
    const int lcx = gn.ghost + w / gn.ny, cy = w % gn.ny;
    const int gcx = gn.col0 + (lcx - gn.ghost);
    const int K = gn.K;
    const size_t obase = (size_t)(lcx * gn.ny + cy) * K;
    int nout = 0;
    for (int ddx = -1; ddx <= 1; ++ddx) {
        if (gn.nx == 2 && ddx < 0) continue;
        const int ocx = gn.ghost ? lcx + ddx : (lcx + ddx + gn.nx) % gn.nx;
        for (int ddy = -1; ddy <= 1; ++ddy) {
            if (gn.ny == 2 && ddy < 0) continue;
            const int ocy = (cy + ddy + gn.ny) % gn.ny;
            const int oc = ocx * gn.ny + ocy;
            const int n = cnt_o[oc];
            for (int s0 = 0; s0 < n; s0 += 32) {
                const int s = s0 + lane;
                bool take = false;
                double2 p = make_double2(0.0, 0.0);
                if (s < n) {
                    p = pos_o[(size_t)oc * K + s];
                    int ngx, ncy;
                    sweep_cell_of(gn, p.x, p.y, ngx, ncy);
                    take = (ngx == gcx && ncy == cy);
                }
                const unsigned m = __ballot_sync(FULL_MASK, take);
                if (take) {
                    const int off = nout + __popc(m & ((1u << lane) - 1u));
                    if (off < K) {
                        pos_n[obase + off] = p;
                        if (ids_n) ids_n[obase + off] = ids_o[(size_t)oc * K + s];
                    }
                }
                nout += __popc(m);
            }
        }
    }
    if (lane == 0) {
        cnt_n[lcx * gn.ny + cy] = min(nout, K);
        if (nout > K) atomicMax(&err[0], nout);
    }
}

// slot layout -> compact array (download).  offsets = exclusive scan of the owned cells' counts.
__global__ void k_compact(const double2 *pos, const int *cnt, const int *ids, const int *offsets, int cell0,
                          int ncell, int K, double2 *out, int *ids_out) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= ncell) return;
    const int c = cell0 + w, n = cnt[c], o = offsets[w];
    for (int s = lane; s < n; s += 32) {
        out[o + s] = pos[(size_t)c * K + s];
        if (ids_out) ids_out[o + s] = ids ? ids[(size_t)c * K + s] : -1;
    }
}

__global__ void k_sum_counts(const int *cnt, int cell0, int ncell, unsigned long long *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (i < ncell) ? cnt[cell0 + i] : 0;
    n = warp_sum_int(n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, (unsigned long long)n);
}

// ---------------------------------------------------------------------------------------------
// K2 + K3 — one colour pass of the checkerboard sweep: trial displacements and GCMC
// insert/delete trials for every cell of one colour.
//   replaces displacementMove_cell_list_dE (MC.cpp:96-123), grandCanonicalMove (MC.cpp:189-249),
//   getNeighbors/getNeighbors2 (cell_list.cpp:43-101) and, on the MPI side, do_one_parity_step_ /
//   try_displacement_ / local_energy_of_* (MC_parallel.cpp:117-216).
//
// One warp per active cell.  The 8 surrounding cells (frozen during the pass: same-colour cells
// are >= one cell >= rcut apart) are staged once into shared memory as coordinates LOCAL to the
// active cell (periodic image already applied, so the inner loop has no minimum-image code); the
// cell's own particles live in shared memory and are updated in place.  Each trial: one Philox
// call (stream = global cell id), lanes evaluate pair energies for old and new position against
// their candidates, warp-shuffle reduction gives dE, every lane takes the same Metropolis
// decision.  Detailed balance: a displacement that leaves the cell is rejected; insertions are
// uniform in the cell with acceptance z*A_cell/(n+1)*exp(-beta dU), deletions n/(z*A_cell)*exp(-beta dU)
// (the per-cell form of MC.cpp:201,226).
// MINIMG = true adds a per-pair minimum image (grids with fewer than 4 cells across).
// ---------------------------------------------------------------------------------------------
struct SweepArgs {
    double2 *pos;
    int *cnt;
    int *ids;
    Grid g;
    PairParams pp;
    double r2cut, beta, zA, delta;
    int colour, pass;
    uint32_t sweep_lo, sweep_hi, seed_lo, seed_hi;
    int disp_mult, n_gc;
    double p_ins, p_del;
    double *partial_dU;            // one per CTA
    unsigned long long *counters;  // [0..2] attempted, [3..5] accepted, [6] overflow
    mc2d_trial *trace;
    unsigned long long *trace_n;
    long long trace_cap;
};

template <int POT>
__device__ __forceinline__ double pair_e_cut(double dx, double dy, double r2cut, const PairParams &pp) {
    const double r2 = dx * dx + dy * dy;
    return (r2 <= r2cut) ? pair_u<POT>(r2, pp) : 0.0;
}

template <int POT, bool MINIMG, bool TRACE>
__global__ void __launch_bounds__(256) k_sweep(SweepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ unsigned int cred[8][7];
    const Grid &g = a.g;
    const int K = g.K;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    // per-warp shared memory: own[K] double2, own_id[K] int (if ids), nb[8K] double2
    const int nb_cap = (POT == POT_IDEAL) ? 0 : 8 * K;
    const size_t per_warp = (size_t)(K + nb_cap) * sizeof(double2) + (size_t)K * sizeof(int);
    double2 *own = reinterpret_cast<double2 *>(smem_raw + wib * per_warp);
    double2 *nb = own + K;
    int *own_id = reinterpret_cast<int *>(nb + nb_cap);

    const int nay = g.ny >> 1;
    const int nact = (g.ncol >> 1) * nay;
    double dU_total = 0.0;
    unsigned int c_att[3] = {0, 0, 0}, c_acc[3] = {0, 0, 0}, c_ovf = 0;

    if (w < nact) {
        const int px = a.colour & 1, py = a.colour >> 1;
        const int acx = w / nay, acy = w % nay;
        const int lcx = g.ghost + 2 * acx + px;  // stored column (col0 is even, so parity is global)
        const int cy = 2 * acy + py;
        const int gcx = g.col0 + 2 * acx + px;
        const int cell = lcx * g.ny + cy;
        const uint32_t cell_gid = (uint32_t)(gcx * g.ny + cy);
        const double x_lo = g.ox + gcx * g.dx, y_lo = g.oy + cy * g.dy;
        const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;
        int n_own = a.cnt[cell];
        const int n_disp = a.disp_mult * n_own;
        if (n_disp > 0 || a.n_gc > 0) {
            // ---- stage the own cell (local coordinates) ----
            for (int s = lane; s < n_own; s += 32) {
                double2 p = a.pos[(size_t)cell * K + s];
                double ux = p.x - x_lo, uy = p.y - y_lo;
                if (ux < -hLx) ux += g.Lx; else if (ux >= hLx) ux -= g.Lx;
                if (uy < -hLy) uy += g.Ly; else if (uy >= hLy) uy -= g.Ly;
                own[s] = make_double2(ux, uy);
                if (a.ids) own_id[s] = a.ids[(size_t)cell * K + s];
            }
            // ---- stage the 8 neighbour cells, flattened ----
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
                int tc[8], ts[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    tc[c] = __shfl_sync(FULL_MASK, nb_cell, c);
                    ts[c] = __shfl_sync(FULL_MASK, excl, c);
                }
                for (int k = lane; k < M; k += 32) {
                    int c = 0;
#pragma unroll
                    for (int q = 1; q < 8; ++q)
                        if (k >= ts[q]) c = q;
                    int cc = tc[0], cs = ts[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q)
                        if (c == q) { cc = tc[q]; cs = ts[q]; }
                    double2 p = a.pos[(size_t)cc * K + (k - cs)];
                    double ux = p.x - x_lo, uy = p.y - y_lo;
                    if (ux < -hLx) ux += g.Lx; else if (ux >= hLx) ux -= g.Lx;
                    if (uy < -hLy) uy += g.Ly; else if (uy >= hLy) uy -= g.Ly;
                    nb[k] = make_double2(ux, uy);
                }
            }
            __syncwarp();
            bool dirty = false;

            // ---- K2: displacement trials (MC.cpp:96-123) ----
            for (int t = 0; t < n_disp; ++t) {
                const Philox4 r = philox4x32_10(cell_gid, MC2D_TAG_DISP | (uint32_t)t, a.sweep_lo, a.sweep_hi,
                                                a.seed_lo, a.seed_hi);
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
                        for (int k = lane; k < M; k += 32) {
                            const double2 c = nb[k];
                            double dxn = c.x - un.x, dyn = c.y - un.y, dxo = c.x - uo.x, dyo = c.y - uo.y;
                            if (MINIMG) {
                                if (dxn >= hLx) dxn -= g.Lx; else if (dxn < -hLx) dxn += g.Lx;
                                if (dyn >= hLy) dyn -= g.Ly; else if (dyn < -hLy) dyn += g.Ly;
                                if (dxo >= hLx) dxo -= g.Lx; else if (dxo < -hLx) dxo += g.Lx;
                                if (dyo >= hLy) dyo -= g.Ly; else if (dyo < -hLy) dyo += g.Ly;
                            }
                            acc += pair_e_cut<POT>(dxn, dyn, a.r2cut, a.pp) - pair_e_cut<POT>(dxo, dyo, a.r2cut, a.pp);
                        }
                        for (int s = lane; s < n_own; s += 32) {
                            if (s == i) continue;
                            const double2 c = own[s];
                            acc += pair_e_cut<POT>(c.x - un.x, c.y - un.y, a.r2cut, a.pp) -
                                   pair_e_cut<POT>(c.x - uo.x, c.y - uo.y, a.r2cut, a.pp);
                        }
                        dE = warp_sum(acc);
                    }
                    accept = (dE <= 0.0) || (mc2d_u01(r.w) < exp(-a.beta * dE));  // MC.cpp:110-115
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
                const Philox4 r = philox4x32_10(cell_gid, MC2D_TAG_GC | (uint32_t)t, a.sweep_lo, a.sweep_hi,
                                                a.seed_lo, a.seed_hi);
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
                            for (int k = lane; k < M; k += 32) {
                                const double2 c = nb[k];
                                double dxn = c.x - pt.x, dyn = c.y - pt.y;
                                if (MINIMG) {
                                    if (dxn >= hLx) dxn -= g.Lx; else if (dxn < -hLx) dxn += g.Lx;
                                    if (dyn >= hLy) dyn -= g.Ly; else if (dyn < -hLy) dyn += g.Ly;
                                }
                                acc += pair_e_cut<POT>(dxn, dyn, a.r2cut, a.pp);
                            }
                            for (int s = lane; s < n_own; s += 32) {
                                const double2 c = own[s];
                                acc += pair_e_cut<POT>(c.x - pt.x, c.y - pt.y, a.r2cut, a.pp);
                            }
                            dE = warp_sum(acc);
                        }
                        evaluated = true;
                        const double av = a.zA / (double)(n_own + 1) * exp(-a.beta * dE);  // MC.cpp:201
                        accept = (av >= 1.0) || (mc2d_u01(r.w) < av);
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
                            for (int k = lane; k < M; k += 32) {
                                const double2 c = nb[k];
                                double dxo = c.x - old.x, dyo = c.y - old.y;
                                if (MINIMG) {
                                    if (dxo >= hLx) dxo -= g.Lx; else if (dxo < -hLx) dxo += g.Lx;
                                    if (dyo >= hLy) dyo -= g.Ly; else if (dyo < -hLy) dyo += g.Ly;
                                }
                                acc += pair_e_cut<POT>(dxo, dyo, a.r2cut, a.pp);
                            }
                            for (int s = lane; s < n_own; s += 32) {
                                if (s == i) continue;
                                const double2 c = own[s];
                                acc += pair_e_cut<POT>(c.x - old.x, c.y - old.y, a.r2cut, a.pp);
                            }
                            dE = -warp_sum(acc);  // MC.cpp:225
                        }
                        evaluated = true;
                        const double av = (double)n_own / a.zA * exp(-a.beta * dE);  // MC.cpp:226
                        accept = (av >= 1.0) || (mc2d_u01(r.w) < av);
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
        a.partial_dU[blockIdx.x] = s;
    }
    if (threadIdx.x < 7) {
        unsigned long long s = 0;
        for (int i = 0; i < wpb; ++i) s += cred[i][threadIdx.x];
        if (s) atomicAdd(&a.counters[threadIdx.x], s);
    }
}
