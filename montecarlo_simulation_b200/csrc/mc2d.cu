// mc2d.cu — host side of libmc2d.so: context, launch logic and the C ABI of include/mc2d.h.
// The kernels are in mc2d_kernels.cuh.  There is NO CPU fallback: without a CUDA device every
// compute entry point fails with MC2D_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "mc2d_kernels.cuh"

#if __has_include(<nccl.h>)
#include <nccl.h>
#define MC2D_HAVE_NCCL_H 1
#else
#define MC2D_HAVE_NCCL_H 0
#endif

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

#if MC2D_HAVE_NCCL_H
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) =
        nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (handle) return true;
        // RTLD_NOLOAD first: reuse the NCCL a host program (e.g. torch) already loaded
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW);
        if (!handle) {
            err = std::string("cannot load libnccl: ") + dlerror();
            return false;
        }
#define MC2D_SYM(name)                                                        \
    name = reinterpret_cast<decltype(name)>(dlsym(handle, "nccl" #name));     \
    if (!name) {                                                              \
        err = "libnccl lacks nccl" #name;                                     \
        return false;                                                         \
    }
        MC2D_SYM(GetUniqueId) MC2D_SYM(CommInitRank) MC2D_SYM(CommDestroy) MC2D_SYM(Send) MC2D_SYM(Recv)
        MC2D_SYM(GroupStart) MC2D_SYM(GroupEnd) MC2D_SYM(AllReduce) MC2D_SYM(AllGather) MC2D_SYM(GetErrorString)
#undef MC2D_SYM
        return true;
    }
};
NcclApi g_nccl;
#endif

inline int even_up(int v, int m) { return ((v + m - 1) / m) * m; }

}  // namespace

// Small results come back through page-locked memory: a copy into pageable memory (a stack variable) makes the
// driver stage and wait per copy, a pinned destination lets several copies share one stream synchronisation.
// One named field per use, so nested helpers (count_particles inside fetch_stats ...) never alias.
struct PinnedScalars {
    int flags[4];
    int errv[8];
    int maxcount, pad0, csr_kept, pad1;
    int chunk_off[8];
    unsigned long long kept, count, tn;
    EnergyResult er;  // written by k_energy_finish
    unsigned long long counters[8];
    double energy;
    double global_uwn[4];  // all-reduced U, W, N, overflow flag
};

struct LaunchShape {
    size_t smem = (size_t)-1;
    int wpb = 0, occ = 0;
};

struct mc2d_ctx {
    mc2d_params prm;
    PairParams pp;
    double r2cut = 0, beta = 0;
    double p_ins = 0.3, p_del = 0.3;
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_b = nullptr;  // slab boundary work (boundary kernels + halo pushes) next to the interior kernels
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};  // copy / kernel pipelining of upload and download
    std::string err;

    Grid g{};
    bool have_particles = false;
    bool sweep_ok = true;  // false: box too small for the checkerboard, calculator hooks only
    bool use_ids = false;
    int cur = 0;
    double2 *pos[2] = {nullptr, nullptr};
    int *cnt[2] = {nullptr, nullptr};
    int *ids[2] = {nullptr, nullptr};
    int allocK = 0;
    long long slot_allocs = 0;
    // One arena holds the sync flags and both slot buffers: a neighbour rank maps it with ONE CUDA IPC
    // handle and writes its boundary column straight into this rank's ghost column over NVLink.
    char *arena = nullptr;
    size_t off_pos[2] = {0, 0}, off_cnt[2] = {0, 0}, off_ids[2] = {0, 0};
    char *peer_base[2] = {nullptr, nullptr};  // [0] left, [1] right neighbour's arena (the same mapping when nranks == 2)
    bool peers_valid = false;                 // peer_base matches the neighbours' current arenas
    bool p2p = false;                         // halo over peer memory (k_halo_sync); false: ncclSend/ncclRecv
    unsigned long long halo_seq = 0;          // sync points so far; every rank counts the same ones
    int last_G = 0;                           // lanes per cell of the last sweep launch (0: k_sweep)
    mc2d_tuning tune{};                       // A-B switches (mc2d_set_tuning); the library reads no environment
    bool auto_margin = false;                 // cell_margin = -1: the cell width is chosen at the first upload
    // NPT volume move in flight (mc2d_rescale_try): the rescaled configuration sits in the spare buffer
    double *d_probe = nullptr;  // [0] dU, [1] x, [2] y of the last single-particle probe
    int *d_probe_sel = nullptr;  // [0] stored cell, [1] slot
    int probe_pending = 0;       // 1: a probed insertion, 2: a probed deletion can be applied
    bool rescale_pending = false;
    Grid rescale_g{};
    double rescale_Lx = 0, rescale_Ly = 0, rescale_U = 0;

    unsigned long long *d_counters = nullptr;  // 8
    double *d_energy = nullptr;                // [0] running, [1] U, [2] W, [3] N, [4] k_energy_p overflowed, [5] spare
    unsigned long long *d_pair_stats = nullptr;  // 2 + [2] particle count
    int *d_err = nullptr;                      // [0] overflow need, [1] flags, [2] max count, [3] energy item did not fit,
                                               // [4] an in-kernel wait gave up (never cleared), [5] insertions refused in this call
    DevBuf partial, trace_buf;
    DevBuf order_buf;  // work order of k_sweep_p (k_order_local), rebuilt after every re-bin
    DevBuf done_buf;   // k_sweep_p pipeline: finished items per (colour pass, column pair)
    bool order_valid = false;
    int *d_work = nullptr;  // work-item ticket counters: [p] interior / whole pass p of a sweep, [4 + p] slab boundary
    int num_sms = 0;
    DevBuf xy_in, ids_in, keys, keys_sorted, idx, idx_sorted, spos, sids, start, hist, scan_sums, pts, excl, outE, outC,
        outI, offsets;
    unsigned long long *d_trace_n = nullptr;

    uint64_t sweep = 0;
    long long launches = 0;
    double last_sweep_ms = 0, last_energy_ms = 0;
    long long last_tests = 0, last_evals = 0;
    unsigned long long attr_set = 0;  // bitmask of k_sweep instantiations whose smem attribute is set
    PinnedScalars *hp = nullptr;      // cudaHostAlloc, mapped: the finish kernels write their scalars straight into it
    PinnedScalars *hp_dev = nullptr;  // ... through this device alias
    FinishWords *d_fin = nullptr;     // accumulators / tickets of the finish kernels, zero between calls
    std::map<const void *, LaunchShape> launch_shapes;  // per kernel: shape whose attribute / occupancy is cached
    long long n_last = 0;             // last known particle count (sizes the neighbour staging area)
    int last_refused = 0;             // insertions refused (cell full) by the last mc2d_run_sweeps call
    bool ktiming = false;
    std::vector<cudaEvent_t> kev;  // event pool for per-kernel timing
    std::vector<int> kev_kind;     // 0 sweep, 1 rebin (one entry per event PAIR used in this call)
    double k_sweep_ms = 0, k_rebin_ms = 0, k_halo_ms = 0;
    long long k_sweep_launches = 0, k_rebin_launches = 0, k_halo_launches = 0;

#if MC2D_HAVE_NCCL_H
    ncclComm_t comm = nullptr;
#endif
    bool have_comm = false;
};

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            char b__[512];                                                                            \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            ctx->err = b__;                                                                           \
            return MC2D_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

#define FAIL(code, ...)                           \
    do {                                          \
        char b__[512];                            \
        snprintf(b__, sizeof b__, __VA_ARGS__);   \
        ctx->err = b__;                           \
        return (code);                            \
    } while (0)

#define DISPATCH_POT(pot, ...)                                                         \
    switch (pot) {                                                                     \
    case 0: { constexpr int POT = 0; __VA_ARGS__; } break;                             \
    case 1: { constexpr int POT = 1; __VA_ARGS__; } break;                             \
    case 2: { constexpr int POT = 2; __VA_ARGS__; } break;                             \
    case 3: { constexpr int POT = 3; __VA_ARGS__; } break;                             \
    case 4: { constexpr int POT = 4; __VA_ARGS__; } break;                             \
    default: { constexpr int POT = 5; __VA_ARGS__; } break;                            \
    }

// ------------------------------------------------------------------------------------------------
// pure host geometry
// ------------------------------------------------------------------------------------------------
extern "C" const char *mc2d_status_string(int s) {
    switch (s) {
    case MC2D_OK: return "ok";
    case MC2D_ERR_INVALID: return "invalid argument";
    case MC2D_ERR_CUDA: return "CUDA error";
    case MC2D_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case MC2D_ERR_CELL_OVERFLOW: return "cell capacity exceeded";
    case MC2D_ERR_GEOMETRY: return "box too small for the checkerboard grid";
    case MC2D_ERR_OUT_OF_BOX: return "position outside the box or NaN";
    case MC2D_ERR_NCCL: return "NCCL error";
    case MC2D_ERR_STATE: return "call out of order";
    case MC2D_ERR_PEER_TIMEOUT: return "a neighbour rank did not signal its halo in time";
    }
    return "unknown status";
}
extern "C" const char *mc2d_last_error(const mc2d_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int mc2d_abi_version(void) { return MC2D_ABI_VERSION; }
extern "C" int mc2d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int mc2d_reference_grid(double Lx, double Ly, double rcut, int *nx, int *ny, double *cdx, double *cdy) {
    if (!(Lx > 0) || !(Ly > 0) || !(rcut > 0)) return MC2D_ERR_INVALID;
    int ix = (int)(Lx / rcut), iy = (int)(Ly / rcut);  // cell_list.cpp:8-13
    if (ix < 1) ix = 1;
    if (iy < 1) iy = 1;
    if (nx) *nx = ix;
    if (ny) *ny = iy;
    if (cdx) *cdx = Lx / ix;
    if (cdy) *cdy = Ly / iy;
    return MC2D_OK;
}

extern "C" int mc2d_sweep_grid(double Lx, double Ly, double rcut, int nranks, int *nx, int *ny, double *cdx,
                               double *cdy) {
    if (!(Lx > 0) || !(Ly > 0) || !(rcut > 0) || nranks < 1) return MC2D_ERR_INVALID;
    long long ix = (long long)(Lx / rcut), iy = (long long)(Ly / rcut);
    ix -= ix % (2LL * nranks);  // even count per slab (cell_list_parallel.cpp:97-105)
    iy -= iy % 2;
    if (ix < 2LL * nranks || iy < 2 || ix > (1 << 24) || iy > (1 << 24)) return MC2D_ERR_GEOMETRY;
    if (nx) *nx = (int)ix;
    if (ny) *ny = (int)iy;
    if (cdx) *cdx = Lx / (double)ix;
    if (cdy) *cdy = Ly / (double)iy;
    return MC2D_OK;
}

extern "C" int mc2d_slab_columns(int nx, int nranks, int rank, int *col0, int *ncol) {
    if (nranks < 1 || rank < 0 || rank >= nranks || nx < 2 * nranks || nx % (2 * nranks)) return MC2D_ERR_INVALID;
    int per = nx / nranks;
    if (col0) *col0 = rank * per;
    if (ncol) *ncol = per;
    return MC2D_OK;
}

extern "C" int mc2d_halo_route(int colour, int rank, int nranks, int *send_left, int *send_to, int *recv_from) {
    if (colour < 0 || colour > 3 || nranks < 1 || rank < 0 || rank >= nranks) return MC2D_ERR_INVALID;
    int px = colour & 1;  // column parity of the active cells; a slab starts on an even column
    int left = (rank + nranks - 1) % nranks, right = (rank + 1) % nranks;
    if (send_left) *send_left = (px == 0);
    if (send_to) *send_to = (px == 0) ? left : right;
    if (recv_from) *recv_from = (px == 0) ? right : left;
    return MC2D_OK;
}

extern "C" int mc2d_sweep_schedule(uint64_t seed, uint64_t sweep, double cdx, double cdy, int shift_grid,
                                   int order[4], double *sx, double *sy) {
    double a, b;
    mc2d_schedule(seed, sweep, cdx, cdy, shift_grid, order, &a, &b);
    if (sx) *sx = a;
    if (sy) *sy = b;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
// checkerboard grid of the context for cells >= rcut * (1 + margin) wide
static int setup_grid(mc2d_ctx *ctx, double margin) {
    const mc2d_params *p = &ctx->prm;
    Grid &g = ctx->g;
    ctx->sweep_ok = true;
    int rc = mc2d_sweep_grid(p->Lx, p->Ly, p->rcut * (1.0 + margin), p->nranks, &g.nx, &g.ny, &g.dx, &g.dy);
    if (rc != MC2D_OK) {
        // The calculator hooks (reference grid) still work on such a box; only the checkerboard
        // step loop cannot run.  mc2d_upload / mc2d_run_sweeps report MC2D_ERR_GEOMETRY.
        if (p->nranks > 1)
            FAIL(rc, "box %g x %g cannot hold an even checkerboard of cells >= %g on %d slab(s)", p->Lx, p->Ly,
                 p->rcut * (1.0 + margin), p->nranks);
        ctx->sweep_ok = false;
        g.nx = g.ny = 2;
        g.dx = p->Lx / 2;
        g.dy = p->Ly / 2;
    }
    mc2d_slab_columns(g.nx, p->nranks, p->rank, &g.col0, &g.ncol);
    g.ghost = p->nranks > 1 ? 1 : 0;
    g.ncs = g.ncol + 2 * g.ghost;
    g.Lx = p->Lx;
    g.inv_dx = 1.0 / g.dx;
    g.inv_dy = 1.0 / g.dy;
    g.Ly = p->Ly;
    g.ox = g.oy = 0.0;
    return MC2D_OK;
}

static mc2d_tuning default_tuning() {
    mc2d_tuning t;
    memset(&t, 0, sizeof t);
    t.sweep_lanes = -1;
    t.fuse_passes = 1;
    t.pipeline = 1;
    t.rebin_lanes = 4;
    t.halo_overlap = 1;
    return t;
}

extern "C" int mc2d_get_tuning(mc2d_ctx *ctx, mc2d_tuning *out) {
    if (!ctx || !out) return MC2D_ERR_INVALID;
    *out = ctx->tune;
    return MC2D_OK;
}

extern "C" int mc2d_set_tuning(mc2d_ctx *ctx, const mc2d_tuning *in) {
    if (!ctx || !in) return MC2D_ERR_INVALID;
    const int G = in->sweep_lanes;
    if (G != -1 && G != 0 && G != 2 && G != 4 && G != 8) FAIL(MC2D_ERR_INVALID, "tuning.sweep_lanes must be -1, 0, 2, 4 or 8");
    if (in->rebin_lanes != 4 && in->rebin_lanes != 8) FAIL(MC2D_ERR_INVALID, "tuning.rebin_lanes must be 4 or 8");
    if (in->sweep_slots < 0 || in->sweep_warps_per_cta < 0 || in->sweep_warps_per_cta > 8 || in->spin_timeout_ms < 0 ||
        in->test_upload_lag_us < 0)
        FAIL(MC2D_ERR_INVALID, "tuning: negative or out-of-range value");
    if (in->halo_nccl != ctx->tune.halo_nccl) ctx->peers_valid = false;  // the next exchange re-negotiates the halo mode
    ctx->tune = *in;
    return MC2D_OK;
}

extern "C" int mc2d_create(const mc2d_params *p, mc2d_ctx **out) {
    if (!p || !out) return MC2D_ERR_INVALID;
    *out = nullptr;
    mc2d_ctx *ctx = new mc2d_ctx();
    *out = ctx;  // returned even on failure so that mc2d_last_error works; caller destroys it
    ctx->prm = *p;
    if (!(p->Lx > 0) || !(p->Ly > 0) || !(p->rcut > 0) || p->potential < 0 || p->potential > 5 || p->nranks < 1 ||
        p->rank < 0 || p->rank >= p->nranks || !(p->temperature > 0) || p->delta < 0 || p->activity < 0 ||
        p->cell_capacity < 0)
        FAIL(MC2D_ERR_INVALID, "mc2d_create: invalid parameters");
    int ndev = mc2d_device_count();
    if (ndev <= 0) FAIL(MC2D_ERR_NO_DEVICE, "mc2d_create: no CUDA device visible; this library has no CPU path");
    if (p->device < 0 || p->device >= ndev) FAIL(MC2D_ERR_INVALID, "mc2d_create: device %d of %d", p->device, ndev);
    ctx->dev = p->device;
    ctx->pp = PairParams{p->f_prime, p->f_d_prime, p->kappa, p->alpha};
    ctx->r2cut = p->rcut * p->rcut;
    ctx->beta = 1.0 / p->temperature;
    if (p->p_insert > 0 || p->p_delete > 0) {
        if (p->p_insert != p->p_delete || p->p_insert + p->p_delete > 1.0)
            FAIL(MC2D_ERR_INVALID, "p_insert must equal p_delete (detailed balance) and sum to <= 1");
        ctx->p_ins = p->p_insert;
        ctx->p_del = p->p_delete;
    }
    if (p->cell_margin != -1.0 && (!(p->cell_margin >= 0) || p->cell_margin > 3))
        FAIL(MC2D_ERR_INVALID, "mc2d_create: cell_margin must be in [0, 3], or -1 (chosen from the density at the first upload)");
    ctx->auto_margin = (p->cell_margin == -1.0);
    ctx->tune = default_tuning();
    if (int rc = setup_grid(ctx, ctx->auto_margin ? 0.0 : p->cell_margin)) return rc;
    CU(cudaSetDevice(ctx->dev));
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&ctx->ev0));
    CU(cudaEventCreate(&ctx->ev1));
    CU(cudaStreamCreateWithFlags(&ctx->stream_b, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming));
    for (cudaEvent_t &e : ctx->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
    CU(cudaMalloc(&ctx->d_energy, 6 * sizeof(double)));
    CU(cudaMalloc(&ctx->d_pair_stats, 4 * sizeof(unsigned long long)));
    CU(cudaMalloc(&ctx->d_err, 8 * sizeof(int)));
    CU(cudaMalloc(&ctx->d_work, 8 * sizeof(int)));
    CU(cudaMalloc(&ctx->d_probe, 4 * sizeof(double)));
    CU(cudaMalloc(&ctx->d_probe_sel, 2 * sizeof(int)));
    CU(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->dev));
    CU(cudaMalloc(&ctx->d_trace_n, sizeof(unsigned long long)));
    CU(cudaHostAlloc(&ctx->hp, sizeof(PinnedScalars), cudaHostAllocMapped));
    memset(ctx->hp, 0, sizeof(PinnedScalars));
    CU(cudaHostGetDevicePointer((void **)&ctx->hp_dev, ctx->hp, 0));
    CU(cudaMalloc(&ctx->d_fin, sizeof(FinishWords)));
    CU(cudaMemsetAsync(ctx->d_fin, 0, sizeof(FinishWords), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_energy, 0, 6 * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_err, 0, 8 * sizeof(int), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MC2D_OK;
}

static unsigned long long spin_budget_ns(const mc2d_ctx *ctx) {
    const int ms = ctx->tune.spin_timeout_ms > 0 ? ctx->tune.spin_timeout_ms : 10000;
    return (unsigned long long)ms * 1000000ull;
}

// after a synchronisation that followed halo kernels: did one of their waits give up?  (errv = copy of d_err)
static int check_peer_timeout(mc2d_ctx *ctx, const int *errv) {
    if (!errv[4]) return MC2D_OK;
    ctx->have_particles = false;
    FAIL(MC2D_ERR_PEER_TIMEOUT,
         "a neighbour rank did not signal its halo within %d ms (it failed, returned early or died); the slab contexts "
         "of all ranks are out of step: destroy them [first wait to give up: code %d, detail %d]",
         ctx->tune.spin_timeout_ms > 0 ? ctx->tune.spin_timeout_ms : 10000, errv[6], errv[7]);
}

static void close_peers(mc2d_ctx *ctx) {
    if (ctx->peer_base[0]) cudaIpcCloseMemHandle(ctx->peer_base[0]);
    if (ctx->peer_base[1] && ctx->peer_base[1] != ctx->peer_base[0]) cudaIpcCloseMemHandle(ctx->peer_base[1]);
    ctx->peer_base[0] = ctx->peer_base[1] = nullptr;
    ctx->peers_valid = false;
    ctx->p2p = false;
}

static void free_slots(mc2d_ctx *ctx) {
    close_peers(ctx);
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    for (int b = 0; b < 2; ++b) {
        ctx->pos[b] = nullptr;
        ctx->cnt[b] = nullptr;
        ctx->ids[b] = nullptr;
    }
    ctx->allocK = 0;
}

extern "C" int mc2d_destroy(mc2d_ctx *ctx) {
    if (!ctx) return MC2D_OK;
    if (ctx->stream) {
        cudaSetDevice(ctx->dev);
        cudaStreamSynchronize(ctx->stream);
#if MC2D_HAVE_NCCL_H
        if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
#endif
        free_slots(ctx);
        DevBuf *bufs[] = {&ctx->partial, &ctx->trace_buf, &ctx->xy_in, &ctx->ids_in, &ctx->keys, &ctx->keys_sorted,
                          &ctx->idx, &ctx->idx_sorted, &ctx->spos, &ctx->sids, &ctx->start, &ctx->hist, &ctx->scan_sums, &ctx->pts,
                          &ctx->excl, &ctx->outE, &ctx->outC, &ctx->outI, &ctx->offsets, &ctx->order_buf, &ctx->done_buf};
        for (DevBuf *b : bufs) b->release();
        if (ctx->d_counters) cudaFree(ctx->d_counters);
        if (ctx->d_energy) cudaFree(ctx->d_energy);
        if (ctx->d_pair_stats) cudaFree(ctx->d_pair_stats);
        if (ctx->d_fin) cudaFree(ctx->d_fin);
        if (ctx->d_err) cudaFree(ctx->d_err);
        if (ctx->d_work) cudaFree(ctx->d_work);
        if (ctx->d_probe) cudaFree(ctx->d_probe);
        if (ctx->d_probe_sel) cudaFree(ctx->d_probe_sel);
        if (ctx->d_trace_n) cudaFree(ctx->d_trace_n);
        if (ctx->hp) cudaFreeHost(ctx->hp);
        if (ctx->ev0) cudaEventDestroy(ctx->ev0);
        if (ctx->ev1) cudaEventDestroy(ctx->ev1);
        if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
        if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
        for (cudaEvent_t e : ctx->ev_chunk)
            if (e) cudaEventDestroy(e);
        if (ctx->stream_b) cudaStreamDestroy(ctx->stream_b);
        for (cudaEvent_t e : ctx->kev) cudaEventDestroy(e);
        cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of n ints: out[0..n) and out[n] = total (own kernels, three launches)
// ------------------------------------------------------------------------------------------------
static int exclusive_scan(mc2d_ctx *ctx, const int *in, int n, int *out) {
    const int nb = std::max(1, (n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    CU(ctx->scan_sums.ensure(sizeof(int) * (size_t)(nb + 1)));
    k_scan_chunks<<<nb, SCAN_TPB, 0, ctx->stream>>>(in, n, out, ctx->scan_sums.as<int>());
    k_scan_sums<<<1, SCAN_TPB, 0, ctx->stream>>>(ctx->scan_sums.as<int>(), nb, out + n);
    k_scan_add<<<nb, SCAN_TPB, 0, ctx->stream>>>(out, n, ctx->scan_sums.as<int>());
    ctx->launches += 3;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// CSR build: keys + histogram -> scan (= cell offsets) -> scatter -> rank inside the bucket + gather   (K1)
// ------------------------------------------------------------------------------------------------
// mode 0: reference grid (nx,ny,cdx,cdy given); mode 1: checkerboard grid ctx->g, slab filter.
// On return ctx->spos / ctx->sids (perm or ids) / ctx->start hold the CSR arrays; *n_kept = number of
// particles with a real (non-sentinel) key.
// bad_out != nullptr: a position outside the box is reported there instead of failing (mc2d_upload lets all
// ranks agree on the failure before any of them returns).
static int build_csr(mc2d_ctx *ctx, const double2 *d_xy, const int *d_ids, long long n, int mode, int nx, int ny,
                     double cdx, double cdy, int ncell, long long *n_kept, bool *bad_out = nullptr) {
    const int nn = (int)n;
    CU(ctx->keys.ensure(sizeof(int) * (size_t)(nn + 1)));
    CU(ctx->keys_sorted.ensure(sizeof(int) * (size_t)(nn + 1)));
    CU(ctx->idx_sorted.ensure(sizeof(int) * (size_t)(nn + 1)));
    CU(ctx->spos.ensure(sizeof(double2) * (size_t)(nn + 1)));
    CU(ctx->sids.ensure(sizeof(int) * (size_t)(nn + 1)));
    CU(ctx->start.ensure(sizeof(int) * (size_t)(ncell + 2)));
    CU(ctx->hist.ensure(sizeof(int) * 2 * (size_t)(ncell + 1)));  // histogram, then the fill counters of the scatter
    int *hist = ctx->hist.as<int>(), *fill = hist + (ncell + 1);
    CU(cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(hist, 0, sizeof(int) * 2 * (size_t)(ncell + 1), ctx->stream));
    if (nn > 0) {
        k_cell_keys<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(d_xy, nn, mode, nx, ny, cdx, cdy, ctx->prm.Lx,
                                                               ctx->prm.Ly, ctx->g, ncell, ctx->keys.as<int>(), nullptr,
                                                               ctx->d_err + 1, hist);
        ctx->launches++;
    }
    // start[c] = first sorted position of cell c; start[ncell] = particles kept (the sentinel bucket follows them)
    if (int rc = exclusive_scan(ctx, hist, ncell + 1, ctx->start.as<int>())) return rc;
    if (nn > 0) {
        k_scatter_keys<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->keys.as<int>(), nn, ncell, ctx->start.as<int>(), fill,
                                                                  ctx->idx_sorted.as<int>(), ctx->keys_sorted.as<int>());
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(ctx->hp->flags, ctx->d_err, sizeof ctx->hp->flags, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&ctx->hp->csr_kept, ctx->start.as<int>() + ncell, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    const int *flags = ctx->hp->flags;
    const int kept = ctx->hp->csr_kept;
    if (kept > 0) {
        k_stabilize_gather<<<(kept + 255) / 256, 256, 0, ctx->stream>>>(ctx->idx_sorted.as<int>(), ctx->keys_sorted.as<int>(),
                                                                      kept, ctx->start.as<int>(), d_xy, d_ids,
                                                                      ctx->spos.as<double2>(), ctx->sids.as<int>());
        ctx->launches++;
    }
    if (bad_out)
        *bad_out = (flags[1] & 1) != 0;
    else if (flags[1] & 1)
        FAIL(MC2D_ERR_OUT_OF_BOX, "a position is NaN or outside [0,%g] x [0,%g]", ctx->prm.Lx, ctx->prm.Ly);
    if (n_kept) *n_kept = kept;
    return MC2D_OK;
}

static int h2d_positions(mc2d_ctx *ctx, const double *xy, long long n) {
    if (n < 0 || n > 0x7ffffff0LL) FAIL(MC2D_ERR_INVALID, "particle count %lld out of range", n);
    CU(cudaSetDevice(ctx->dev));
    CU(ctx->xy_in.ensure(sizeof(double2) * (size_t)(n + 1)));
    if (n > 0) CU(cudaMemcpyAsync(ctx->xy_in.p, xy, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// halo exchange (K5): a cell column is contiguous, so no pack kernel is needed
// ------------------------------------------------------------------------------------------------
#if MC2D_HAVE_NCCL_H
#define NC(call)                                                                                  \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) FAIL(MC2D_ERR_NCCL, "%s: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)
#endif

// all ranks meet (and the stream drains)
static int comm_barrier(mc2d_ctx *ctx) {
#if MC2D_HAVE_NCCL_H
    CU(ctx->outI.ensure(256));
    NC(g_nccl.AllReduce(ctx->outI.p, ctx->outI.p, 1, ncclInt, ncclSum, ctx->comm, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
#endif
    return MC2D_OK;
}

// element-wise MAX of one int over the ranks (agrees on the cell capacity)
static int comm_max_int(mc2d_ctx *ctx, int *v) {
#if MC2D_HAVE_NCCL_H
    CU(ctx->outI.ensure(256));
    CU(cudaMemcpyAsync(ctx->outI.p, v, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllReduce(ctx->outI.p, ctx->outI.p, 1, ncclInt, ncclMax, ctx->comm, ctx->stream));
    CU(cudaMemcpyAsync(v, ctx->outI.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
#endif
    return MC2D_OK;
}

// Maps the arenas of the left and right neighbour (CUDA IPC; the 64-byte handles travel by ncclAllGather).
// Collective.  If any rank cannot map a neighbour, every rank falls back to ncclSend/ncclRecv.
static int connect_peers(mc2d_ctx *ctx) {
#if MC2D_HAVE_NCCL_H
    const int R = ctx->prm.nranks, me = ctx->prm.rank;
    close_peers(ctx);
    int ok = ctx->tune.halo_nccl ? 0 : 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine, ctx->arena) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    CU(ctx->outC.ensure(hb * (size_t)(R + 1)));
    char *dsend = ctx->outC.as<char>(), *drecv = dsend + hb;
    CU(cudaMemcpyAsync(dsend, &mine, hb, cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllGather(dsend, drecv, hb, ncclChar, ctx->comm, ctx->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)R);
    CU(cudaMemcpyAsync(all.data(), drecv, hb * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const int nb[2] = {(me + R - 1) % R, (me + 1) % R};
    for (int side = 0; side < 2 && ok; ++side) {
        if (side == 1 && nb[1] == nb[0]) {
            ctx->peer_base[1] = ctx->peer_base[0];
            break;
        }
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[(size_t)nb[side]], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        } else {
            ctx->peer_base[side] = static_cast<char *>(ptr);
        }
    }
    int all_ok = ok ? 0 : 1;  // max over ranks of "failed"
    if (int rc = comm_max_int(ctx, &all_ok)) return rc;
    if (all_ok != 0) close_peers(ctx);
    ctx->p2p = (all_ok == 0);
    ctx->peers_valid = true;
#endif
    return MC2D_OK;
}

static int halo_exchange(mc2d_ctx *ctx, int buf, bool left, bool right, bool wait_left = true, bool wait_right = true,
                         bool signal = true, cudaStream_t stream = nullptr) {
#if MC2D_HAVE_NCCL_H
    const Grid &g = ctx->g;
    if (!g.ghost) return MC2D_OK;
    if (!ctx->have_comm) FAIL(MC2D_ERR_STATE, "nranks > 1 but mc2d_comm_init was not called");
    if (!ctx->peers_valid)
        if (int rc = connect_peers(ctx)) return rc;
    const int first = g.ghost, last = g.ghost + g.ncol - 1, lghost = 0, rghost = g.ghost + g.ncol;
    if (ctx->p2p) {
        const size_t colp = (size_t)g.ny * g.K, colc = (size_t)g.ny;
        HaloArgs h;
        h.pos = ctx->pos[buf];
        h.cnt = ctx->cnt[buf];
        h.ids = ctx->use_ids ? ctx->ids[buf] : nullptr;
        const int ghost_col[2] = {rghost, lghost};  // my first column lands in the LEFT neighbour's RIGHT ghost
        for (int side = 0; side < 2; ++side) {
            char *pb = ctx->peer_base[side];
            h.dpos[side] = reinterpret_cast<double2 *>(pb + ctx->off_pos[buf]) + (size_t)ghost_col[side] * colp;
            h.dcnt[side] = reinterpret_cast<int *>(pb + ctx->off_cnt[buf]) + (size_t)ghost_col[side] * colc;
            h.dids[side] = ctx->use_ids ? reinterpret_cast<int *>(pb + ctx->off_ids[buf]) + (size_t)ghost_col[side] * colp
                                        : nullptr;
            h.dflag[side] = reinterpret_cast<unsigned long long *>(pb) + (side == 0 ? 1 : 0);
        }
        h.myflag = reinterpret_cast<unsigned long long *>(ctx->arena);
        h.col[0] = first;
        h.col[1] = last;
        h.push[0] = left ? 1 : 0;
        h.push[1] = right ? 1 : 0;
        h.wait[0] = wait_left ? 1 : 0;
        h.wait[1] = wait_right ? 1 : 0;
        h.signal = signal ? 1 : 0;
        h.ny = g.ny;
        h.K = g.K;
        if (signal) ++ctx->halo_seq;  // a launch that only waits waits for the sync point reached last
        h.seq = ctx->halo_seq;
        h.err = ctx->d_err;
        h.spin_budget_ns = spin_budget_ns(ctx);
        const int total = (int)colp * (h.push[0] + h.push[1]);
        const int nblk = std::max(1, std::min(4 * ctx->num_sms, (total + 255) / 256));
        k_halo_sync<<<nblk, 256, 0, stream ? stream : ctx->stream>>>(h);
        ctx->launches++;
        return MC2D_OK;
    }
    const int R = ctx->prm.nranks, me = ctx->prm.rank;
    const int lrank = (me + R - 1) % R, rrank = (me + 1) % R;
    const size_t colp = (size_t)g.ny * g.K, colc = (size_t)g.ny;
    auto col_pos = [&](int c) { return ctx->pos[buf] + (size_t)c * colp; };
    auto col_cnt = [&](int c) { return ctx->cnt[buf] + (size_t)c * colc; };
    auto col_ids = [&](int c) { return ctx->ids[buf] + (size_t)c * colp; };
    NC(g_nccl.GroupStart());
    if (left) {  // my first owned column -> left neighbour's right ghost; my right ghost <- right neighbour
        NC(g_nccl.Send(col_pos(first), colp * sizeof(double2), ncclChar, lrank, ctx->comm, ctx->stream));
        NC(g_nccl.Send(col_cnt(first), colc * sizeof(int), ncclChar, lrank, ctx->comm, ctx->stream));
        if (ctx->use_ids) NC(g_nccl.Send(col_ids(first), colp * sizeof(int), ncclChar, lrank, ctx->comm, ctx->stream));
        NC(g_nccl.Recv(col_pos(rghost), colp * sizeof(double2), ncclChar, rrank, ctx->comm, ctx->stream));
        NC(g_nccl.Recv(col_cnt(rghost), colc * sizeof(int), ncclChar, rrank, ctx->comm, ctx->stream));
        if (ctx->use_ids) NC(g_nccl.Recv(col_ids(rghost), colp * sizeof(int), ncclChar, rrank, ctx->comm, ctx->stream));
    }
    if (right) {  // my last owned column -> right neighbour's left ghost; my left ghost <- left neighbour
        NC(g_nccl.Send(col_pos(last), colp * sizeof(double2), ncclChar, rrank, ctx->comm, ctx->stream));
        NC(g_nccl.Send(col_cnt(last), colc * sizeof(int), ncclChar, rrank, ctx->comm, ctx->stream));
        if (ctx->use_ids) NC(g_nccl.Send(col_ids(last), colp * sizeof(int), ncclChar, rrank, ctx->comm, ctx->stream));
        NC(g_nccl.Recv(col_pos(lghost), colp * sizeof(double2), ncclChar, lrank, ctx->comm, ctx->stream));
        NC(g_nccl.Recv(col_cnt(lghost), colc * sizeof(int), ncclChar, lrank, ctx->comm, ctx->stream));
        if (ctx->use_ids) NC(g_nccl.Recv(col_ids(lghost), colp * sizeof(int), ncclChar, lrank, ctx->comm, ctx->stream));
    }
    NC(g_nccl.GroupEnd());
    return MC2D_OK;
#else
    (void)buf; (void)left; (void)right;
    if (ctx->g.ghost) FAIL(MC2D_ERR_NCCL, "built without nccl.h");
    return MC2D_OK;
#endif
}

// ------------------------------------------------------------------------------------------------
// slot storage
// ------------------------------------------------------------------------------------------------
static int comm_barrier(mc2d_ctx *ctx);

static int alloc_slots(mc2d_ctx *ctx, int K) {
    // Neighbours may have this arena mapped and this rank theirs: every rank closes its mappings, then all
    // meet, then each frees.  (mc2d_upload is collective when nranks > 1.)
    close_peers(ctx);
    if (ctx->arena && ctx->have_comm && ctx->prm.nranks > 1)
        if (int rc = comm_barrier(ctx)) return rc;
    free_slots(ctx);
    const size_t ncell = (size_t)ctx->g.ncs * ctx->g.ny;
    if (ncell * (size_t)K >= 0x7fffffffULL)  // kernels index slots with 32-bit integers
        FAIL(MC2D_ERR_GEOMETRY, "%zu cells x %d slots do not fit a 32-bit slot index; use more ranks", ncell, K);
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t off = 256;  // [0] flag from left, [1] flag from right, [2] CTA ticket of k_halo_sync
    for (int b = 0; b < 2; ++b) {
        ctx->off_pos[b] = off;
        off += up(sizeof(double2) * ncell * K);
    }
    for (int b = 0; b < 2; ++b) {
        ctx->off_cnt[b] = off;
        off += up(sizeof(int) * ncell);
    }
    for (int b = 0; b < 2; ++b) {
        ctx->off_ids[b] = off;
        if (ctx->use_ids) off += up(sizeof(int) * ncell * K);
    }
    CU(cudaMalloc(&ctx->arena, off));
    CU(cudaMemsetAsync(ctx->arena, 0, 256, ctx->stream));
    ctx->halo_seq = 0;
    for (int b = 0; b < 2; ++b) {
        ctx->pos[b] = reinterpret_cast<double2 *>(ctx->arena + ctx->off_pos[b]);
        ctx->cnt[b] = reinterpret_cast<int *>(ctx->arena + ctx->off_cnt[b]);
        ctx->ids[b] = ctx->use_ids ? reinterpret_cast<int *>(ctx->arena + ctx->off_ids[b]) : nullptr;
        CU(cudaMemsetAsync(ctx->cnt[b], 0, sizeof(int) * ncell, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->allocK = K;
    ctx->g.K = K;
    ctx->slot_allocs++;
    return MC2D_OK;
}

static int auto_capacity(const mc2d_ctx *ctx, int maxcount) {
    const Grid &g = ctx->g;
    const double A = g.dx * g.dy;
    double bound = 0.0;
    switch (ctx->prm.potential) {
    case MC2D_IDEAL: {
        double m = ctx->prm.activity * A;
        bound = m + 8.0 * sqrt(m) + 8.0;
    } break;
    case MC2D_LENNARD_JONES:
    case MC2D_WCA: bound = 1.6 * A + 6.0; break;  // above 2D close packing of unit-diameter cores (1.155/sigma^2)
    default: bound = 0.0;
    }
    int K = std::max(8, std::max(maxcount + maxcount / 2 + 4, (int)ceil(bound)));
    if (ctx->prm.cell_capacity > 0) K = std::max(ctx->prm.cell_capacity, maxcount);
    return even_up(K, 4);
}

// Capacity for an upload whose fullest cell holds maxcount particles.  Existing storage is kept while it still
// leaves 25 % + 2 slots of headroom and is not more than twice what a fresh allocation would take: the largest
// count of a big system moves by one or two between uploads, and a re-allocation costs a cudaMalloc plus, with
// neighbours, closing and re-mapping their memory.  (The capacity does not enter the trajectory.)
static int choose_capacity(const mc2d_ctx *ctx, int maxcount) {
    const int want = auto_capacity(ctx, maxcount);
    if (ctx->allocK > 0 && ctx->prm.cell_capacity <= 0 && ctx->allocK >= even_up(maxcount + maxcount / 4 + 2, 4) &&
        ctx->allocK <= 2 * want)
        return ctx->allocK;
    return want;
}

static int energy_on_slots(mc2d_ctx *ctx, double *U, double *W, long long *N, bool resync, bool allreduce);

extern "C" int mc2d_upload(mc2d_ctx *ctx, const double *xy, const int *ids, long long n) {
    return mc2d_upload_ex(ctx, xy, ids, n, 0u);
}

extern "C" int mc2d_upload_ex(mc2d_ctx *ctx, const double *xy, const int *ids, long long n, unsigned flags_in) {
    if (!ctx || (n > 0 && !xy) || (flags_in & ~(unsigned)MC2D_UPLOAD_NO_ENERGY)) return MC2D_ERR_INVALID;
    if (!ctx->sweep_ok)
        FAIL(MC2D_ERR_GEOMETRY, "box %g x %g cannot hold an even checkerboard of cells >= rcut=%g (needs >= 2 cells per "
             "direction); only the calculator hooks are available", ctx->prm.Lx, ctx->prm.Ly, ctx->prm.rcut);
    // Slot storage of a fitting capacity exists (every upload but a context's first): the host -> device copy goes out in
    // chunks on the second stream and k_bin_direct bins a chunk while the next one is still on the bus.
    const bool pipelined = ctx->arena && ctx->allocK > 0 && ctx->g.K == ctx->allocK && ((ids == nullptr) || ctx->ids[0]) &&
                           !ctx->tune.upload_sort && !ctx->auto_margin && n >= 4096;
    int rc = MC2D_OK;
    if (pipelined) {
        if (n > 0x7ffffff0LL) FAIL(MC2D_ERR_INVALID, "particle count %lld out of range", n);
        CU(cudaSetDevice(ctx->dev));
        CU(ctx->xy_in.ensure(sizeof(double2) * (size_t)(n + 1)));
    } else {
        rc = h2d_positions(ctx, xy, n);
    }
    if (rc) return rc;
    ctx->use_ids = (ids != nullptr);
    if (ids) {
        CU(ctx->ids_in.ensure(sizeof(int) * (size_t)(n + 1)));
        if (n > 0) CU(cudaMemcpyAsync(ctx->ids_in.p, ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
    Grid &g = ctx->g;
    if (ctx->auto_margin && ctx->prm.nranks == 1 && n > 0) {
        // Cells of exactly rcut are right for ~3 particles per cell (LJ at liquid densities); a short cutoff or a dilute
        // system leaves most cells empty and the per-cell work (staging, Philox, one GC draw) dominates.  Wider cells
        // are legal (the checkerboard only needs >= rcut): aim at 2.5 particles per cell.  Measured on 1 M WCA
        // particles at rho = 0.5: 2.4e9 -> 3.8e9 moves/s with cells of 2 rcut.
        ctx->auto_margin = false;
        int nx0, ny0;
        double dx0, dy0;
        if (mc2d_sweep_grid(ctx->prm.Lx, ctx->prm.Ly, ctx->prm.rcut, 1, &nx0, &ny0, &dx0, &dy0) == MC2D_OK) {
            const double n_c = (double)n / ((double)nx0 * ny0);
            if (n_c < 2.0) {
                double margin = std::min(2.0, sqrt(2.5 / n_c) - 1.0);
                while (margin > 0.05 && setup_grid(ctx, margin) != MC2D_OK) margin *= 0.5;  // small boxes
                if (!ctx->sweep_ok || g.nx < 4 || g.ny < 4)
                    if (int rc2 = setup_grid(ctx, 0.0)) return rc2;
                free_slots(ctx);
            }
        }
    }
    g.ox = g.oy = 0.0;
    const int ncell = g.ncs * g.ny;
    const int cell0 = g.ghost * g.ny, nown = g.ncol * g.ny;
    long long kept = 0;
    int maxcount = 0;
    // fault injection (tools/check_multi_gpu.py): odd ranks fall behind between the collective and their conversion
    // kernel, so the neighbours' halo pushes land first
    // A bad position on one rank fails the upload on every rank: the flag travels with the collective maximum.
    constexpr int kOutOfBox = 0x7fffffff;
    auto test_lag = [&]() {
#ifdef MC2D_FAULT_INJECTION
        if ((ctx->prm.rank & 1) && ctx->tune.test_upload_lag_us > 0) usleep((useconds_t)ctx->tune.test_upload_lag_us);
#endif
    };
    // Slot storage of a fitting capacity already exists (a re-upload): bin straight into the spare buffer, no sort.
    // Every rank takes the same branch: allocK and use_ids are the same on all ranks, and so is the capacity
    // computed from the global maximum below.
    bool direct = ctx->arena && ctx->allocK > 0 && g.K == ctx->allocK && (!ctx->use_ids || ctx->ids[0]) &&
                  !ctx->tune.upload_sort;
    if (!direct && !ctx->tune.upload_sort) {
        // No fitting slot storage yet (the first upload of a context): count the particles per cell (the histogram pass
        // of the key sort), agree on the capacity, allocate — and then bin directly like every later upload.
        const int nn = (int)n;
        CU(ctx->keys.ensure(sizeof(int) * (size_t)(nn + 1)));
        CU(ctx->hist.ensure(sizeof(int) * 2 * (size_t)(ncell + 1)));
        CU(cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream));
        CU(cudaMemsetAsync(ctx->hist.p, 0, sizeof(int) * (size_t)(ncell + 1), ctx->stream));
        if (nn > 0) {
            k_cell_keys<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->xy_in.as<double2>(), nn, 1, g.nx, g.ny, g.dx, g.dy,
                                                                   ctx->prm.Lx, ctx->prm.Ly, g, ncell, ctx->keys.as<int>(),
                                                                   nullptr, ctx->d_err + 1, ctx->hist.as<int>());
            k_max_int<<<(ncell + 255) / 256, 256, 0, ctx->stream>>>(ctx->hist.as<int>(), ncell, ctx->d_err + 2);
            ctx->launches += 2;
        }
        CU(cudaMemcpyAsync(ctx->hp->flags, ctx->d_err, sizeof ctx->hp->flags, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
        int mc0 = (ctx->hp->flags[1] & 1) ? kOutOfBox : ctx->hp->flags[2];
        if (g.ghost && ctx->have_comm)
            if ((rc = comm_max_int(ctx, &mc0))) return rc;
        if (mc0 == kOutOfBox)
            FAIL(MC2D_ERR_OUT_OF_BOX, "a position is NaN or outside [0,%g] x [0,%g]", ctx->prm.Lx, ctx->prm.Ly);
        const int K0 = choose_capacity(ctx, mc0);
        if (K0 > 1024)
            FAIL(MC2D_ERR_CELL_OVERFLOW, "a cell holds %d particles; capacity %d exceeds the 1024-slot limit", mc0, K0);
        if (K0 != ctx->allocK || (ctx->use_ids && !ctx->ids[0]))
            if ((rc = alloc_slots(ctx, K0))) return rc;
        direct = true;
    }
    const bool was_direct = direct;
    if (direct) {
        ctx->have_particles = false;  // the spare buffer may be the current one: a failed upload leaves nothing
        const int K = g.K;
        int *tag = ctx->use_ids ? ctx->ids[1] : nullptr;
        if (!tag) {
            CU(ctx->idx.ensure(sizeof(int) * (size_t)ncell * K));
            tag = ctx->idx.as<int>();
        }
        CU(cudaMemsetAsync(ctx->cnt[1], 0, sizeof(int) * (size_t)ncell, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream));
        CU(cudaMemsetAsync(ctx->d_pair_stats + 2, 0, sizeof(unsigned long long), ctx->stream));
        const int nch = pipelined ? 4 : 1;
        for (int c = 0; c < nch && n > 0; ++c) {
            const long long i0 = n * c / nch, i1 = n * (c + 1) / nch;
            if (pipelined) {
                CU(cudaMemcpyAsync(ctx->xy_in.as<double2>() + i0, xy + 2 * i0, sizeof(double2) * (size_t)(i1 - i0),
                                   cudaMemcpyHostToDevice, ctx->stream_b));
                CU(cudaEventRecord(ctx->ev_chunk[c], ctx->stream_b));
                CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c], 0));
            }
            k_bin_direct<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, ctx->stream>>>(
                ctx->xy_in.as<double2>() + i0, (int)(i1 - i0), (int)i0, ctx->prm.Lx, ctx->prm.Ly, g, K, ctx->pos[1], ctx->cnt[1],
                tag, ctx->d_err, ctx->d_pair_stats + 2);
            ctx->launches++;
        }
        CU(cudaMemcpyAsync(ctx->hp->flags, ctx->d_err, sizeof ctx->hp->flags, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(&ctx->hp->kept, ctx->d_pair_stats + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
        const int *flags = ctx->hp->flags;
        const unsigned long long nk = ctx->hp->kept;
        kept = (long long)nk;
        maxcount = (flags[1] & 1) ? kOutOfBox : flags[2];
        if (g.ghost && ctx->have_comm)
            if ((rc = comm_max_int(ctx, &maxcount))) return rc;
        if (maxcount == kOutOfBox)
            FAIL(MC2D_ERR_OUT_OF_BOX, "a position is NaN or outside [0,%g] x [0,%g]", ctx->prm.Lx, ctx->prm.Ly);
        test_lag();
        if (choose_capacity(ctx, maxcount) != ctx->allocK) {
            direct = false;  // the cells got fuller (or much emptier) than the storage was made for: sorted path
        } else {
            k_order_cells<<<(nown * 8 + 255) / 256, 256, 0, ctx->stream>>>(
                ctx->pos[1], ctx->cnt[1], tag, ids ? ctx->ids_in.as<int>() : nullptr, cell0, nown, K, ctx->pos[0],
                ctx->cnt[0], ctx->use_ids ? ctx->ids[0] : nullptr);
            ctx->launches++;
            ctx->cur = 0;
        }
    }
    if (!direct) {
        bool bad = false;
        rc = build_csr(ctx, ctx->xy_in.as<double2>(), ids ? ctx->ids_in.as<int>() : nullptr, n, 1, g.nx, g.ny, g.dx, g.dy,
                       ncell, &kept, &bad);
        if (rc) return rc;
        k_max_cell_count<<<(ncell + 255) / 256, 256, 0, ctx->stream>>>(ctx->start.as<int>(), ncell, ctx->d_err + 2);
        ctx->launches++;
        CU(cudaMemcpyAsync(&ctx->hp->maxcount, ctx->d_err + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        maxcount = bad ? kOutOfBox : ctx->hp->maxcount;
        if (g.ghost && ctx->have_comm)  // every slab must use the same capacity: the halo copies whole slot columns
            if ((rc = comm_max_int(ctx, &maxcount))) return rc;
        if (maxcount == kOutOfBox)
            FAIL(MC2D_ERR_OUT_OF_BOX, "a position is NaN or outside [0,%g] x [0,%g]", ctx->prm.Lx, ctx->prm.Ly);
        int K = choose_capacity(ctx, maxcount);
        if (K > 1024)
            FAIL(MC2D_ERR_CELL_OVERFLOW, "a cell holds %d particles; capacity %d exceeds the 1024-slot limit", maxcount, K);
        if (K != ctx->allocK || (ctx->use_ids && !ctx->ids[0])) {
            rc = alloc_slots(ctx, K);
            if (rc) return rc;
        }
        ctx->cur = 0;
        if (!was_direct) test_lag();
        CU(cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream));
        k_csr_to_slots<<<(nown + 7) / 8, 256, 0, ctx->stream>>>(ctx->spos.as<double2>(), ctx->sids.as<int>(),
                                                                ctx->start.as<int>(), cell0, nown, g.K, ctx->pos[0],
                                                                ctx->cnt[0], ctx->use_ids ? ctx->ids[0] : nullptr,
                                                                ctx->d_err);
        ctx->launches++;
    }
    ctx->have_particles = true;
    ctx->rescale_pending = false;
    ctx->probe_pending = 0;
    ctx->order_valid = false;
    ctx->n_last = kept;
    // The sweep counter (part of every Philox counter) is NOT reset: a context that is re-uploaded
    // between batches (the end-to-end pattern: upload, sweep, download) must not replay the random
    // streams of its previous batch.  Move counters are left alone too (mc2d_reset_counters).
    if (g.ghost && ctx->have_comm) {
        rc = halo_exchange(ctx, ctx->cur, true, true);
        if (rc) return rc;
    }
    if (flags_in & MC2D_UPLOAD_NO_ENERGY) {
        CU(cudaMemsetAsync(ctx->d_energy, 0, sizeof(double), ctx->stream));  // the running energy counts from zero
    } else if (!g.ghost || ctx->have_comm) {
        double U, W;
        long long N;
        rc = energy_on_slots(ctx, &U, &W, &N, true, false);
        if (rc) return rc;
    }
    if (g.ghost) CU(cudaMemcpyAsync(ctx->hp->errv, ctx->d_err, sizeof ctx->hp->errv, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (g.ghost)
        if ((rc = check_peer_timeout(ctx, ctx->hp->errv))) return rc;
    return MC2D_OK;
}

// enqueues the particle count of the owned cells; ctx->hp->count is valid after the next stream synchronisation
static int count_particles_async(mc2d_ctx *ctx) {
    const Grid &g = ctx->g;
    const int cell0 = g.ghost * g.ny, ncell = g.ncol * g.ny;
    const int nb = std::max(1, std::min((ncell + 1023) / 1024, 4 * ctx->num_sms));
    k_sum_counts<<<nb, 256, 0, ctx->stream>>>(ctx->cnt[ctx->cur], cell0, ncell, ctx->d_fin, &ctx->hp_dev->count, ctx->d_energy + 3);
    ctx->launches++;
    return MC2D_OK;
}

static int count_particles(mc2d_ctx *ctx, long long *n) {
    if (int rc = count_particles_async(ctx)) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    *n = (long long)ctx->hp->count;
    ctx->n_last = *n;
    return MC2D_OK;
}

extern "C" int mc2d_num_particles(mc2d_ctx *ctx, long long *n_local) {
    if (!ctx || !n_local) return MC2D_ERR_INVALID;
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    CU(cudaSetDevice(ctx->dev));
    return count_particles(ctx, n_local);
}

// download, optionally with the energy/virial reduction running behind the device -> host copies (mc2d_sample)
static int download_impl(mc2d_ctx *ctx, double *xy, int *ids, long long cap, long long *n_out, bool with_energy, bool resync,
                         bool allreduce, double *U, double *W, long long *N) {
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    CU(cudaSetDevice(ctx->dev));
    const Grid &g = ctx->g;
    const int cell0 = g.ghost * g.ny, ncell = g.ncol * g.ny;
    CU(ctx->offsets.ensure(sizeof(int) * (size_t)(ncell + 2)));
    int rc = exclusive_scan(ctx, ctx->cnt[ctx->cur] + cell0, ncell, ctx->offsets.as<int>());  // offsets[ncell] = particle count
    if (rc) return rc;
    // The cells are compacted in four blocks of columns; the device -> host copy of a block runs on the second stream
    // while the next block is compacted.  The block boundaries (offsets at 1/4, 2/4, 3/4 of the cells) come back with the
    // particle count in one synchronisation.
    const int nch = ncell >= 4096 ? 4 : 1;
    int cb[5];
    for (int c = 0; c <= nch; ++c) {
        cb[c] = (int)((long long)ncell * c / nch);
        CU(cudaMemcpyAsync(&ctx->hp->chunk_off[c], ctx->offsets.as<int>() + cb[c], sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    const long long n = ctx->hp->chunk_off[nch];
    ctx->n_last = n;
    *n_out = n;
    if (n > cap) FAIL(MC2D_ERR_INVALID, "download needs room for %lld particles, caller gave %lld", n, cap);
    if (n == 0) {
        if (with_energy) return energy_on_slots(ctx, U, W, N, resync, allreduce);
        return MC2D_OK;
    }
    CU(ctx->spos.ensure(sizeof(double2) * (size_t)n));
    CU(ctx->sids.ensure(sizeof(int) * (size_t)n));
    for (int c = 0; c < nch; ++c) {
        const int nc = cb[c + 1] - cb[c];
        const long long o0 = ctx->hp->chunk_off[c], o1 = ctx->hp->chunk_off[c + 1];
        if (nc <= 0) continue;
        k_compact<<<(nc * 4 + 255) / 256, 256, 0, ctx->stream>>>(ctx->pos[ctx->cur], ctx->cnt[ctx->cur],
                                                             ctx->use_ids ? ctx->ids[ctx->cur] : nullptr,
                                                             ctx->offsets.as<int>() + cb[c], cell0 + cb[c], nc, g.K,
                                                             ctx->spos.as<double2>(), ids ? ctx->sids.as<int>() : nullptr);
        ctx->launches++;
        if (o1 <= o0) continue;
        cudaStream_t cs = nch > 1 ? ctx->stream_b : ctx->stream;
        if (nch > 1) {
            CU(cudaEventRecord(ctx->ev_chunk[c], ctx->stream));
            CU(cudaStreamWaitEvent(cs, ctx->ev_chunk[c], 0));
        }
        if (xy) CU(cudaMemcpyAsync(xy + 2 * o0, ctx->spos.as<double2>() + o0, sizeof(double2) * (size_t)(o1 - o0), cudaMemcpyDeviceToHost, cs));
        if (ids) CU(cudaMemcpyAsync(ids + o0, ctx->sids.as<int>() + o0, sizeof(int) * (size_t)(o1 - o0), cudaMemcpyDeviceToHost, cs));
    }
    // (mc2d_sample: the reduction runs on the main stream while the copies are still on the bus)
    if (with_energy)
        if ((rc = energy_on_slots(ctx, U, W, N, resync, allreduce))) return rc;
    if (nch > 1) CU(cudaStreamSynchronize(ctx->stream_b));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return MC2D_OK;
}

extern "C" int mc2d_download(mc2d_ctx *ctx, double *xy, int *ids, long long cap, long long *n_out) {
    if (!ctx || !n_out) return MC2D_ERR_INVALID;
    return download_impl(ctx, xy, ids, cap, n_out, false, false, false, nullptr, nullptr, nullptr);
}

extern "C" int mc2d_sample(mc2d_ctx *ctx, int resync, int allreduce, double *U, double *W, long long *N, double *xy, int *ids,
                           long long cap, long long *n_out) {
    if (!ctx || !n_out) return MC2D_ERR_INVALID;
    double u = 0, w = 0;
    long long n = 0;
    const int rc = download_impl(ctx, xy, ids, cap, n_out, true, resync != 0, allreduce != 0, &u, &w, &n);
    if (rc) return rc;
    if (U) *U = u;
    if (W) *W = w;
    if (N) *N = n;
    return MC2D_OK;
}

// Launch shape of k_energy_p (fixed-size candidate array per group): lanes per cell, neighbour slots and
// bytes per group, warps per CTA.  G = 0: the grid or the cells do not fit that layout.
struct ItemLayout {
    int G = 0, nb_cap = 0, wpb = 8;
    size_t per_group = 0;
};

static int item_layout(mc2d_ctx *ctx, ItemLayout *out) {
    const Grid &g = ctx->g;
    const int K = g.K;
    const bool ideal = ctx->prm.potential == MC2D_IDEAL;
    int wpb = 8, G = 0, nb_cap = 0;
    size_t per_group = 0;  // bytes between the candidate arrays of two groups
    if (g.nx >= 4 && g.ny >= 4 && K <= 255) {
        const double n_c = (double)std::max<long long>(ctx->n_last, 1) / std::max(1, g.ncol * g.ny);
        G = (5.0 * n_c > 40.0) ? 8 : 4;  // own cell + 4 forward neighbours
        // A quarter warp (8 lanes x 16 B) must touch 32 distinct banks: the groups inside it sit G*16 bytes
        // apart modulo 128 (profiles/r01_notes.md: an unpadded stride doubled every LDS.128).
        const size_t resid = (size_t)(G * 16) % 128;
        auto stride_of = [&](int nb) {
            const size_t b = (size_t)(nb + K) * sizeof(double2);
            size_t st = b / 128 * 128 + resid;
            return st < b ? st + 128 : st;
        };
        // forward-neighbour slots: generous when that still leaves 32 warps per SM (7 KB per warp); an item
        // that does not fit makes the caller fall back to k_energy_virial_g
        const int nb_want = ideal ? 0 : std::min(4 * K, even_up((int)(4.0 * n_c * 1.8 + 12.0), 4));
        const int nb_min = ideal ? 0 : std::min(4 * K, even_up((int)(4.0 * n_c * 1.3 + 8.0), 4));
        const size_t warp_budget = 7 * 1024;
        bool found = false;
        for (int nb = nb_want; nb >= nb_min && !found; nb -= 4)
            if (stride_of(nb) * (32 / G) <= warp_budget) {
                nb_cap = nb;
                found = true;
            }
        if (!found) {  // big cells: fewer warps per SM
            nb_cap = nb_want;
            while (wpb > 1 && stride_of(nb_cap) * (32 / G) * wpb > 200 * 1024) wpb >>= 1;
            if (stride_of(nb_cap) * (32 / G) * wpb > 200 * 1024) G = 0;
        }
        per_group = stride_of(nb_cap);
    }
    out->G = G;
    out->nb_cap = nb_cap;
    out->wpb = wpb;
    out->per_group = per_group;
    return MC2D_OK;
}

// per colour and column pair: active cells by falling (occupancy, candidates) -> ctx->order_buf
static SlabHalo no_slab_halo() {
    SlabHalo sh;
    memset(&sh, 0, sizeof sh);
    return sh;
}

static int build_order(mc2d_ctx *ctx, const SlabHalo *wait = nullptr, int *zero_work = nullptr, int n_work = 0,
                       int *zero_done = nullptr, int n_done = 0) {
    const Grid &g = ctx->g;
    const int ncell = g.ncol * g.ny;
    const size_t smem = sizeof(int) * (size_t)(3 * g.ny + g.ny / 2);
    if (smem > 200 * 1024) FAIL(MC2D_ERR_GEOMETRY, "grid has %d rows: too many for k_order_local", g.ny);
    LaunchShape &ls = ctx->launch_shapes[(const void *)k_order_local];
    if (smem > 40 * 1024 && !ls.occ) {
        CU(cudaFuncSetAttribute(k_order_local, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ls.occ = 1;
    }
    CU(ctx->order_buf.ensure(sizeof(int) * (size_t)ncell));
    k_order_local<<<dim3(g.ncol / 2, 4), ORDL_WARPS * 32, smem, ctx->stream>>>(g, ctx->cnt[ctx->cur], ctx->order_buf.as<int>(),
                                                                                  wait ? *wait : no_slab_halo(), zero_work,
                                                                                  n_work, zero_done, n_done);
    ctx->launches++;
    ctx->order_valid = true;
    return MC2D_OK;
}


// K4 on the slot layout through the sorted items (k_energy_p).  *done = false: not applicable here (tiny grid,
// huge cells) or an item did not fit its shared memory — the caller uses k_energy_virial_g instead.
static int launch_energy_items(mc2d_ctx *ctx, double *U, double *W, bool *done, bool resync, bool allreduce, bool *reduced) {
    *done = false;
    const Grid &g = ctx->g;
    if (g.nx < 4 || g.ny < 4) return MC2D_OK;
    if (sizeof(int) * (size_t)(3 * g.ny + g.ny / 2) > 200 * 1024) return MC2D_OK;  // too tall for k_order_local
    ItemLayout lay;
    if (int rc = item_layout(ctx, &lay)) return rc;
    if (!lay.G) return MC2D_OK;
    if (!ctx->order_valid)  // any permutation of the cells is correct; a fresh one balances the warps
        if (int rc = build_order(ctx)) return rc;
    const int G = lay.G, wpb = lay.wpb;
    const int nitems = (g.ncol / 2) * ((g.ny / 2 + 32 / G - 1) / (32 / G));
    const int nblocks = (4 * nitems + wpb - 1) / wpb;
    const size_t smem = lay.per_group * (32 / G) * wpb;
    const int cap = lay.nb_cap + g.K;
    const int nparts = nblocks * wpb;  // one partial sum per warp
    CU(ctx->partial.ensure(sizeof(double) * 2 * (size_t)nparts));
    double *pU = ctx->partial.as<double>(), *pW = pU + nparts;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    DISPATCH_POT(ctx->prm.potential, {
        if (G == 4) {
            LaunchShape &ls = ctx->launch_shapes[(const void *)k_energy_p<POT, 4>];
            if (!ls.occ) {
                CU(cudaFuncSetAttribute(k_energy_p<POT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                ls.occ = 1;
            }
            k_energy_p<POT, 4><<<nblocks, wpb * 32, smem, ctx->stream>>>(
                ctx->pos[ctx->cur], ctx->cnt[ctx->cur], ctx->order_buf.as<int>(), g, ctx->pp, ctx->r2cut, nitems, cap,
                (int)lay.per_group, pU, pW, ctx->d_fin->pair_stats, &ctx->d_fin->energy_ovf);
        } else {
            LaunchShape &ls = ctx->launch_shapes[(const void *)k_energy_p<POT, 8>];
            if (!ls.occ) {
                CU(cudaFuncSetAttribute(k_energy_p<POT, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                ls.occ = 1;
            }
            k_energy_p<POT, 8><<<nblocks, wpb * 32, smem, ctx->stream>>>(
                ctx->pos[ctx->cur], ctx->cnt[ctx->cur], ctx->order_buf.as<int>(), g, ctx->pp, ctx->r2cut, nitems, cap,
                (int)lay.per_group, pU, pW, ctx->d_fin->pair_stats, &ctx->d_fin->energy_ovf);
        }
    });
    // U and W in fixed order; the CTA that finishes last writes them, the pair statistics and the overflow flag into the
    // pinned scalars and re-synchronises the running energy (unless an item overflowed its shared memory — never seen —
    // and the caller's fallback redoes the reduction): one kernel, one stream synchronisation, no copies
    k_energy_finish<<<2, 1024, 0, ctx->stream>>>(pU, nparts, ctx->d_energy, resync ? 1 : 0, ctx->d_fin, &ctx->hp_dev->er);
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 2;
    PinnedScalars *hp = ctx->hp;
    *reduced = false;
#if MC2D_HAVE_NCCL_H
    if (allreduce) {  // {U, W, N, overflow} summed over the slabs on the device, behind the same synchronisation
        if (!ctx->have_comm) FAIL(MC2D_ERR_STATE, "allreduce requested without a communicator");
        ncclResult_t r = g_nccl.AllReduce(ctx->d_energy + 1, ctx->d_energy + 1, 4, ncclDouble, ncclSum, ctx->comm, ctx->stream);
        if (r != ncclSuccess) FAIL(MC2D_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
        CU(cudaMemcpyAsync(hp->global_uwn, ctx->d_energy + 1, sizeof hp->global_uwn, cudaMemcpyDeviceToHost, ctx->stream));
        *reduced = true;
    }
#else
    (void)allreduce;
#endif
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    // (an item that overflowed its shared memory — never seen — sends EVERY rank to the fallback: the flag is summed)
    if (hp->er.ovf || (*reduced && hp->global_uwn[3] != 0.0)) {
        *reduced = false;
        return MC2D_OK;
    }
    const double *uw = hp->er.uw;
    const unsigned long long *ps = hp->er.ps;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_energy_ms = ms;
    ctx->last_tests = (long long)ps[0];
    ctx->last_evals = (long long)ps[1];
    *U = uw[0];
    *W = uw[1];
    *done = true;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// K4 launches
// ------------------------------------------------------------------------------------------------
template <class V>
static int launch_energy(mc2d_ctx *ctx, const V &view, int ncells, int half, bool perp, int *count_out,
                         double *eloc_out, double *U, double *W) {
    // per-particle outputs: one warp per cell; totals only: 8 lanes per cell, 4 cells per warp
    constexpr int GE = 8;
    const int wpb = 8, cells_per_block = perp ? wpb : wpb * (32 / GE);
    const int nblocks = std::max(1, (ncells + cells_per_block - 1) / cells_per_block);
    CU(ctx->partial.ensure(sizeof(double) * 2 * (size_t)nblocks));
    double *pU = ctx->partial.as<double>(), *pW = pU + nblocks;
    CU(cudaMemsetAsync(ctx->d_pair_stats, 0, 2 * sizeof(unsigned long long), ctx->stream));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    DISPATCH_POT(ctx->prm.potential, {
        if (perp)
            k_energy_virial<V, POT, true><<<nblocks, wpb * 32, 0, ctx->stream>>>(
                view, ctx->pp, ctx->r2cut, half, pU, pW, ctx->d_pair_stats, count_out, eloc_out);
        else
            k_energy_virial_g<V, POT, GE><<<nblocks, wpb * 32, 0, ctx->stream>>>(view, ctx->pp, ctx->r2cut, half, pU,
                                                                                   pW, ctx->d_pair_stats);
    });
    const double scale = half ? 1.0 : 0.5;  // directed double count is halved (thermodynamic_calculator.cpp:211,232)
    k_reduce_partials<<<1, 1024, 0, ctx->stream>>>(pU, nblocks, scale, ctx->d_energy + 1, 0);
    k_reduce_partials<<<1, 1024, 0, ctx->stream>>>(pW, nblocks, scale, ctx->d_energy + 2, 0);
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 3;
    double uw[2];
    unsigned long long ps[2];
    CU(cudaMemcpyAsync(uw, ctx->d_energy + 1, sizeof uw, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ps, ctx->d_pair_stats, sizeof ps, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_energy_ms = ms;
    ctx->last_tests = (long long)ps[0];
    ctx->last_evals = (long long)ps[1];
    *U = uw[0];
    *W = uw[1];
    return MC2D_OK;
}

static int energy_on_slots(mc2d_ctx *ctx, double *U, double *W, long long *N, bool resync, bool allreduce) {
    const Grid &g = ctx->g;
    SlotView v{ctx->pos[ctx->cur], ctx->cnt[ctx->cur], g};
    const int half = (g.nx >= 3 && g.ny >= 3) ? 1 : 0;
    bool done = false;
    int rc = MC2D_OK;
    // the particle count rides on the synchronisation of the energy launch
    if ((rc = count_particles_async(ctx))) return rc;
    const bool want_global = allreduce && ctx->g.ghost;
    bool reduced = false;
    if (!ctx->tune.energy_unstaged) rc = launch_energy_items(ctx, U, W, &done, resync, want_global, &reduced);
    if (rc) return rc;
    bool pending = false;  // work on the stream behind the last synchronisation
    if (!done) {
        rc = launch_energy(ctx, v, g.ncol * g.ny, half, false, nullptr, nullptr, U, W);
        if (rc) return rc;
        if (resync) {
            CU(cudaMemcpyAsync(ctx->d_energy, ctx->d_energy + 1, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            pending = true;
        }
    }
    long long n = (long long)ctx->hp->count;
    ctx->n_last = n;
    if (want_global && reduced) {
        *U = ctx->hp->global_uwn[0];
        *W = ctx->hp->global_uwn[1];
        n = (long long)llround(ctx->hp->global_uwn[2]);
    } else if (want_global) {
#if MC2D_HAVE_NCCL_H
        if (!ctx->have_comm) FAIL(MC2D_ERR_STATE, "allreduce requested without a communicator");
        double h[3] = {*U, *W, (double)n};
        CU(cudaMemcpyAsync(ctx->d_energy + 1, h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
        ncclResult_t r = g_nccl.AllReduce(ctx->d_energy + 1, ctx->d_energy + 1, 3, ncclDouble, ncclSum, ctx->comm,
                                          ctx->stream);
        if (r != ncclSuccess) FAIL(MC2D_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
        CU(cudaMemcpyAsync(h, ctx->d_energy + 1, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *U = h[0];
        *W = h[1];
        n = (long long)llround(h[2]);
        pending = false;
#endif
    }
    if (pending) CU(cudaStreamSynchronize(ctx->stream));
    if (N) *N = n;
    return MC2D_OK;
}

extern "C" int mc2d_total_energy_virial(mc2d_ctx *ctx, int resync, int allreduce, double *U, double *W, long long *N) {
    if (!ctx) return MC2D_ERR_INVALID;
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    CU(cudaSetDevice(ctx->dev));
    double u, w;
    long long n;
    int rc = energy_on_slots(ctx, &u, &w, &n, resync != 0, allreduce != 0);
    if (rc) return rc;
    if (U) *U = u;
    if (W) *W = w;
    if (N) *N = n;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// single-particle probes (Gibbs-ensemble particle exchange)
// ------------------------------------------------------------------------------------------------
static int probe_common(mc2d_ctx *ctx, int mode, double x, double y, long long index, double *xo, double *yo, double *dU) {
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    if (ctx->prm.nranks != 1) FAIL(MC2D_ERR_INVALID, "single-particle probes need a single-rank context");
    CU(cudaSetDevice(ctx->dev));
    ctx->probe_pending = 0;
    ctx->rescale_pending = false;
    const Grid &g = ctx->g;
    const int ncell = g.ncol * g.ny;
    if (mode == 1) {
        long long n = 0;
        if (int rc = count_particles(ctx, &n)) return rc;
        if (index < 0 || index >= n) FAIL(MC2D_ERR_INVALID, "particle index %lld outside [0, %lld)", index, n);
        CU(ctx->offsets.ensure(sizeof(int) * (size_t)(ncell + 2)));
        if (int rc = exclusive_scan(ctx, ctx->cnt[ctx->cur], ncell, ctx->offsets.as<int>())) return rc;
    } else if (!(x >= 0.0 && x < ctx->prm.Lx && y >= 0.0 && y < ctx->prm.Ly)) {
        FAIL(MC2D_ERR_OUT_OF_BOX, "probe point (%g, %g) outside [0, %g) x [0, %g)", x, y, ctx->prm.Lx, ctx->prm.Ly);
    }
    DISPATCH_POT(ctx->prm.potential, {
        k_probe_slots<POT><<<1, 32, 0, ctx->stream>>>(ctx->pos[ctx->cur], ctx->cnt[ctx->cur], g, ctx->pp, ctx->r2cut, mode,
                                                      make_double2(x, y), index, ctx->offsets.as<int>(), ctx->d_probe,
                                                      ctx->d_probe_sel);
    });
    ctx->launches++;
    double h[3];
    CU(cudaMemcpyAsync(h, ctx->d_probe, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (dU) *dU = h[0];
    if (xo) *xo = h[1];
    if (yo) *yo = h[2];
    ctx->probe_pending = mode == 1 ? 2 : 1;
    return MC2D_OK;
}

extern "C" int mc2d_probe_point(mc2d_ctx *ctx, double x, double y, double *dU) {
    if (!ctx) return MC2D_ERR_INVALID;
    return probe_common(ctx, 0, x, y, 0, nullptr, nullptr, dU);
}

extern "C" int mc2d_probe_particle(mc2d_ctx *ctx, long long index, double *x, double *y, double *dU) {
    if (!ctx) return MC2D_ERR_INVALID;
    return probe_common(ctx, 1, 0.0, 0.0, index, x, y, dU);
}

extern "C" int mc2d_apply_probe(mc2d_ctx *ctx, int what) {
    if (!ctx) return MC2D_ERR_INVALID;
    const int pending = ctx->probe_pending;
    ctx->probe_pending = 0;
    if (what == 0) return MC2D_OK;
    if (what != pending) FAIL(MC2D_ERR_STATE, "no matching probe pending (asked %d, pending %d)", what, pending);
    CU(cudaSetDevice(ctx->dev));
    CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    k_apply_probe<<<1, 32, 0, ctx->stream>>>(ctx->pos[ctx->cur], ctx->cnt[ctx->cur], ctx->use_ids ? ctx->ids[ctx->cur] : nullptr,
                                             ctx->g.K, what, ctx->d_probe, ctx->d_probe_sel, ctx->d_energy, ctx->d_err);
    ctx->launches++;
    int full = 0;
    CU(cudaMemcpyAsync(&full, ctx->d_err, sizeof full, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (full) FAIL(MC2D_ERR_CELL_OVERFLOW, "the cell of the inserted particle has no free slot (capacity %d)", ctx->g.K);
    ctx->n_last += what == 1 ? 1 : -1;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// NPT volume move
// ------------------------------------------------------------------------------------------------
extern "C" int mc2d_rescale_try(mc2d_ctx *ctx, double Lx_new, double Ly_new, double *U_new, double *W_new) {
    if (!ctx) return MC2D_ERR_INVALID;
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    if (ctx->prm.nranks != 1) FAIL(MC2D_ERR_INVALID, "volume moves need a single-rank context");
    if (!(Lx_new > 0) || !(Ly_new > 0)) FAIL(MC2D_ERR_INVALID, "box lengths must be positive");
    ctx->rescale_pending = false;
    CU(cudaSetDevice(ctx->dev));
    const Grid go = ctx->g;
    Grid gn = go;
    gn.Lx = Lx_new;
    gn.Ly = Ly_new;
    gn.dx = Lx_new / go.nx;
    gn.dy = Ly_new / go.ny;
    if (gn.dx < ctx->prm.rcut || gn.dy < ctx->prm.rcut)
        FAIL(MC2D_ERR_GEOMETRY, "box %g x %g on %d x %d cells: cells would be narrower than rcut = %g", Lx_new, Ly_new,
             go.nx, go.ny, ctx->prm.rcut);
    gn.inv_dx = 1.0 / gn.dx;
    gn.inv_dy = 1.0 / gn.dy;
    const double rx = Lx_new / go.Lx, ry = Ly_new / go.Ly;
    gn.ox = go.ox * rx;  // cell walls scale with the particles
    gn.oy = go.oy * ry;
    if (gn.ox >= gn.dx) gn.ox = 0.0;
    if (gn.oy >= gn.dy) gn.oy = 0.0;
    const int cur = ctx->cur, nb = 1 - cur;
    const int ncell = go.ncs * go.ny;
    const long long nslot = (long long)ncell * go.K;
    k_scale_slots<<<(unsigned)((nslot + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->pos[cur], ctx->cnt[cur], ctx->use_ids ? ctx->ids[cur] : nullptr, ctx->pos[nb], ctx->cnt[nb],
        ctx->use_ids ? ctx->ids[nb] : nullptr, ncell, go.K, rx, ry, Lx_new, Ly_new);
    ctx->launches++;
    // the reduction runs on the rescaled state; the context goes back to the present one afterwards
    const double Lx_old = ctx->prm.Lx, Ly_old = ctx->prm.Ly;
    ctx->g = gn;
    ctx->prm.Lx = Lx_new;
    ctx->prm.Ly = Ly_new;
    ctx->cur = nb;
    double U = 0, W = 0;
    long long N = 0;
    const int rc = energy_on_slots(ctx, &U, &W, &N, false, false);
    ctx->g = go;
    ctx->prm.Lx = Lx_old;
    ctx->prm.Ly = Ly_old;
    ctx->cur = cur;
    if (rc) return rc;
    ctx->rescale_pending = true;
    ctx->rescale_g = gn;
    ctx->rescale_Lx = Lx_new;
    ctx->rescale_Ly = Ly_new;
    ctx->rescale_U = U;
    if (U_new) *U_new = U;
    if (W_new) *W_new = W;
    return MC2D_OK;
}

extern "C" int mc2d_rescale_finish(mc2d_ctx *ctx, int accept) {
    if (!ctx) return MC2D_ERR_INVALID;
    if (!ctx->rescale_pending) FAIL(MC2D_ERR_STATE, "no volume move pending (mc2d_rescale_try first)");
    ctx->rescale_pending = false;
    if (!accept) return MC2D_OK;
    CU(cudaSetDevice(ctx->dev));
    ctx->g = ctx->rescale_g;
    ctx->prm.Lx = ctx->rescale_Lx;
    ctx->prm.Ly = ctx->rescale_Ly;
    ctx->cur = 1 - ctx->cur;
    CU(cudaMemcpyAsync(ctx->d_energy, &ctx->rescale_U, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// ThermodynamicCalculator / CellList hooks on the reference grid
// ------------------------------------------------------------------------------------------------
static int build_reference_csr(mc2d_ctx *ctx, const double *xy, long long n, CsrView *view) {
    int nx, ny;
    double cdx, cdy;
    mc2d_reference_grid(ctx->prm.Lx, ctx->prm.Ly, ctx->prm.rcut, &nx, &ny, &cdx, &cdy);
    if ((long long)nx * ny > 0x3fffffffLL) FAIL(MC2D_ERR_GEOMETRY, "reference grid %d x %d too large", nx, ny);
    int rc = h2d_positions(ctx, xy, n);
    if (rc) return rc;
    long long kept = 0;
    rc = build_csr(ctx, ctx->xy_in.as<double2>(), nullptr, n, 0, nx, ny, cdx, cdy, nx * ny, &kept);
    if (rc) return rc;
    *view = CsrView{ctx->spos.as<double2>(), ctx->sids.as<int>(), ctx->start.as<int>(), nx, ny, cdx, cdy,
                    ctx->prm.Lx, ctx->prm.Ly};
    return MC2D_OK;
}

extern "C" int mc2d_calc_energy_virial(mc2d_ctx *ctx, const double *xy, long long n, double *U, double *W) {
    if (!ctx || (n > 0 && !xy)) return MC2D_ERR_INVALID;
    CsrView v;
    int rc = build_reference_csr(ctx, xy, n, &v);
    if (rc) return rc;
    double u = 0, w = 0;
    const int half = (v.nx >= 3 && v.ny >= 3) ? 1 : 0;
    rc = launch_energy(ctx, v, v.nx * v.ny, half, false, nullptr, nullptr, &u, &w);
    if (rc) return rc;
    if (U) *U = u;
    if (W) *W = w;
    return MC2D_OK;
}

extern "C" int mc2d_calc_cell_ids(mc2d_ctx *ctx, const double *xy, long long n, int *cell_out) {
    if (!ctx || (n > 0 && (!xy || !cell_out))) return MC2D_ERR_INVALID;
    int nx, ny;
    double cdx, cdy;
    mc2d_reference_grid(ctx->prm.Lx, ctx->prm.Ly, ctx->prm.rcut, &nx, &ny, &cdx, &cdy);
    int rc = h2d_positions(ctx, xy, n);
    if (rc) return rc;
    if (n == 0) return MC2D_OK;
    const int nn = (int)n;
    CU(ctx->keys.ensure(sizeof(int) * (size_t)nn));
    CU(ctx->idx.ensure(sizeof(int) * (size_t)nn));
    CU(cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream));
    k_cell_keys<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->xy_in.as<double2>(), nn, 0, nx, ny, cdx, cdy,
                                                           ctx->prm.Lx, ctx->prm.Ly, ctx->g, nx * ny,
                                                           ctx->keys.as<int>(), nullptr, ctx->d_err + 1, nullptr);
    ctx->launches++;
    CU(cudaMemcpyAsync(cell_out, ctx->keys.p, sizeof(int) * (size_t)nn, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return MC2D_OK;
}

extern "C" int mc2d_calc_neighbor_counts(mc2d_ctx *ctx, const double *xy, long long n, int *count_out,
                                         double *local_energy_out) {
    if (!ctx || (n > 0 && (!xy || !count_out))) return MC2D_ERR_INVALID;
    CsrView v;
    int rc = build_reference_csr(ctx, xy, n, &v);
    if (rc) return rc;
    if (n == 0) return MC2D_OK;
    CU(ctx->outC.ensure(sizeof(int) * (size_t)n));
    CU(ctx->outE.ensure(sizeof(double) * (size_t)n));
    CU(cudaMemsetAsync(ctx->outC.p, 0, sizeof(int) * (size_t)n, ctx->stream));
    double u, w;
    rc = launch_energy(ctx, v, v.nx * v.ny, 0, true, ctx->outC.as<int>(), ctx->outE.as<double>(), &u, &w);
    if (rc) return rc;
    CU(cudaMemcpyAsync(count_out, ctx->outC.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (local_energy_out)
        CU(cudaMemcpyAsync(local_energy_out, ctx->outE.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MC2D_OK;
}

extern "C" int mc2d_calc_point_energies(mc2d_ctx *ctx, const double *xy, long long n, const double *points,
                                        const int *exclude, long long nq, double *energy_out, int *count_out) {
    if (!ctx || (n > 0 && !xy) || nq < 0 || (nq > 0 && (!points || !energy_out))) return MC2D_ERR_INVALID;
    CsrView v;
    int rc = build_reference_csr(ctx, xy, n, &v);
    if (rc) return rc;
    if (nq == 0) return MC2D_OK;
    CU(ctx->pts.ensure(sizeof(double2) * (size_t)nq));
    CU(ctx->excl.ensure(sizeof(int) * (size_t)nq));
    CU(ctx->outE.ensure(sizeof(double) * (size_t)nq));
    CU(ctx->outC.ensure(sizeof(int) * (size_t)nq));
    CU(cudaMemcpyAsync(ctx->pts.p, points, sizeof(double2) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    if (exclude) CU(cudaMemcpyAsync(ctx->excl.p, exclude, sizeof(int) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    const int q = (int)nq;
    DISPATCH_POT(ctx->prm.potential, {
        k_point_energy<POT><<<(q + 7) / 8, 256, 0, ctx->stream>>>(v, ctx->pp, ctx->r2cut, ctx->pts.as<double2>(),
                                                                  exclude ? ctx->excl.as<int>() : nullptr, q,
                                                                  ctx->outE.as<double>(), ctx->outC.as<int>());
    });
    ctx->launches++;
    CU(cudaMemcpyAsync(energy_out, ctx->outE.p, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    if (count_out) CU(cudaMemcpyAsync(count_out, ctx->outC.p, sizeof(int) * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return MC2D_OK;
}

extern "C" int mc2d_calc_neighbor_list(mc2d_ctx *ctx, const double *xy, long long n, long long query_index,
                                       double qx, double qy, int *out_j, double *out_r2, int cap, int *count_out) {
    if (!ctx || (n > 0 && !xy) || cap < 0 || !count_out || (cap > 0 && (!out_j || !out_r2)) || query_index >= n)
        return MC2D_ERR_INVALID;
    CsrView v;
    int rc = build_reference_csr(ctx, xy, n, &v);
    if (rc) return rc;
    int qs = -1;
    if (query_index >= 0) {  // sorted position of the caller's particle: invert the permutation on the host
        std::vector<int> perm((size_t)n);
        CU(cudaMemcpyAsync(perm.data(), ctx->sids.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (long long k = 0; k < n; ++k)
            if (perm[(size_t)k] == (int)query_index) {
                qs = (int)k;
                break;
            }
    }
    CU(ctx->outC.ensure(sizeof(int) * (size_t)(cap + 2)));
    CU(ctx->outE.ensure(sizeof(double) * (size_t)(cap + 1)));
    int *d_count = ctx->outC.as<int>() + cap;
    k_neighbor_list<<<1, 32, 0, ctx->stream>>>(v, ctx->r2cut, qs, qx, qy, ctx->outC.as<int>(), ctx->outE.as<double>(),
                                               cap, d_count);
    ctx->launches++;
    int cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    *count_out = cnt;
    const int m = std::min(cnt, cap);
    if (m > 0) {
        CU(cudaMemcpy(out_j, ctx->outC.p, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(out_r2, ctx->outE.p, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost));
    }
    return MC2D_OK;
}

extern "C" int mc2d_calc_energy_virial_bruteforce(mc2d_ctx *ctx, const double *xy, long long n, double *U, double *W) {
    if (!ctx || (n > 0 && !xy)) return MC2D_ERR_INVALID;
    int rc = h2d_positions(ctx, xy, n);
    if (rc) return rc;
    const int nn = (int)n, nblocks = std::max(1, (nn + 255) / 256);
    CU(ctx->partial.ensure(sizeof(double) * 2 * (size_t)nblocks));
    double *pU = ctx->partial.as<double>(), *pW = pU + nblocks;
    DISPATCH_POT(ctx->prm.potential, {
        k_bruteforce<POT><<<nblocks, 256, 0, ctx->stream>>>(ctx->xy_in.as<double2>(), nn, ctx->prm.Lx, ctx->prm.Ly,
                                                            1.0 / ctx->prm.Lx, 1.0 / ctx->prm.Ly, ctx->pp,
                                                            ctx->r2cut, pU, pW);
    });
    k_reduce_partials<<<1, 1024, 0, ctx->stream>>>(pU, nblocks, 1.0, ctx->d_energy + 1, 0);
    k_reduce_partials<<<1, 1024, 0, ctx->stream>>>(pW, nblocks, 1.0, ctx->d_energy + 2, 0);
    ctx->launches += 3;
    double uw[2];
    CU(cudaMemcpyAsync(uw, ctx->d_energy + 1, sizeof uw, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (U) *U = uw[0];
    if (W) *W = uw[1];
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// the step loop
// ------------------------------------------------------------------------------------------------
template <int POT, bool MINIMG, bool TRACE>
static int launch_sweep_kernel(mc2d_ctx *ctx, const SweepArgs &a, int nblocks, int wpb, size_t smem, int inst) {
    if (!(ctx->attr_set & (1ull << inst))) {
        CU(cudaFuncSetAttribute(k_sweep<POT, MINIMG, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->attr_set |= (1ull << inst);
    }
    k_sweep<POT, MINIMG, TRACE><<<nblocks, wpb * 32, smem, ctx->stream>>>(a);
    ctx->launches++;
    return MC2D_OK;
}

static int ktime_begin(mc2d_ctx *ctx, int kind) {
    if (!ctx->ktiming) return MC2D_OK;
    const size_t need = 2 * (ctx->kev_kind.size() + 1);
    while (ctx->kev.size() < need) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ctx->kev.push_back(e);
    }
    ctx->kev_kind.push_back(kind);
    CU(cudaEventRecord(ctx->kev[need - 2], ctx->stream));
    return MC2D_OK;
}
static int ktime_end(mc2d_ctx *ctx) {
    if (!ctx->ktiming) return MC2D_OK;
    CU(cudaEventRecord(ctx->kev[2 * ctx->kev_kind.size() - 1], ctx->stream));
    return MC2D_OK;
}

template <int POT, int G, bool TRACE>
static int launch_sweep_p(mc2d_ctx *ctx, const SweepArgs &a, int wpb, size_t smem, int R, cudaStream_t stream,
                          int reserve_ctas) {
    // per context (= per device) and instantiation: resident CTAs per SM for this launch shape; the query costs
    // tens of microseconds, so it is not repeated per launch
    LaunchShape &ls = ctx->launch_shapes[(const void *)k_sweep_p<POT, G, TRACE>];
    if (smem != ls.smem || wpb != ls.wpb) {
        CU(cudaFuncSetAttribute(k_sweep_p<POT, G, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sweep_p<POT, G, TRACE>, wpb * 32, smem));
        ls.occ = std::max(1, nb);
        ls.smem = smem;
        ls.wpb = wpb;
    }
    const int occ = ls.occ;
    const int want = (a.nitems + wpb - 1) / wpb;
    // reserve_ctas: resident-CTA slots left free for the boundary kernel that runs next to this one
    const int grid = std::max(1, std::min(want, ctx->num_sms * occ - reserve_ctas));
    if (a.npass > 1) {  // all passes of a sweep in one launch: the grid barrier needs a cooperative launch
        SweepArgs aa = a;
        int Rv = R;
        void *kargs[] = {&aa, &Rv};
        CU(cudaLaunchCooperativeKernel((const void *)k_sweep_p<POT, G, TRACE>, dim3(grid), dim3(wpb * 32), kargs, smem,
                                       stream));
    } else {
        k_sweep_p<POT, G, TRACE><<<grid, wpb * 32, smem, stream>>>(a, R);
    }
    ctx->launches++;
    return MC2D_OK;
}

static int enqueue_stats_fetch(mc2d_ctx *ctx, bool counters_too);

static int run_sweeps_impl(mc2d_ctx *ctx, long long nsweeps, const mc2d_sweep_plan *plan, mc2d_trial *trace,
                           long long trace_cap, long long *trace_n, bool with_stats = false) {
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    ctx->rescale_pending = false;  // the re-bin overwrites the spare buffer
    ctx->probe_pending = 0;
    if (!plan || plan->disp_per_particle < 0 || plan->gc_per_cell < 0 || plan->disp_per_cell < 0 || nsweeps < 0)
        FAIL(MC2D_ERR_INVALID, "bad sweep plan");
    if (plan->disp_per_particle * 1024LL + plan->disp_per_cell + plan->gc_per_cell >= (1 << 28))
        FAIL(MC2D_ERR_INVALID, "too many trials per cell visit");
    CU(cudaSetDevice(ctx->dev));
    Grid &g = ctx->g;
    const int K = g.K;
    const bool ideal = ctx->prm.potential == MC2D_IDEAL;
    const bool minimg = (g.nx < 4 || g.ny < 4);
    const size_t per_warp = (size_t)(K + (ideal ? 0 : 8 * K)) * sizeof(double2) + (size_t)K * sizeof(int);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
    if (per_warp * wpb > 200 * 1024) FAIL(MC2D_ERR_CELL_OVERFLOW, "cell capacity %d needs too much shared memory", K);
    const int nact = (g.ncol / 2) * (g.ny / 2);
    // Kernel choice: k_sweep_p (G lanes per cell, persistent warps over the sorted work list, R candidate
    // slots of shared memory per warp shared out by need) unless the grid needs the per-pair minimum image or
    // one cell with all its neighbours cannot fit a warp's slots; then k_sweep (one warp per cell).
    // MC2D_SWEEP_G overrides the lanes per cell (0 = k_sweep), for tuning and tests.
    int G = 0, R = 0;
    const bool order_fits = sizeof(int) * (size_t)(3 * g.ny + g.ny / 2) <= 200 * 1024;  // k_order_local's shared memory
    if (!minimg && K <= 255 && order_fits) {
        const double n_c = (double)std::max<long long>(ctx->n_last, 1) / std::max(1, g.ncol * g.ny);
        G = (9.0 * n_c > 64.0) ? 8 : 4;
        if (ctx->tune.sweep_lanes >= 0) G = ctx->tune.sweep_lanes;
        if (G) {
            const size_t slot_bytes = sizeof(double2) + (ctx->use_ids ? sizeof(int) : 0);
            const int need_max = (ideal ? 0 : 8 * K) + K;  // one cell, every slot of its neighbourhood full
            wpb = 8;
            if (ctx->tune.sweep_warps_per_cta > 0) wpb = ctx->tune.sweep_warps_per_cta;
            // 6 KB per warp: 3 CTAs x 8 warps per SM (the register limit) with ~80 KB of the SM left as L1 for the
            // count / order / position loads.  Measured on config #4: 61 us per pass for R = 320..512 slots,
            // 67 us at 576 (the third CTA no longer fits), 79 us at 256 (one item in four runs in two batches).
            R = (int)((6 * 1024) / slot_bytes) & ~3;  // a multiple of 4: the warps' regions stay 16-byte aligned with ids
            if (ctx->tune.sweep_slots > 0) R = (std::max(need_max, ctx->tune.sweep_slots) + 3) & ~3;
            if (R < need_max) {  // big cells: fewer warps per SM
                R = (need_max + 3) & ~3;
                while (wpb > 1 && (size_t)R * slot_bytes * wpb > 200 * 1024) wpb >>= 1;
                if ((size_t)R * slot_bytes * wpb > 200 * 1024) G = 0;
            }
        }
    }
    if (!G) {
        wpb = 8;
        while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
    }
    const bool persistent = G != 0;
    ctx->last_G = G;
    const int cells_per_block = G ? wpb * (32 / G) : wpb;
    // k_sweep_p: a work item = 32/G cells of one column pair; k_sweep: one CTA per cells_per_block
    const int nitems = G ? (g.ncol / 2) * ((g.ny / 2 + 32 / G - 1) / (32 / G)) : 0;
    const int nblocks = persistent ? nitems : std::max(1, (nact + cells_per_block - 1) / cells_per_block);
    CU(ctx->partial.ensure(sizeof(double) * 4 * (size_t)nblocks));
    // per-CTA (per-item) dU accumulators live across the whole batch and are reduced once at its end
    CU(cudaMemsetAsync(ctx->partial.p, 0, sizeof(double) * 4 * (size_t)nblocks, ctx->stream));
    if (trace) {
        CU(ctx->trace_buf.ensure(sizeof(mc2d_trial) * (size_t)trace_cap));
        CU(cudaMemsetAsync(ctx->d_trace_n, 0, sizeof(unsigned long long), ctx->stream));
    }
    const int ncell_owned = g.ncol * g.ny;
    const int rebin_G = ctx->tune.rebin_lanes;
    ctx->kev_kind.clear();
    CU(cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_err + 5, 0, sizeof(int), ctx->stream));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    // Slab mode with peer-memory halos: the boundary column pair of a pass (and the two boundary columns of
    // a re-bin) run on a second stream, followed there by the push into the neighbour and the signal, while
    // the interior runs on the main stream.  The neighbour's data and the NVLink latency arrive during the
    // interior kernel; the wait in front of the next boundary kernel normally finds its flag already set.
    const int npairs = g.ncol / 2;
    bool overlap = persistent && g.ghost && ctx->have_comm && K <= 64 && npairs >= 4 && !trace;
    if (overlap && !ctx->peers_valid)
        if (int rc = connect_peers(ctx)) return rc;
    overlap = overlap && ctx->p2p;
    overlap = overlap && ctx->tune.halo_overlap != 0;
    cudaStream_t sa = ctx->stream, sb = ctx->stream_b;
    auto join = [&]() -> int {  // each stream continues after everything queued so far on the other one
        CU(cudaEventRecord(ctx->ev_a, sa));
        CU(cudaEventRecord(ctx->ev_b, sb));
        CU(cudaStreamWaitEvent(sb, ctx->ev_a, 0));
        CU(cudaStreamWaitEvent(sa, ctx->ev_b, 0));
        return MC2D_OK;
    };
    auto launch_rebin = [&](cudaStream_t st, const Grid &go, const Grid &gn, int nb, int w0, int n0, int w1, int n1) -> int {
        const double2 *po = ctx->pos[ctx->cur];
        const int *co = ctx->cnt[ctx->cur], *io = ctx->use_ids ? ctx->ids[ctx->cur] : nullptr;
        int *in = ctx->use_ids ? ctx->ids[nb] : nullptr;
        const int n = n0 + n1;
        const double sdx = gn.ox - go.ox, sdy = gn.oy - go.oy;
        if (K <= 64 && rebin_G == 4 && fabs(sdx) > 1e-9 * gn.dx && fabs(sdy) > 1e-9 * gn.dy && gn.nx >= 3 && gn.ny >= 3)
            k_rebin4<4, false><<<(n + 63) / 64, 256, 0, st>>>(po, co, io, ctx->pos[nb], ctx->cnt[nb], in, gn, sdx > 0 ? 1 : -1,
                                                        sdy > 0 ? 1 : -1, w0, n0, w1, n1, ctx->d_err, no_slab_halo());
        else if (K <= 64 && rebin_G == 4)  // 4 lanes per new cell, 8 cells per warp
            k_rebin_g<4><<<(n + 63) / 64, 256, 0, st>>>(po, co, io, ctx->pos[nb], ctx->cnt[nb], in, gn, gn.ox - go.ox,
                                                         gn.oy - go.oy, w0, n0, w1, n1, ctx->d_err);
        else if (K <= 64)  // 8 lanes per new cell, 4 cells per warp
            k_rebin_g<8><<<(n + 31) / 32, 256, 0, st>>>(po, co, io, ctx->pos[nb], ctx->cnt[nb], in, gn, gn.ox - go.ox,
                                                         gn.oy - go.oy, w0, n0, w1, n1, ctx->d_err);
        else  // very populated cells: one warp per new cell (whole slab only)
            k_rebin<<<(ncell_owned + 7) / 8, 256, 0, st>>>(po, co, io, ctx->pos[nb], ctx->cnt[nb], in, go, gn, ctx->d_err);
        ctx->launches++;
        CU(cudaGetLastError());
        return MC2D_OK;
    };
    // single rank / no overlap: the four passes of a sweep go out as ONE cooperative launch (completion counters or a
    // grid barrier between colours); tuning.fuse_passes = 0 launches them one by one
    const bool fuse = persistent && !overlap && !g.ghost && ctx->tune.fuse_passes;
    const size_t sweep_smem = G ? (size_t)R * (sizeof(double2) + (ctx->use_ids ? sizeof(int) : 0)) * wpb : per_warp * wpb;
    // Slab mode, everything on ONE stream: the re-bin kernel itself waits for the old ghost columns, pushes the
    // re-binned boundary columns into the neighbours' ghost columns and signals; k_order_local waits for the
    // neighbours' signal before it reads the new ghost counts; the sweep kernel does the rest (halo_fused + pipeline).
    // Without the four stream joins and the three side-stream launches per sweep of the two-stream variant (which
    // stays for the grid-barrier / per-pass A-B variants and for shifts too small for k_rebin4).
    const bool slab_single = overlap && ctx->tune.fuse_passes && ctx->tune.pipeline && K <= 64 && rebin_G == 4 && g.nx >= 3 &&
                             g.ny >= 3;
    auto slab_halo = [&](int buf, unsigned long long wait_seq, unsigned long long signal_seq) {
        SlabHalo sh = no_slab_halo();
        const size_t colp = (size_t)g.ny * g.K, colc = (size_t)g.ny;
        const int ghost_col[2] = {g.ghost + g.ncol, 0};  // my first column lands in the LEFT neighbour's RIGHT ghost
        sh.on = 1;
        sh.myflag = reinterpret_cast<const unsigned long long *>(ctx->arena);
        sh.wait_seq = wait_seq;
        sh.signal_seq = signal_seq;
        for (int side = 0; side < 2; ++side) {
            char *pb = ctx->peer_base[side];
            sh.hpos[side] = reinterpret_cast<double2 *>(pb + ctx->off_pos[buf]) + (size_t)ghost_col[side] * colp;
            sh.hcnt[side] = reinterpret_cast<int *>(pb + ctx->off_cnt[buf]) + (size_t)ghost_col[side] * colc;
            sh.hids[side] = ctx->use_ids ? reinterpret_cast<int *>(pb + ctx->off_ids[buf]) + (size_t)ghost_col[side] * colp
                                         : nullptr;
            sh.hflag[side] = reinterpret_cast<unsigned long long *>(pb) + (side == 0 ? 1 : 0);
        }
        sh.bcount = ctx->d_work + 7;
        // k_rebin4 is launched with the last owned column first and the first owned column right behind it (see below):
        // the cells of the two slab-edge columns are the first 2 ny cells of its list, 8 cells per warp
        sh.nbwarps = (2 * g.ny + 7) / 8;
        sh.err = ctx->d_err;
        sh.spin_budget_ns = spin_budget_ns(ctx);
        return sh;
    };
    SlabHalo order_wait = no_slab_halo();
    for (long long s = 0; s < nsweeps; ++s) {
        int order[4];
        double sx, sy;
        mc2d_schedule(ctx->prm.seed, ctx->sweep, g.dx, g.dy, plan->shift_grid, order, &sx, &sy);
        order_wait.on = 0;
        bool used_sb = !slab_single;  // the two-stream variants queue work on the boundary stream in every sweep
        if ((sx != g.ox || sy != g.oy) && slab_single && fabs(sx - g.ox) > 1e-9 * g.dx && fabs(sy - g.oy) > 1e-9 * g.dy) {
            Grid gn = g;
            gn.ox = sx;
            gn.oy = sy;
            const int nb = 1 - ctx->cur;
            const SlabHalo sh = slab_halo(nb, ctx->halo_seq, ctx->halo_seq + 1);
            ++ctx->halo_seq;
            CU(cudaMemsetAsync(ctx->d_work + 7, 0, sizeof(int), sa));
            int rc = ktime_begin(ctx, 1);
            if (rc) return rc;
            k_rebin4<4, true><<<(ncell_owned + 63) / 64, 256, 0, sa>>>(
                ctx->pos[ctx->cur], ctx->cnt[ctx->cur], ctx->use_ids ? ctx->ids[ctx->cur] : nullptr, ctx->pos[nb], ctx->cnt[nb],
                ctx->use_ids ? ctx->ids[nb] : nullptr, gn, sx > g.ox ? 1 : -1, sy > g.oy ? 1 : -1,
                // slab-edge columns first: their pushes over NVLink and the signal to the neighbours (whose k_order_local
                // waits for it) travel while the interior columns are re-binned
                (g.ncol - 1) * g.ny, g.ny, 0, (g.ncol - 1) * g.ny, ctx->d_err, sh);
            ctx->launches++;
            CU(cudaGetLastError());
            if ((rc = ktime_end(ctx))) return rc;
            ctx->cur = nb;
            g = gn;
            ctx->order_valid = false;
            order_wait = sh;  // k_order_local waits for the neighbours' signal_seq
        } else if (sx != g.ox || sy != g.oy) {
            Grid gn = g;
            gn.ox = sx;
            gn.oy = sy;
            const int nb = 1 - ctx->cur;
            if (overlap) {
                used_sb = true;
                int rc = join();
                if (rc) return rc;
                // the boundary columns read both ghosts of the old buffer: wait for the neighbours' last pushes
                if ((rc = halo_exchange(ctx, ctx->cur, false, false, true, true, false, sb))) return rc;
                if ((rc = launch_rebin(sb, g, gn, nb, 0, g.ny, (g.ncol - 1) * g.ny, g.ny))) return rc;
                if ((rc = ktime_begin(ctx, 1))) return rc;
                if ((rc = launch_rebin(sa, g, gn, nb, g.ny, (g.ncol - 2) * g.ny, 0, 0))) return rc;
                if ((rc = ktime_end(ctx))) return rc;
                ctx->cur = nb;
                g = gn;
                ctx->order_valid = false;
                if ((rc = halo_exchange(ctx, ctx->cur, true, true, false, false, true, sb))) return rc;
                if ((rc = join())) return rc;
            } else {
                int rc = ktime_begin(ctx, 1);
                if (rc) return rc;
                if ((rc = launch_rebin(sa, g, gn, nb, 0, ncell_owned, 0, 0))) return rc;
                if ((rc = ktime_end(ctx))) return rc;
                ctx->cur = nb;
                g = gn;
                ctx->order_valid = false;
                const bool even_next = (order[0] & 1) == 0;  // the first pass reads the left ghost (even columns) or the right one
                if (g.ghost)
                    if ((rc = ktime_begin(ctx, 2))) return rc;
                rc = halo_exchange(ctx, ctx->cur, true, true, even_next, !even_next);
                if (rc) return rc;
                if (g.ghost)
                    if ((rc = ktime_end(ctx))) return rc;
            }
        }
        if (persistent) {
            if (overlap && used_sb)  // the boundary stream may still be using the counters and the order of the last sweep
                if (int rc = join()) return rc;
            // ticket counters; finished items per (colour pass, column pair) + 8 words of the slab pipeline (edge items per
            // pass, frontier): zero before the sweep kernel — by k_order_local when it runs anyway (every sweep with a grid
            // shift), by two memsets otherwise
            const int n_done = 4 * npairs + 8;
            CU(ctx->done_buf.ensure(sizeof(int) * (size_t)n_done));
            if (!ctx->order_valid) {
                if (int rc = build_order(ctx, order_wait.on ? &order_wait : nullptr, ctx->d_work, 8, ctx->done_buf.as<int>(), n_done))
                    return rc;
            } else {
                CU(cudaMemsetAsync(ctx->d_work, 0, 8 * sizeof(int), ctx->stream));
                CU(cudaMemsetAsync(ctx->done_buf.p, 0, sizeof(int) * (size_t)n_done, ctx->stream));
            }
        }
        // slab mode: the same single launch, with the halo exchange inside the kernel (boundary items wait on the
        // neighbour's flag word, push their cells into its ghost column, the last one signals)
        // (grid-barrier variant: its slab-edge items must all be first items of a warp; the pipeline has no such limit)
        const bool fuse_halo = overlap && ctx->tune.fuse_passes &&
                               (ctx->tune.pipeline || (nitems / npairs) <= ctx->num_sms * wpb);
        for (int p = 0; p < ((fuse || fuse_halo) ? 1 : 4); ++p) {
            SweepArgs a;
            a.pos = ctx->pos[ctx->cur];
            a.cnt = ctx->cnt[ctx->cur];
            a.ids = ctx->use_ids ? ctx->ids[ctx->cur] : nullptr;
            a.order = persistent ? ctx->order_buf.as<int>() : nullptr;
            a.work = ctx->d_work + p;
            a.nitems = nitems;
            a.pair_lo = 0;
            a.pair_n = npairs;
            a.npass = (fuse || fuse_halo) ? 4 : 1;
            for (int q = 0; q < 4; ++q) a.colours[q] = order[q];
            a.dU_stride = nblocks;
            a.pipeline = (fuse && ctx->tune.pipeline) ? 1 : 0;
            a.done = ctx->done_buf.as<int>();
            a.halo_fused = 0;
            a.myflag = nullptr;
            a.seq_base = 0;
            for (int q = 0; q < 2; ++q) {
                a.hpos[q] = nullptr;
                a.hcnt[q] = nullptr;
                a.hids[q] = nullptr;
                a.hflag[q] = nullptr;
            }
            a.g = g;
            a.pp = ctx->pp;
            a.r2cut = ctx->r2cut;
            a.beta = ctx->beta;
            a.zA = ctx->prm.activity * g.dx * g.dy;
            a.delta = ctx->prm.delta;
            a.colour = order[p];
            a.pass = p;
            a.sweep_lo = (uint32_t)ctx->sweep;
            a.sweep_hi = (uint32_t)(ctx->sweep >> 32);
            a.seed_lo = (uint32_t)ctx->prm.seed;
            a.seed_hi = (uint32_t)(ctx->prm.seed >> 32);
            a.disp_mult = plan->disp_per_particle;
            a.disp_cell = plan->disp_per_cell;
            a.n_gc = plan->gc_per_cell;
            a.p_ins = ctx->p_ins;
            a.p_del = ctx->p_del;
            a.partial_dU = ctx->partial.as<double>() + (size_t)p * nblocks;
            a.counters = ctx->d_counters;
            a.err = ctx->d_err;
            a.spin_budget_ns = spin_budget_ns(ctx);
            a.trace = trace ? ctx->trace_buf.as<mc2d_trial>() : nullptr;
            a.trace_n = ctx->d_trace_n;
            a.trace_cap = trace_cap;
            const size_t smem = sweep_smem;
            const int pot = ctx->prm.potential;
            auto launch_items = [&](const SweepArgs &aa, cudaStream_t st, int reserve) -> int {  // k_sweep_p over aa's column pairs
                int rc = MC2D_OK;
                DISPATCH_POT(pot, {
                    if (G == 2) {
                        if (trace) rc = launch_sweep_p<POT, 2, true>(ctx, aa, wpb, smem, R, st, reserve);
                        else rc = launch_sweep_p<POT, 2, false>(ctx, aa, wpb, smem, R, st, reserve);
                    } else if (G == 4) {
                        if (trace) rc = launch_sweep_p<POT, 4, true>(ctx, aa, wpb, smem, R, st, reserve);
                        else rc = launch_sweep_p<POT, 4, false>(ctx, aa, wpb, smem, R, st, reserve);
                    } else {
                        if (trace) rc = launch_sweep_p<POT, 8, true>(ctx, aa, wpb, smem, R, st, reserve);
                        else rc = launch_sweep_p<POT, 8, false>(ctx, aa, wpb, smem, R, st, reserve);
                    }
                });
                return rc;
            };
            int rc = MC2D_OK;
            if (fuse_halo) {
                if (used_sb && (rc = join())) return rc;  // the re-bin's boundary part and its push ran on the other stream
                const size_t colp = (size_t)g.ny * g.K, colc = (size_t)g.ny;
                const int ghost_col[2] = {g.ghost + g.ncol, 0};  // my first column lands in the LEFT neighbour's RIGHT ghost
                for (int side = 0; side < 2; ++side) {
                    char *pb = ctx->peer_base[side];
                    a.hpos[side] = reinterpret_cast<double2 *>(pb + ctx->off_pos[ctx->cur]) + (size_t)ghost_col[side] * colp;
                    a.hcnt[side] = reinterpret_cast<int *>(pb + ctx->off_cnt[ctx->cur]) + (size_t)ghost_col[side] * colc;
                    a.hids[side] = ctx->use_ids ? reinterpret_cast<int *>(pb + ctx->off_ids[ctx->cur]) + (size_t)ghost_col[side] * colp
                                                : nullptr;
                    a.hflag[side] = reinterpret_cast<unsigned long long *>(pb) + (side == 0 ? 1 : 0);
                }
                a.myflag = reinterpret_cast<unsigned long long *>(ctx->arena);
                a.seq_base = ctx->halo_seq;
                a.halo_fused = 1;
                a.pipeline = ctx->tune.pipeline ? 1 : 0;
                ctx->halo_seq += 4;  // one sync point per colour pass, signalled from inside the kernel
                if ((rc = ktime_begin(ctx, 3))) return rc;
                if ((rc = launch_items(a, sa, 0))) return rc;
                if ((rc = ktime_end(ctx))) return rc;
                continue;
            }
            if (overlap) {
                const bool left = (order[p] & 1) == 0;  // even columns active: the first owned column changes
                const int items_per_pair = nitems / npairs;
                if ((rc = join())) return rc;
                // boundary pair on stream B: wait for the ghost it reads, run, push, signal
                if ((rc = halo_exchange(ctx, ctx->cur, false, false, left, !left, false, sb))) return rc;
                SweepArgs ab = a;
                ab.work = ctx->d_work + 4 + p;
                ab.pair_lo = left ? 0 : npairs - 1;
                ab.pair_n = 1;
                ab.nitems = items_per_pair;
                if ((rc = launch_items(ab, sb, 0))) return rc;
                if ((rc = halo_exchange(ctx, ctx->cur, left, !left, false, false, true, sb))) return rc;
                // interior pairs on the main stream
                a.pair_lo = left ? 1 : 0;
                a.pair_n = npairs - 1;
                a.nitems = items_per_pair * (npairs - 1);
                if ((rc = ktime_begin(ctx, 0))) return rc;
                if ((rc = launch_items(a, sa, (items_per_pair + wpb - 1) / wpb))) return rc;
                if ((rc = ktime_end(ctx))) return rc;
                continue;
            }
            rc = ktime_begin(ctx, fuse ? 3 : 0);  // 3: one launch = the four passes of a sweep
            if (rc) return rc;
            if (persistent) {
                rc = launch_items(a, sa, 0);
            } else {
                DISPATCH_POT(pot, {
                    const int inst = POT * 4 + (minimg ? 2 : 0) + (trace ? 1 : 0);
                    if (trace) {
                        if (minimg) rc = launch_sweep_kernel<POT, true, true>(ctx, a, nblocks, wpb, smem, inst);
                        else rc = launch_sweep_kernel<POT, false, true>(ctx, a, nblocks, wpb, smem, inst);
                    } else {
                        if (minimg) rc = launch_sweep_kernel<POT, true, false>(ctx, a, nblocks, wpb, smem, inst);
                        else rc = launch_sweep_kernel<POT, false, false>(ctx, a, nblocks, wpb, smem, inst);
                    }
                });
            }
            if (rc) return rc;
            if ((rc = ktime_end(ctx))) return rc;
            if (g.ghost) {
                const bool left = (order[p] & 1) == 0;
                // what the next kernel on the stream reads: one ghost for a pass, both for re-bin / energy / download
                const bool even_next = p < 3 && (order[p + 1] & 1) == 0;
                if ((rc = ktime_begin(ctx, 2))) return rc;
                rc = halo_exchange(ctx, ctx->cur, left, !left, p == 3 || even_next, p == 3 || !even_next);
                if (rc) return rc;
                if ((rc = ktime_end(ctx))) return rc;
            }
        }
        ctx->sweep++;
    }
    if (overlap) {  // back to one stream; what follows (energy, download, the next call) reads both ghosts
        int rc = join();
        if (rc) return rc;
        if ((rc = halo_exchange(ctx, ctx->cur, false, false, true, true, false, sa))) return rc;
    }
    // running energy += the batch's energy changes; error words, counters and the energy land in the pinned scalars
    k_sweeps_finish<<<1, 1024, 0, ctx->stream>>>(ctx->partial.as<double>(), 4 * nblocks, ctx->d_energy, ctx->d_err, ctx->d_counters,
                                                 ctx->hp_dev->errv, ctx->hp_dev->counters, &ctx->hp_dev->energy);
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->hp->tn = 0;
    if (trace) CU(cudaMemcpyAsync(&ctx->hp->tn, ctx->d_trace_n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (with_stats)  // the caller's statistics ride on this synchronisation (k_sweeps_finish has sent the counters)
        if (int rc = enqueue_stats_fetch(ctx, false)) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    const int *errv = ctx->hp->errv;
    const unsigned long long tn = ctx->hp->tn;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_sweep_ms = ms;
    if (ctx->ktiming) {
        ctx->k_sweep_ms = ctx->k_rebin_ms = ctx->k_halo_ms = 0;
        ctx->k_sweep_launches = ctx->k_rebin_launches = ctx->k_halo_launches = 0;
        for (size_t i = 0; i < ctx->kev_kind.size(); ++i) {
            float t = 0;
            CU(cudaEventElapsedTime(&t, ctx->kev[2 * i], ctx->kev[2 * i + 1]));
            if (ctx->kev_kind[i] == 0 || ctx->kev_kind[i] == 3) {
                ctx->k_sweep_ms += t;
                ctx->k_sweep_launches += ctx->kev_kind[i] == 3 ? 4 : 1;  // counted in colour passes
            } else if (ctx->kev_kind[i] == 1) {
                ctx->k_rebin_ms += t;
                ctx->k_rebin_launches++;
            } else {
                ctx->k_halo_ms += t;
                ctx->k_halo_launches++;
            }
        }
    }
    if (trace) {
        long long ncopy = std::min<long long>((long long)tn, trace_cap);
        if (ncopy > 0) CU(cudaMemcpy(trace, ctx->trace_buf.p, sizeof(mc2d_trial) * (size_t)ncopy, cudaMemcpyDeviceToHost));
        if (trace_n) *trace_n = (long long)tn;
    }
    if (int rc = check_peer_timeout(ctx, errv)) return rc;
    if (errv[0] > K) {
        ctx->have_particles = false;
        FAIL(MC2D_ERR_CELL_OVERFLOW,
             "re-binning needed %d slots in one cell but cell_capacity is %d: particles were dropped, the "
             "configuration on the device is no longer valid; re-upload with a larger cell_capacity",
             errv[0], K);
    }
    ctx->last_refused = errv[5];
    return MC2D_OK;
}

// the counters, the running energy and the particle count on their way to the pinned scalars (no synchronisation)
static int enqueue_stats_fetch(mc2d_ctx *ctx, bool counters_too = true) {
    if (counters_too) {
        CU(cudaMemcpyAsync(ctx->hp->counters, ctx->d_counters, sizeof ctx->hp->counters, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(&ctx->hp->energy, ctx->d_energy, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return count_particles_async(ctx);
}

// ... and read after the stream has been synchronised
static void read_stats(mc2d_ctx *ctx, mc2d_stats *st) {
    const long long n = (long long)ctx->hp->count;
    ctx->n_last = n;
    const unsigned long long *c = ctx->hp->counters;
    const double e = ctx->hp->energy;
    for (int i = 0; i < 3; ++i) {
        st->attempted[i] = (long long)c[i];
        st->accepted[i] = (long long)c[3 + i];
    }
    st->overflow = (long long)c[6];
    st->n_particles = n;
    st->energy = e;
    st->sweeps_done = ctx->sweep;
}

static int fetch_stats(mc2d_ctx *ctx, mc2d_stats *st) {
    if (int rc = enqueue_stats_fetch(ctx)) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    read_stats(ctx, st);
    return MC2D_OK;
}

extern "C" int mc2d_run_sweeps(mc2d_ctx *ctx, long long nsweeps, const mc2d_sweep_plan *plan, mc2d_stats *st) {
    if (!ctx) return MC2D_ERR_INVALID;
    int rc = run_sweeps_impl(ctx, nsweeps, plan, nullptr, 0, nullptr, st != nullptr);
    if (rc) return rc;
    if (st) read_stats(ctx, st);
    // A refused insertion biases the ensemble (detailed balance): it is reported whether or not the caller asked for
    // the statistics, and per CALL (stats.overflow is the running total since mc2d_reset_counters).  The sweeps of this
    // call have run and the statistics are valid; the caller re-uploads with a larger cell_capacity.
    if (ctx->last_refused > 0)
        FAIL(MC2D_ERR_CELL_OVERFLOW, "%d insertion(s) of this call were refused because a cell had no free slot (capacity %d)",
             ctx->last_refused, ctx->g.K);
    return MC2D_OK;
}

extern "C" int mc2d_run_sweep_traced(mc2d_ctx *ctx, const mc2d_sweep_plan *plan, mc2d_trial *trials, long long cap,
                                     long long *n_out, mc2d_stats *st) {
    if (!ctx || !trials || cap <= 0) return MC2D_ERR_INVALID;
    int rc = run_sweeps_impl(ctx, 1, plan, trials, cap, n_out);
    if (rc) return rc;
    if (st) return fetch_stats(ctx, st);
    return MC2D_OK;
}

extern "C" int mc2d_get_stats(mc2d_ctx *ctx, mc2d_stats *st) {
    if (!ctx || !st) return MC2D_ERR_INVALID;
    if (!ctx->have_particles) FAIL(MC2D_ERR_STATE, "no particles uploaded");
    CU(cudaSetDevice(ctx->dev));
    return fetch_stats(ctx, st);
}

extern "C" int mc2d_reset_counters(mc2d_ctx *ctx) {
    if (!ctx) return MC2D_ERR_INVALID;
    CU(cudaSetDevice(ctx->dev));
    CU(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    return MC2D_OK;
}

extern "C" int mc2d_set_delta(mc2d_ctx *ctx, double delta) {
    if (!ctx || !(delta >= 0)) return MC2D_ERR_INVALID;
    ctx->prm.delta = delta;
    return MC2D_OK;
}

extern "C" int mc2d_set_activity(mc2d_ctx *ctx, double z) {
    if (!ctx || !(z >= 0)) return MC2D_ERR_INVALID;
    ctx->prm.activity = z;
    return MC2D_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU
// ------------------------------------------------------------------------------------------------
extern "C" int mc2d_comm_unique_id(void *id_out) {
#if MC2D_HAVE_NCCL_H
    static_assert(sizeof(ncclUniqueId) == MC2D_UNIQUE_ID_BYTES, "ncclUniqueId size");
    std::string err;
    if (!id_out || !g_nccl.load(err)) return MC2D_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MC2D_ERR_NCCL;
    memcpy(id_out, &id, sizeof id);
    return MC2D_OK;
#else
    (void)id_out;
    return MC2D_ERR_NCCL;
#endif
}

extern "C" int mc2d_comm_init(mc2d_ctx *ctx, const void *unique_id) {
    if (!ctx || !unique_id) return MC2D_ERR_INVALID;
#if MC2D_HAVE_NCCL_H
    if (!g_nccl.load(ctx->err)) return MC2D_ERR_NCCL;
    CU(cudaSetDevice(ctx->dev));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, ctx->prm.nranks, id, ctx->prm.rank);
    if (r != ncclSuccess) FAIL(MC2D_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    ctx->have_comm = true;
    if (ctx->have_particles && ctx->g.ghost) {  // particles were uploaded before the communicator existed
        int rc = halo_exchange(ctx, ctx->cur, true, true);
        if (rc) return rc;
        double U, W;
        long long N;
        rc = energy_on_slots(ctx, &U, &W, &N, true, false);
        if (rc) return rc;
    }
    return MC2D_OK;
#else
    FAIL(MC2D_ERR_NCCL, "built without nccl.h");
#endif
}

extern "C" int mc2d_allreduce_stats(mc2d_ctx *ctx, mc2d_stats *st) {
    if (!ctx || !st) return MC2D_ERR_INVALID;
    if (ctx->prm.nranks == 1) return MC2D_OK;
#if MC2D_HAVE_NCCL_H
    if (!ctx->have_comm) FAIL(MC2D_ERR_STATE, "no communicator");
    CU(cudaSetDevice(ctx->dev));
    double h[9];
    for (int i = 0; i < 3; ++i) {
        h[i] = (double)st->attempted[i];
        h[3 + i] = (double)st->accepted[i];
    }
    h[6] = (double)st->n_particles;
    h[7] = st->energy;
    h[8] = (double)st->overflow;
    CU(ctx->outE.ensure(sizeof h));
    CU(cudaMemcpyAsync(ctx->outE.p, h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = g_nccl.AllReduce(ctx->outE.p, ctx->outE.p, 9, ncclDouble, ncclSum, ctx->comm, ctx->stream);
    if (r != ncclSuccess) FAIL(MC2D_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    CU(cudaMemcpyAsync(h, ctx->outE.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) {
        st->attempted[i] = llround(h[i]);
        st->accepted[i] = llround(h[3 + i]);
    }
    st->n_particles = llround(h[6]);
    st->energy = h[7];
    st->overflow = llround(h[8]);
    return MC2D_OK;
#else
    FAIL(MC2D_ERR_NCCL, "built without nccl.h");
#endif
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------
extern "C" int mc2d_measure_fp64_peak(mc2d_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return MC2D_ERR_INVALID;
    CU(cudaSetDevice(ctx->dev));
    const int iters = 8192, blocks = ctx->num_sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(ctx->d_probe, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    ctx->launches += 4;
    *tflops = best;
    return MC2D_OK;
}

extern "C" int mc2d_get_info(mc2d_ctx *ctx, mc2d_info *o) {
    if (!ctx || !o) return MC2D_ERR_INVALID;
    const Grid &g = ctx->g;
    o->nx = g.nx;
    o->ny = g.ny;
    o->col0 = g.col0;
    o->ncol = g.ncol;
    o->cell_capacity = g.K;
    o->cell_dx = g.dx;
    o->cell_dy = g.dy;
    o->launches = ctx->launches;
    o->last_sweep_ms = ctx->last_sweep_ms;
    o->last_energy_ms = ctx->last_energy_ms;
    o->last_pair_tests = ctx->last_tests;
    o->last_pair_evals = ctx->last_evals;
    o->k_sweep_ms = ctx->k_sweep_ms;
    o->k_sweep_launches = ctx->k_sweep_launches;
    o->k_rebin_ms = ctx->k_rebin_ms;
    o->k_rebin_launches = ctx->k_rebin_launches;
    o->halo_mode = !ctx->g.ghost ? 0 : (ctx->p2p ? 2 : 1);
    o->k_halo_ms = ctx->k_halo_ms;
    o->k_halo_launches = ctx->k_halo_launches;
    o->slot_allocs = ctx->slot_allocs;
    o->sweep_kernel_lanes = ctx->last_G ? ctx->last_G : 32;
    return MC2D_OK;
}

extern "C" int mc2d_set_kernel_timing(mc2d_ctx *ctx, int on) {
    if (!ctx) return MC2D_ERR_INVALID;
    ctx->ktiming = on != 0;
    return MC2D_OK;
}

extern "C" int mc2d_get_stream(mc2d_ctx *ctx, void **s) {
    if (!ctx || !s) return MC2D_ERR_INVALID;
    *s = (void *)ctx->stream;
    return MC2D_OK;
}

extern "C" int mc2d_synchronize(mc2d_ctx *ctx) {
    if (!ctx) return MC2D_ERR_INVALID;
    CU(cudaSetDevice(ctx->dev));
    CU(cudaStreamSynchronize(ctx->stream));
    return MC2D_OK;
}
