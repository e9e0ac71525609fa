// mpi_shim.cpp — TEST INFRASTRUCTURE: implementation of oracle/mpi_shim/mpi.h (threads as ranks) and the
// launcher of the reference's MPI driver.
//
//   mc2d_ref_mpi <nranks> <input.txt>
//
// starts <nranks> threads; each calls the reference's own main() (mpi_src/main.cpp, compiled UNMODIFIED
// with -Dmain=mc2d_ref_mpi_main, see oracle/Makefile) with argv = {prog, input.txt}.  Rank 0's thread
// reports the wall time of the whole run and the number of attempted moves the reference performs
// (MC_parallel.cpp:57-114: every sweep makes owned.size() "attempts" of 4 parity micro-steps per rank,
// each micro-step one displacement trial => 4 N trials per sweep):
//
//   MC2D_MPI_TIMING ranks=R N=.. sweeps=.. run_s=.. attempted=..
#include "mpi.h"

#include <fcntl.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

struct mc2d_mpi_request {
    bool is_recv = false;
    void *buf = nullptr;
    size_t bytes = 0;
    int source = 0, tag = 0;
};
struct mc2d_mpi_file {
    int fd = -1;
};

namespace {

int g_size = 1;
thread_local int tl_rank = 0;

// ---- datatypes ---------------------------------------------------------------------------------
std::mutex g_type_mutex;
std::vector<size_t> g_type_size = {0, 1, sizeof(int), sizeof(double), sizeof(long long), 8, 8, sizeof(unsigned)};
size_t type_size(MPI_Datatype dt) {
    std::lock_guard<std::mutex> lk(g_type_mutex);
    if (dt <= 0 || (size_t)dt >= g_type_size.size()) {
        fprintf(stderr, "mpi shim: unknown datatype %d\n", dt);
        abort();
    }
    return g_type_size[(size_t)dt];
}

inline void relax(unsigned &spins) {
    if (++spins < 2000) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    } else {
        sched_yield();
    }
}

// ---- barrier (sense reversing, spinning) -----------------------------------------------------------
std::atomic<int> g_bar_count{0};
std::atomic<int> g_bar_gen{0};
void barrier() {
    const int gen = g_bar_gen.load(std::memory_order_acquire);
    if (g_bar_count.fetch_add(1, std::memory_order_acq_rel) == g_size - 1) {
        g_bar_count.store(0, std::memory_order_relaxed);
        g_bar_gen.store(gen + 1, std::memory_order_release);
    } else {
        unsigned spins = 0;
        while (g_bar_gen.load(std::memory_order_acquire) == gen) relax(spins);
    }
}

// ---- collectives: every rank publishes its arguments, all meet, each copies what it needs, all meet ----
struct CollSlot {
    const void *send = nullptr;
    void *recv = nullptr;
    const int *scounts = nullptr, *sdispls = nullptr;
    int count = 0;
};
std::vector<CollSlot> g_coll;

// ---- point to point: eager copies, FIFO per (source, tag) at the destination --------------------------
struct Mailbox {
    std::mutex m;
    std::map<std::pair<int, int>, std::deque<std::vector<char>>> q;  // (source, tag) -> messages in send order
    std::atomic<long> arrivals{0};
};
std::vector<Mailbox> *g_mail = nullptr;

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) {
    fflush(nullptr);
    _exit(code);
}
int MPI_Comm_rank(MPI_Comm, int *rank) {
    *rank = tl_rank;
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm, int *size) {
    *size = g_size;
    return MPI_SUCCESS;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm, MPI_Request *req) {
    const size_t bytes = (size_t)count * type_size(dt);
    std::vector<char> msg(bytes);
    if (bytes) memcpy(msg.data(), buf, bytes);
    Mailbox &mb = (*g_mail)[(size_t)dest];
    {
        std::lock_guard<std::mutex> lk(mb.m);
        mb.q[{tl_rank, tag}].push_back(std::move(msg));
    }
    mb.arrivals.fetch_add(1, std::memory_order_release);
    *req = new mc2d_mpi_request();  // complete on return (eager)
    return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm, MPI_Request *req) {
    mc2d_mpi_request *r = new mc2d_mpi_request();
    r->is_recv = true;
    r->buf = buf;
    r->bytes = (size_t)count * type_size(dt);
    r->source = source;
    r->tag = tag;
    *req = r;
    return MPI_SUCCESS;
}

// Receives complete in array order, which is the posting order at every call site of the reference
// (particle_exchange.cpp:151-167): the non-overtaking rule per (source, tag) is kept.
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *) {
    Mailbox &mb = (*g_mail)[(size_t)tl_rank];
    for (int i = 0; i < n; ++i) {
        mc2d_mpi_request *r = reqs[i];
        if (r && r->is_recv) {
            unsigned spins = 0;
            for (;;) {
                bool got = false;
                {
                    std::lock_guard<std::mutex> lk(mb.m);
                    auto it = mb.q.find({r->source, r->tag});
                    if (it != mb.q.end() && !it->second.empty()) {
                        std::vector<char> &msg = it->second.front();
                        if (msg.size() > r->bytes) {
                            fprintf(stderr, "mpi shim: message of %zu bytes truncated to %zu\n", msg.size(), r->bytes);
                            abort();
                        }
                        if (!msg.empty()) memcpy(r->buf, msg.data(), msg.size());
                        it->second.pop_front();
                        got = true;
                    }
                }
                if (got) break;
                relax(spins);
            }
        }
        delete r;
        reqs[i] = nullptr;
    }
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm) {
    g_coll[(size_t)tl_rank].send = buf;
    barrier();
    if (tl_rank != root) memcpy(buf, g_coll[(size_t)root].send, (size_t)count * type_size(dt));
    barrier();
    return MPI_SUCCESS;
}

static void reduce_into(void *acc, const void *src, int count, MPI_Datatype dt) {
    switch (dt) {
    case MPI_INT:
        for (int i = 0; i < count; ++i) ((int *)acc)[i] += ((const int *)src)[i];
        break;
    case MPI_DOUBLE:
        for (int i = 0; i < count; ++i) ((double *)acc)[i] += ((const double *)src)[i];
        break;
    case MPI_LONG_LONG_INT:
    case MPI_INT64_T:
        for (int i = 0; i < count; ++i) ((long long *)acc)[i] += ((const long long *)src)[i];
        break;
    case MPI_UINT64_T:
        for (int i = 0; i < count; ++i) ((unsigned long long *)acc)[i] += ((const unsigned long long *)src)[i];
        break;
    default:
        fprintf(stderr, "mpi shim: reduction on datatype %d not supported\n", dt);
        abort();
    }
}

int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
    g_coll[(size_t)tl_rank].send = send;
    barrier();
    const size_t bytes = (size_t)count * type_size(dt);
    std::vector<char> acc(bytes, 0);
    for (int r = 0; r < g_size; ++r) reduce_into(acc.data(), g_coll[(size_t)r].send, count, dt);  // rank order: same sum everywhere
    barrier();  // everyone has read the send buffers (recv may alias a neighbour's send)
    memcpy(recv, acc.data(), bytes);
    return MPI_SUCCESS;
}

int MPI_Exscan(const void *send, void *recv, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
    g_coll[(size_t)tl_rank].send = send;
    barrier();
    const size_t bytes = (size_t)count * type_size(dt);
    std::vector<char> acc(bytes, 0);
    for (int r = 0; r < tl_rank; ++r) reduce_into(acc.data(), g_coll[(size_t)r].send, count, dt);
    barrier();
    if (tl_rank > 0) memcpy(recv, acc.data(), bytes);  // rank 0's result is undefined in MPI: left untouched
    return MPI_SUCCESS;
}

int MPI_Gatherv(const void *send, int scount, MPI_Datatype sdt, void *recv, const int *rcounts, const int *displs,
                MPI_Datatype rdt, int root, MPI_Comm) {
    g_coll[(size_t)tl_rank].send = send;
    g_coll[(size_t)tl_rank].count = scount;
    barrier();
    if (tl_rank == root) {
        const size_t rs = type_size(rdt), ss = type_size(sdt);
        for (int r = 0; r < g_size; ++r) {
            const size_t bytes = (size_t)g_coll[(size_t)r].count * ss;
            if ((size_t)rcounts[r] * rs < bytes) {
                fprintf(stderr, "mpi shim: gatherv truncation\n");
                abort();
            }
            if (bytes) memcpy((char *)recv + (size_t)displs[r] * rs, g_coll[(size_t)r].send, bytes);
        }
    }
    barrier();
    return MPI_SUCCESS;
}

int MPI_Gather(const void *send, int scount, MPI_Datatype sdt, void *recv, int rcount, MPI_Datatype rdt, int root,
               MPI_Comm comm) {
    std::vector<int> counts((size_t)g_size, rcount), displs((size_t)g_size);
    for (int r = 0; r < g_size; ++r) displs[(size_t)r] = r * rcount;
    return MPI_Gatherv(send, scount, sdt, recv, counts.data(), displs.data(), rdt, root, comm);
}

int MPI_Alltoallv(const void *send, const int *scounts, const int *sdispls, MPI_Datatype sdt, void *recv,
                  const int *rcounts, const int *rdispls, MPI_Datatype rdt, MPI_Comm) {
    CollSlot &me = g_coll[(size_t)tl_rank];
    me.send = send;
    me.scounts = scounts;
    me.sdispls = sdispls;
    barrier();
    const size_t ss = type_size(sdt), rs = type_size(rdt);
    for (int r = 0; r < g_size; ++r) {
        const CollSlot &o = g_coll[(size_t)r];
        const size_t bytes = (size_t)o.scounts[tl_rank] * ss;
        if ((size_t)rcounts[r] * rs < bytes) {
            fprintf(stderr, "mpi shim: alltoallv truncation\n");
            abort();
        }
        if (bytes) memcpy((char *)recv + (size_t)rdispls[r] * rs, (const char *)o.send + (size_t)o.sdispls[tl_rank] * ss, bytes);
    }
    barrier();
    return MPI_SUCCESS;
}

int MPI_Alltoall(const void *send, int scount, MPI_Datatype sdt, void *recv, int rcount, MPI_Datatype rdt,
                 MPI_Comm comm) {
    std::vector<int> sc((size_t)g_size, scount), sd((size_t)g_size), rc((size_t)g_size, rcount), rd((size_t)g_size);
    for (int r = 0; r < g_size; ++r) {
        sd[(size_t)r] = r * scount;
        rd[(size_t)r] = r * rcount;
    }
    return MPI_Alltoallv(send, sc.data(), sd.data(), sdt, recv, rc.data(), rd.data(), rdt, comm);
}

int MPI_Type_match_size(int, int size, MPI_Datatype *dt) {
    *dt = size == 8 ? MPI_INT64_T : (size == 4 ? MPI_INT : MPI_CHAR);
    return MPI_SUCCESS;
}
int MPI_Get_address(const void *location, MPI_Aint *address) {
    *address = (MPI_Aint)(intptr_t)location;
    return MPI_SUCCESS;
}
// extent = end of the last member rounded up to 8 bytes (the reference's one struct, PackedParticle
// {double, double, int64}, particle_exchange.cpp:10-32, has no holes: 24 bytes)
int MPI_Type_create_struct(int count, const int *blocklens, const MPI_Aint *displs, const MPI_Datatype *types,
                           MPI_Datatype *newtype) {
    size_t end = 0;
    for (int i = 0; i < count; ++i) {
        const size_t e = (size_t)displs[i] + (size_t)blocklens[i] * type_size(types[i]);
        if (e > end) end = e;
    }
    end = (end + 7) / 8 * 8;
    std::lock_guard<std::mutex> lk(g_type_mutex);
    g_type_size.push_back(end);
    *newtype = (MPI_Datatype)(g_type_size.size() - 1);
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }

int MPI_File_open(MPI_Comm, const char *name, int amode, MPI_Info, MPI_File *fh) {
    int flags = O_WRONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    const int fd = open(name, flags, 0644);
    if (fd < 0) return 1;
    mc2d_mpi_file *f = new mc2d_mpi_file();
    f->fd = fd;
    *fh = f;
    barrier();  // collective: the file exists for every rank afterwards
    return MPI_SUCCESS;
}
int MPI_File_close(MPI_File *fh) {
    if (fh && *fh) {
        close((*fh)->fd);
        delete *fh;
        *fh = MPI_FILE_NULL;
    }
    return MPI_SUCCESS;
}
int MPI_File_get_size(MPI_File fh, MPI_Offset *size) {
    struct stat st;
    if (fstat(fh->fd, &st) != 0) return 1;
    *size = (MPI_Offset)st.st_size;
    return MPI_SUCCESS;
}
int MPI_File_write_at_all(MPI_File fh, MPI_Offset offset, const void *buf, int count, MPI_Datatype dt, MPI_Status *) {
    const size_t bytes = (size_t)count * type_size(dt);
    size_t done = 0;
    while (done < bytes) {
        const ssize_t w = pwrite(fh->fd, (const char *)buf + done, bytes - done, (off_t)(offset + (MPI_Offset)done));
        if (w <= 0) return 1;
        done += (size_t)w;
    }
    return MPI_SUCCESS;
}

}  // extern "C"

// ---- launcher -------------------------------------------------------------------------------------------------
int mc2d_ref_mpi_main(int argc, char *argv[]);  // = main() of /root/reference/mpi_src/main.cpp (-Dmain=mc2d_ref_mpi_main)

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <nranks> <input.txt>\n", argv[0]);
        return 1;
    }
    g_size = atoi(argv[1]);
    if (g_size < 1 || g_size > 256) {
        fprintf(stderr, "nranks out of range\n");
        return 1;
    }
    std::vector<Mailbox> mail((size_t)g_size);
    g_mail = &mail;
    g_coll.assign((size_t)g_size, CollSlot());
    // what the run will do, read with the reference's own "key value" format (input.cpp:12-35)
    double N = 0, nSteps = 0, eSteps = 0;
    {
        std::ifstream in(argv[2]);
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream iss(line);
            std::string key;
            double v;
            if (iss >> key >> v) {
                if (key == "N") N = v;
                if (key == "nSteps") nSteps = v;
                if (key == "eSteps") eSteps = v;
            }
        }
    }
    std::vector<int> rcs((size_t)g_size, 0);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < g_size; ++r)
        th.emplace_back([&, r]() {
            tl_rank = r;
            char *av[3] = {argv[0], argv[2], nullptr};
            rcs[(size_t)r] = mc2d_ref_mpi_main(2, av);
        });
    for (auto &t : th) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int rc = 0;
    for (int v : rcs) rc = rc ? rc : v;
    const long long sweeps = (long long)(nSteps + eSteps);
    printf("MC2D_MPI_TIMING ranks=%d N=%lld sweeps=%lld run_s=%.6f attempted=%lld\n", g_size, (long long)N, sweeps, secs,
           4LL * (long long)N * sweeps);
    return rc;
}
