// slab_ranks.hpp — the handful of rank-to-rank services the slab driver needs OUTSIDE the GPU library.
//
// The reference starts its ranks under mpirun and uses MPI for everything (mpi_src/main.cpp:28-50).  Here the GPU
// library moves all particle data itself (NVLink peer memory / NCCL); the host side only has to pass the 128-byte
// communicator id around, keep the ranks in step, and put the trajectory files together.  One forked process per
// GPU shares ONE anonymous page (`Shared`) for that: a sense-reversing barrier, a 256-byte broadcast buffer, a
// turn counter and one 64-bit word per rank.  Bulk data for the gather-to-root file mode travels through a
// scratch file under /dev/shm (`gatherv`), written by every rank at its own offset.
#ifndef MC2D_SLAB_RANKS_HPP
#define MC2D_SLAB_RANKS_HPP

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mc2d_slabs {

constexpr int kMaxRanks = 64;

// one page shared by the ranks (mmap MAP_SHARED | MAP_ANONYMOUS before the fork)
struct Shared {
    std::atomic<int> arrived;
    std::atomic<int> generation;
    std::atomic<int> turn;
    std::atomic<int> failed;
    unsigned char bcast_buf[256];
    long long counts[kMaxRanks];
    long long session;  // parent pid: names the scratch files of this run
};

struct Ranks {
    Shared *sh;
    int rank, size;

    void barrier() const {
        const int gen = sh->generation.load();
        if (sh->arrived.fetch_add(1) + 1 == size) {
            sh->arrived.store(0);
            sh->generation.fetch_add(1);
        } else {
            while (sh->generation.load() == gen) {
                if (sh->failed.load()) throw std::runtime_error("another rank failed");
                usleep(50);
            }
        }
    }
    void bcast(void *buf, std::size_t bytes, int root) const {
        if (bytes > sizeof sh->bcast_buf) throw std::runtime_error("bcast buffer too small");
        if (rank == root) std::memcpy(sh->bcast_buf, buf, bytes);
        barrier();
        if (rank != root) std::memcpy(buf, sh->bcast_buf, bytes);
        barrier();
    }
    // ranks run f one after the other in rank order (append to one file)
    template <class F>
    void in_turn(F &&f) const {
        barrier();
        while (sh->turn.load() != rank) {
            if (sh->failed.load()) throw std::runtime_error("another rank failed");
            usleep(50);
        }
        f();
        sh->turn.store(rank + 1 == size ? 0 : rank + 1);
        barrier();
    }
    // every rank contributes one number and gets all of them (MPI_Allgather of one long long)
    std::vector<long long> allgather(long long mine) const {
        sh->counts[rank] = mine;
        barrier();
        std::vector<long long> all(sh->counts, sh->counts + size);
        barrier();  // nobody overwrites its word before everybody has read
        return all;
    }
    // MPI_Gatherv of raw bytes to `root`: returns the concatenation in rank order there (empty elsewhere) and
    // the per-rank byte counts everywhere.
    std::vector<unsigned char> gatherv(const void *mine, std::size_t bytes, int root, std::vector<long long> *counts_out) const {
        const std::vector<long long> counts = allgather((long long)bytes);
        if (counts_out) *counts_out = counts;
        long long total = 0, my_off = 0;
        for (int r = 0; r < size; ++r) {
            if (r < rank) my_off += counts[r];
            total += counts[r];
        }
        std::vector<unsigned char> out;
        if (total == 0) return out;
        const std::string path = scratch_path();
        if (rank == root) {  // the root creates the file; the others open it after the barrier
            const int fd = ::open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0600);
            if (fd < 0) throw std::runtime_error("cannot create gather scratch file " + path);
            ::close(fd);
        }
        barrier();
        const int fd = ::open(path.c_str(), O_RDWR);
        if (fd < 0) throw std::runtime_error("cannot open gather scratch file " + path);
        write_all(fd, mine, bytes, (off_t)my_off, path);
        barrier();
        if (rank == root) {
            out.resize((std::size_t)total);
            std::size_t got = 0;
            while (got < out.size()) {
                const ssize_t n = ::pread(fd, out.data() + got, out.size() - got, (off_t)got);
                if (n <= 0) {
                    ::close(fd);
                    throw std::runtime_error("short read from gather scratch file " + path);
                }
                got += (std::size_t)n;
            }
            ::unlink(path.c_str());
        }
        ::close(fd);
        barrier();
        return out;
    }
    static void write_all(int fd, const void *buf, std::size_t bytes, off_t at, const std::string &what) {
        const unsigned char *p = static_cast<const unsigned char *>(buf);
        std::size_t done = 0;
        while (done < bytes) {
            const ssize_t n = ::pwrite(fd, p + done, bytes - done, at + (off_t)done);
            if (n <= 0) throw std::runtime_error("short write to " + what);
            done += (std::size_t)n;
        }
    }

private:
    std::string scratch_path() const {
        struct stat st;
        const char *dir = (::stat("/dev/shm", &st) == 0 && S_ISDIR(st.st_mode)) ? "/dev/shm" : "/tmp";
        return std::string(dir) + "/mc2d_slabs_gather_" + std::to_string(sh->session);
    }
};

}  // namespace mc2d_slabs

#endif
