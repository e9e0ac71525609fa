// slab_traj_logger.hpp — LoggingTrajSlabs: the reference's LoggingTrajMPI (mpi_src/logging_traj_mpi.h:11-90,
// logging_traj_mpi.cpp:98-337) for the forked GPU-slab ranks, byte-compatible file formats.
//
// Same interface: LoggingTrajSlabs(basename, mode, ranks, append, rank_as_type); log_xyz(local, box);
// log_dump(local, box, timestep); close().  The three modes and what they write (the formats are the contract:
// tests/test_traj_logger.py compares every mode byte for byte with frames written by the unmodified reference
// class on the threads-as-ranks MPI shim):
//
//   PerRank     <base>.rankNNNN.xyz / .dump, one file per rank, local counts, ids restart at 1 on every rank,
//               coordinates at the stream's default 6 significant digits (logging_traj_mpi.cpp:117-123, 189-206)
//   GatherRoot  <base>.xyz / .dump written by rank 0 from everybody's coordinates; xyz at 10 digits (the header's
//               setprecision sticks to the stream), dump at 6 (logging_traj_mpi.cpp:125-160, 208-266)
//   MPIIO       <base>.xyz / .dump, ONE shared file: rank 0 writes the header, every rank its own byte range at
//               header + (sum of the lower ranks' body sizes); xyz bodies at 6 digits, dump bodies at 17 with
//               global 1-based ids (logging_traj_mpi.cpp:162-178, 268-337).  Here the ranges are written with
//               pwrite() by the forked ranks — the same layout MPI_File_write_at_all produces.
//
// One deliberate difference: a fresh (non-append) MPIIO file is truncated when it is opened.  The reference opens
// it MPI_MODE_WRONLY | MPI_MODE_CREATE and starts at offset 0 (logging_traj_mpi.cpp:77-84), which leaves the tail
// of a longer file from an earlier run behind the new frames.
#ifndef MC2D_SLAB_TRAJ_LOGGER_HPP
#define MC2D_SLAB_TRAJ_LOGGER_HPP

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "mc2d_host.hpp"
#include "slab_ranks.hpp"

namespace mc2d_slabs {

class LoggingTrajSlabs {
public:
    enum class Mode { PerRank, GatherRoot, MPIIO };

    LoggingTrajSlabs(std::string basename, Mode mode, const Ranks &ranks, bool append = false, bool rank_as_type = false)
        : base_(std::move(basename)), mode_(mode), R_(ranks), append_(append), rank_as_type_(rank_as_type) {}
    ~LoggingTrajSlabs() { close(); }
    LoggingTrajSlabs(const LoggingTrajSlabs &) = delete;
    LoggingTrajSlabs &operator=(const LoggingTrajSlabs &) = delete;

    // logging_traj_mpi.cpp:108-114
    void log_xyz(const std::vector<Particle> &local, const SimulationBox &box) { write(xyz_, Kind::Xyz, local, box, 0); }
    // logging_traj_mpi.cpp:181-187
    void log_dump(const std::vector<Particle> &local, const SimulationBox &box, int timestep) {
        write(dump_, Kind::Dump, local, box, timestep);
    }
    void close() {
        xyz_.shut();
        dump_.shut();
    }

    // <base>.rankNNNN<ext> (logging_traj_mpi.cpp:37-41)
    static std::string rank_filename(const std::string &base, int rank, const char *ext) {
        char num[16];
        std::snprintf(num, sizeof num, "%04d", rank);
        return base + ".rank" + num + ext;
    }

private:
    enum class Kind { Xyz, Dump };
    // one output of one kind: a text stream (PerRank, GatherRoot) or a shared file written by byte range (MPIIO)
    struct Sink {
        std::ofstream text;
        int fd = -1;
        long long offset = 0;  // MPIIO: where the next frame starts
        bool open = false;
        void shut() {
            if (text.is_open()) text.close();
            if (fd >= 0) ::close(fd);
            fd = -1;
            open = false;
            offset = 0;
        }
    };

    std::string base_;
    Mode mode_;
    Ranks R_;
    bool append_, rank_as_type_;
    Sink xyz_, dump_;

    static const char *ext_of(Kind k) { return k == Kind::Xyz ? ".xyz" : ".dump"; }
    // basename + ext unless it ends in ext already (logging_traj_mpi.cpp:8-11)
    static std::string with_ext(const std::string &base, const char *ext) {
        const std::string e(ext);
        if (base.size() >= e.size() && base.compare(base.size() - e.size(), e.size(), e) == 0) return base;
        return base + e;
    }

    void open_text(Sink &s, const std::string &name, const char *what) {
        if (s.open) return;
        s.text.open(name, append_ ? (std::ios::out | std::ios::app) : (std::ios::out | std::ios::trunc));
        if (!s.text) throw std::runtime_error(std::string("Cannot open ") + what + ": " + name);
        s.open = true;
    }
    void open_shared(Sink &s, const std::string &name, const char *what) {
        if (s.open) return;
        // rank 0 creates (and, for a fresh file, empties) it; everybody opens it afterwards
        if (R_.rank == 0) {
            const int fd = ::open(name.c_str(), O_WRONLY | O_CREAT | (append_ ? 0 : O_TRUNC), 0644);
            if (fd < 0) {
                R_.sh->failed = 1;
                throw std::runtime_error(std::string("Cannot open shared ") + what + ": " + name);
            }
            ::close(fd);
        }
        R_.barrier();
        s.fd = ::open(name.c_str(), O_WRONLY);
        if (s.fd < 0) throw std::runtime_error(std::string("Cannot open shared ") + what + ": " + name);
        if (append_) {
            struct stat st;
            if (::fstat(s.fd, &st) == 0) s.offset = (long long)st.st_size;
        }
        s.open = true;
        R_.barrier();  // nobody writes before everybody has seen the size
    }

    // ---- frame pieces ------------------------------------------------------------------------------------------
    static void xyz_header(std::ostream &o, long long n, const SimulationBox &box) {
        o << n << "\n" << std::setprecision(10) << "Lx=" << box.getLx() << " Ly=" << box.getLy() << "\n";
    }
    void dump_header(std::ostream &o, long long n, const SimulationBox &box, int ts) const {
        o << "ITEM: TIMESTEP\n" << ts << "\nITEM: NUMBER OF ATOMS\n" << n << "\nITEM: BOX BOUNDS pp pp ff\n"
          << 0.0 << " " << box.getLx() << "\n" << 0.0 << " " << box.getLy() << "\n" << 0.0 << " " << 0.0 << "\n"
          << "ITEM: ATOMS " << (rank_as_type_ ? "id type x y z\n" : "id x y z\n");
    }
    // "id [type] x y z" rows for the particles of one rank; ids continue from first_id
    // (z is the double 0.0 in two of the reference's modes and the character 0 in the third: "0" either way)
    void dump_rows(std::ostream &o, const double *xy, long long n, long long first_id, int rank) const {
        for (long long i = 0; i < n; ++i) {
            o << first_id + i << " ";
            if (rank_as_type_) o << rank + 1 << " ";
            o << xy[2 * i] << " " << xy[2 * i + 1] << " 0\n";
        }
    }
    static void xyz_rows(std::ostream &o, const double *xy, long long n) {
        for (long long i = 0; i < n; ++i) o << "1 " << xy[2 * i] << " " << xy[2 * i + 1] << "\n";
    }

    void write(Sink &s, Kind kind, const std::vector<Particle> &local, const SimulationBox &box, int ts) {
        const double *xy = local.empty() ? nullptr : &local[0].x;  // Particle is a packed {x, y} (mc2d_host.hpp)
        const long long nloc = (long long)local.size();
        switch (mode_) {
        case Mode::PerRank: {
            open_text(s, rank_filename(base_, R_.rank, ext_of(kind)), kind == Kind::Xyz ? "per-rank XYZ" : "per-rank DUMP");
            if (kind == Kind::Xyz) {
                s.text << nloc << "\nrank=" << R_.rank << "\n";
                xyz_rows(s.text, xy, nloc);
            } else {
                dump_header(s.text, nloc, box, ts);
                dump_rows(s.text, xy, nloc, 1, R_.rank);
            }
            s.text.flush();
            break;
        }
        case Mode::GatherRoot: {
            if (R_.rank == 0) open_text(s, with_ext(base_, ext_of(kind)), kind == Kind::Xyz ? "gather XYZ" : "gather DUMP");
            std::vector<long long> bytes;
            const std::vector<unsigned char> all = R_.gatherv(xy, (std::size_t)nloc * sizeof(Particle), 0, &bytes);
            if (R_.rank != 0) break;
            const double *g = reinterpret_cast<const double *>(all.data());
            const long long ntot = (long long)(all.size() / sizeof(Particle));
            if (kind == Kind::Xyz) {
                xyz_header(s.text, ntot, box);
                xyz_rows(s.text, g, ntot);
            } else {
                dump_header(s.text, ntot, box, ts);
                long long first = 0;
                for (int r = 0; r < R_.size; ++r) {  // rank order = id order; the type is the rank of origin
                    const long long n = bytes[(std::size_t)r] / (long long)sizeof(Particle);
                    dump_rows(s.text, g + 2 * first, n, first + 1, r);
                    first += n;
                }
            }
            s.text.flush();
            break;
        }
        case Mode::MPIIO: {
            open_shared(s, with_ext(base_, ext_of(kind)), kind == Kind::Xyz ? "XYZ" : "DUMP");
            const std::vector<long long> counts = R_.allgather(nloc);
            long long ntot = 0, before = 0;
            for (int r = 0; r < R_.size; ++r) {
                ntot += counts[(std::size_t)r];
                if (r < R_.rank) before += counts[(std::size_t)r];
            }
            std::ostringstream head, body;
            if (R_.rank == 0) {
                if (kind == Kind::Xyz) xyz_header(head, ntot, box); else dump_header(head, ntot, box, ts);
            }
            if (kind == Kind::Xyz) {
                xyz_rows(body, xy, nloc);
            } else {
                body << std::setprecision(17);
                dump_rows(body, xy, nloc, before + 1, R_.rank);
            }
            const std::string h = head.str(), b = body.str();
            // byte ranges: [offset, offset + header) by rank 0, then the bodies in rank order
            long long hdr = (long long)h.size();
            R_.bcast(&hdr, sizeof hdr, 0);
            const std::vector<long long> sizes = R_.allgather((long long)b.size());
            long long lower = 0, sum = 0;
            for (int r = 0; r < R_.size; ++r) {
                sum += sizes[(std::size_t)r];
                if (r < R_.rank) lower += sizes[(std::size_t)r];
            }
            if (R_.rank == 0 && hdr > 0) Ranks::write_all(s.fd, h.data(), h.size(), (off_t)s.offset, "trajectory file");
            if (!b.empty()) Ranks::write_all(s.fd, b.data(), b.size(), (off_t)(s.offset + hdr + lower), "trajectory file");
            s.offset += hdr + sum;
            R_.barrier();  // the frame is complete when the call returns
            break;
        }
        }
    }
};

}  // namespace mc2d_slabs

#endif
