/*
 * mpi.h — TEST INFRASTRUCTURE.  A threads-as-ranks stand-in for the ~25 MPI calls the reference's
 * mpi_src/ uses (this image has no MPI: no mpicxx, mpirun or mpi.h).  It exists for ONE purpose:
 * compiling /root/reference/mpi_src/*.cpp UNMODIFIED (oracle/Makefile, target ref_mpi) so that
 * bench.py can time the reference's MPI code path on the host cores beside the GPU numbers
 * ("cpu_baseline_mpi": reference MPI code on shim, N threads).  Every "rank" is a thread of one process;
 * point-to-point messages are eager copies through per-rank mailboxes with MPI's non-overtaking order
 * per (source, tag); collectives meet at a spinning barrier and copy straight out of each other's
 * buffers.  It is not a general MPI: one communicator, blocking collectives, the datatypes and
 * reduction the reference uses.  Never on the product path; never a source of ground truth
 * (SURVEY.md §7 defect 1: the reference's MPI energies are wrong for Py = 1).
 */
#ifndef MC2D_MPI_SHIM_H
#define MPI_H 1
#define MC2D_MPI_SHIM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef ptrdiff_t MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef struct mc2d_mpi_request *MPI_Request;
typedef struct mc2d_mpi_file *MPI_File;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 1
#define MPI_INFO_NULL 0
#define MPI_FILE_NULL ((MPI_File)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_SUM 1
#define MPI_TYPECLASS_INTEGER 1
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 4

/* predefined datatypes (indices into the shim's size table) */
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_DOUBLE 3
#define MPI_LONG_LONG_INT 4
#define MPI_UINT64_T 5
#define MPI_INT64_T 6
#define MPI_UNSIGNED 7

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *statuses);

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Exscan(const void *send, void *recv, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void *send, int scount, MPI_Datatype sdt, void *recv, int rcount, MPI_Datatype rdt, int root,
               MPI_Comm comm);
int MPI_Gatherv(const void *send, int scount, MPI_Datatype sdt, void *recv, const int *rcounts, const int *displs,
                MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Alltoall(const void *send, int scount, MPI_Datatype sdt, void *recv, int rcount, MPI_Datatype rdt,
                 MPI_Comm comm);
int MPI_Alltoallv(const void *send, const int *scounts, const int *sdispls, MPI_Datatype sdt, void *recv,
                  const int *rcounts, const int *rdispls, MPI_Datatype rdt, MPI_Comm comm);

int MPI_Type_match_size(int typeclass, int size, MPI_Datatype *dt);
int MPI_Get_address(const void *location, MPI_Aint *address);
int MPI_Type_create_struct(int count, const int *blocklens, const MPI_Aint *displs, const MPI_Datatype *types,
                           MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *dt);

int MPI_File_open(MPI_Comm comm, const char *name, int amode, MPI_Info info, MPI_File *fh);
int MPI_File_close(MPI_File *fh);
int MPI_File_get_size(MPI_File fh, MPI_Offset *size);
int MPI_File_write_at_all(MPI_File fh, MPI_Offset offset, const void *buf, int count, MPI_Datatype dt,
                          MPI_Status *status);

#ifdef __cplusplus
}
#endif
#endif
