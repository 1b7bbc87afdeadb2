/*
 * mpi.h -- minimal single-node MPI stand-in (TEST / BASELINE INFRASTRUCTURE).
 *
 * The image has no MPI, so the reference's attention-mpi.c cannot be built as
 * shipped.  This header + mpi_shim.c provide exactly the subset that file uses
 * (12 functions, 5 types, 8 constants -- SURVEY.md section 8c) so it compiles
 * UNMODIFIED from /root/reference into oracle/_ref/.  Ranks are forked inside
 * MPI_Init (count from env MPI_SHIM_NP, default 1) and talk through one
 * anonymous shared mapping.  Non-blocking collectives complete eagerly at the
 * call (legal: every rank issues collectives in the same order) and MPI_Wait
 * is a no-op.  Nothing here is derived from an MPI implementation's sources.
 */
#ifndef SDPA_ORACLE_MPI_SHIM_H
#define SDPA_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_FLOAT 2
#define MPI_DOUBLE 3
#define MPI_MAX 1
#define MPI_SUM 2
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_SUCCESS 0

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
double MPI_Wtime(void);

int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Ibcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm,
               MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* status);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs,
                 MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt,
               MPI_Op op, int root, MPI_Comm comm);
int MPI_Ireduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt,
                MPI_Op op, int root, MPI_Comm comm, MPI_Request* req);
int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt,
                   MPI_Op op, MPI_Comm comm, MPI_Request* req);

#ifdef __cplusplus
}
#endif
#endif
