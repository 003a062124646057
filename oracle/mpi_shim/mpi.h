/* mpi.h -- minimal MPI shim used ONLY to build the CPU oracle (oracle/_ref).
 *
 * TEST INFRASTRUCTURE, not product code.  The reference (geodynamics/citcomcu)
 * is pure C99 + MPI; this container and the GPU boxes ship no MPI.  This header
 * plus mpi_shim.c implement exactly the 18 MPI entry points the reference uses
 * (list probed with grep over /root/reference/src: Init, Finalize, Comm_rank,
 * Comm_size, Allreduce, Bcast, Isend, Irecv, Send, Recv, Waitall, Barrier,
 * Comm_group, Group_incl, Comm_create, Comm_free, Group_free, Wtime) on top of
 * fork() + one MAP_SHARED arena, so the unmodified reference sources run as N
 * cooperating processes on the host cores (N from env CCU_MPI_NP, default 1).
 */
#ifndef CCU_ORACLE_MPI_SHIM_H
#define CCU_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_INT 1
#define MPI_FLOAT 2
#define MPI_DOUBLE 3
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *g);
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *newg);
int MPI_Comm_create(MPI_Comm comm, MPI_Group g, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Group_free(MPI_Group *g);
double MPI_Wtime(void);

#ifdef __cplusplus
}
#endif
#endif
