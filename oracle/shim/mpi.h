/*
 * TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product.
 *
 * Minimal single-box stand-in for the handful of MPI calls the reference
 * `cuppens` makes (see SURVEY.md section 2.3 for the call-site inventory:
 * /root/reference/src/main.c:27-30,215-235,397-417,504-542 and
 * /root/reference/src/filehandling.c:263,347-348,415-437,547).
 * This image has no mpicc/mpirun, so `oracle/Makefile` compiles the
 * UNMODIFIED reference sources against this header and `mpishim.c`:
 * `MPI_Init` forks `$MPISHIM_NP - 1` children that talk over socketpairs.
 */
#ifndef CUPPEN_ORACLE_MPI_SHIM_H
#define CUPPEN_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INT 4        /* value = element size in bytes */
#define MPI_DOUBLE 8
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_SUCCESS 0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Get_processor_name(char *name, int *resultlen);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);

/* The reference calls these two upper-case names, which no MPI declares
 * (/root/reference/src/main.c:113,696; README.md:40-41 "Bad Termination"). */
#define MPI_ABORT(c, e) MPI_Abort((c), (e))
#define MPI_FINALIZE() MPI_Finalize()

#ifdef __cplusplus
}
#endif
#endif
