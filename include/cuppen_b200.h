/*
 * cuppen_b200.h -- C ABI of libcuppen_b200.so: Cuppen's divide-and-conquer eigensolver for
 * symmetric tridiagonal matrices on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of chrhenning/symmetric_eigenvalue (`cuppens`).
 * Plain pointers and sizes only; every function returns 0 on success and a negative code on
 * error (cuppen_last_error() holds the text); the library never aborts the process and has no
 * CPU fallback: without a CUDA device cuppen_create() fails with CUPPEN_ERR_CUDA.
 *
 * Reference interfaces replaced (paths under /root/reference):
 *   cuppen_create / cuppen_set_tridiagonal   initEVRepTree (src/backtransformation.h:117,
 *                                            src/backtransformation.c:28-114) + the divide phase
 *                                            of main (src/main.c:274-421)
 *   cuppen_solve                             leaf solve LAPACKE_dsteqr (src/main.c:460), conquer loop
 *                                            (src/main.c:495-664) = computeZ (src/helper.h:57),
 *                                            computeEigenvalues / computeNormalizationFactors /
 *                                            getEigenVector (src/eigenvalues.h:21-36), and the
 *                                            back-transformation inside writeResults
 *                                            (src/filehandling.h:81, src/filehandling.c:332-508)
 *   cuppen_get_eigenvalues                   sorted L of the root node (src/filehandling.c:315-321)
 *   cuppen_get_residuals                     ||T x - lambda x||_2 column of the output file
 *                                            (src/filehandling.c:511-538)
 *   cuppen_get_merge_stats                   the deflation bookkeeping G / numGR of EVRepNode
 *                                            (src/backtransformation.h:16-88), one record per merge
 *   cuppen_get_timers                        the five stdout timer lines (src/main.c:672-679,
 *                                            src/filehandling.c:564-570)
 *   cuppen_scheme                            createMatrixScheme1/2 (src/helper.h:28,39)
 *   cuppen_read_mtx                          readSymmTriadiagonalMatrixFromSparseMTX
 *                                            (src/filehandling.h:56, src/filehandling.c:76-153)
 *   cuppen_read_ev_file                      determineEigenvectorsToCompute (src/filehandling.h:69)
 *   cuppen_write_results                     the fprintf loop of writeResults (src/filehandling.c:537,544)
 *
 * One process drives one GPU.  Multi-GPU runs are SPMD (one process per GPU, like the
 * reference's one-rank-per-leaf MPI layout): every rank calls the same functions with its rank
 * and the world size, plus either an NCCL unique id (cuppen_nccl_unique_id on rank 0, shipped
 * to the others by the launcher) or a table of communication callbacks.
 */
#ifndef CUPPEN_B200_H
#define CUPPEN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cuppen_handle_s* cuppen_handle;

enum {
    CUPPEN_OK = 0,
    CUPPEN_ERR_ARG = -1,        /* bad argument */
    CUPPEN_ERR_IO = -2,         /* unreadable / invalid input file (reference exit code 2 or 3) */
    CUPPEN_ERR_ZERO = -3,       /* zero diagonal / off-diagonal entry (the reference asserts, src/main.c:196-200) */
    CUPPEN_ERR_LEAF = -4,       /* n < reference leaves "Leaf Size is too small" (reference exit code 4) */
    CUPPEN_ERR_STATE = -5,      /* call order */
    CUPPEN_ERR_CUDA = -10,      /* CUDA failure or no device */
    CUPPEN_ERR_NOMEM = -11,
    CUPPEN_ERR_COMM = -12,      /* NCCL / communicator failure */
    CUPPEN_ERR_CONVERGENCE = -13
};

enum {
    CUPPEN_FLAG_VECTORS = 1,    /* materialise eigenvectors (the reference's -e) */
    CUPPEN_FLAG_NO_RESIDUALS = 2,
    CUPPEN_FLAG_SELECT = 4      /* eigenvectors of selected eigenvalues only (the reference's -eFILE): no n x n matrix
                                   is formed, the selected columns are pushed through the implicit U factors of
                                   the tree; see cuppen_select_eigenvectors.  Several ranks: the (cheap) eigenvalue-only
                                   decomposition is replicated, the selected vectors are dealt to the ranks and
                                   gathered, so every rank returns all of them */
};

/* One record per merge of the tree. */
typedef struct {
    int offset;        /* first global index of the node */
    int m;             /* size of the merged problem */
    int n1;            /* size of the left child */
    int mode;          /* 1: merge exists in the reference's P-leaf tree (reference thresholds), 0: below a reference leaf */
    int zdefl;         /* entries with G == -2 (src/eigenvalues.c:75-81) */
    int givens;        /* numGR (src/eigenvalues.c:98-135) */
    int k;             /* secular roots solved */
    int height;        /* tree level, leaves = 0 */
    double rho;        /* beta*theta (mode 1) */
} cuppen_merge_stat;

typedef struct {
    double total_s;          /* "Required time to compute all eigenvalues" */
    double root_finding_s;   /* "Required time for root finding" */
    double ev_extract_s;     /* "eigenvector extraction from U_i's" (normalisation + Loewner + U generation) */
    double backtransform_s;  /* "Required time for backtransformation" (pack + GEMM + residual) */
    double backtransform_ev_s; /* U generation inside the back-transformation */
    double gemm_s;           /* device time of the DMMA GEMM launches */
    double gemm_flop;        /* executed flop: sum 2*M*N*K over GEMM problems */
    double leaf_s, deflation_s, pack_s, residual_s;
    double device_s;         /* CUDA-event time from the first to the last launch of cuppen_solve */
    double pack_bytes;       /* algorithmic bytes of pack_kernel (in place): per merge 8*rows*(zdefl/2 + live + 1.5*rotated) */
    double ugen_bytes;       /* algorithmic bytes of ugen_kernel: 8*K*N written */
    double secular_root_iters; /* reserved */
    long   kernel_launches;
    double apply_s;          /* CUPPEN_FLAG_SELECT: device time of the back-application of the selected columns */
    double comm_s;           /* several GPUs, peer-memory back end: device time of the row redistribution and the barriers */
    long   comm_mode;        /* 0 one GPU, 1 NCCL collectives / callbacks between the kernels, 2 peer memory (NVLink loads/stores in our kernels) */
} cuppen_timers;

/* Communication callbacks for world > 1 when NCCL is not used (tests drive these with gloo).
 * Buffers are device buffers of the library (host memory in the test-only host build). */
typedef struct {
    void* user;
    /* broadcast `bytes` from rank `root` to the contiguous rank group [lo, lo+cnt) */
    int (*group_bcast)(void* user, void* buf, size_t bytes, int root, int lo, int cnt);
    int (*allreduce_sum_f64)(void* user, double* buf, size_t count);
    int (*allgather)(void* user, const void* send, void* recv, size_t bytes_per_rank);
    int (*allreduce_sum_i32)(void* user, int* buf, size_t count);
    /* personalised exchange: send[r] (sbytes[r] bytes) goes to rank r, recv[r] (rbytes[r]) comes from rank r;
     * entries for the calling rank itself are ignored */
    int (*alltoallv)(void* user, const void* const* send, const size_t* sbytes, void* const* recv, const size_t* rbytes);
} cuppen_comm_callbacks;

#define CUPPEN_NCCL_ID_BYTES 128

/* ---- life cycle ---------------------------------------------------------------------------- */
/* n: matrix size; ref_leaves: P of `mpirun -n P` whose tree and thresholds are reproduced
 * (1 = no reference merge levels, LAPACK-grade tolerances throughout); device: CUDA ordinal. */
int cuppen_create(cuppen_handle* h, int n, int ref_leaves, int flags, int device);
/* SPMD variants: rank/world + NCCL unique id (128 bytes) or callbacks. */
int cuppen_nccl_unique_id(unsigned char id[CUPPEN_NCCL_ID_BYTES]);
int cuppen_create_nccl(cuppen_handle* h, int n, int ref_leaves, int flags, int device,
                       int rank, int world, const unsigned char id[CUPPEN_NCCL_ID_BYTES]);
int cuppen_create_callbacks(cuppen_handle* h, int n, int ref_leaves, int flags, int device,
                            int rank, int world, const cuppen_comm_callbacks* cb);
int cuppen_destroy(cuppen_handle h);

/* ---- the path ------------------------------------------------------------------------------ */
/* D[n], E[n-1]: host pointers, copied.  Non-finite entries are rejected (CUPPEN_ERR_ARG). */
int cuppen_set_tridiagonal(cuppen_handle h, const double* D, const double* E);
/* May be called repeatedly: the matrix stays resident on the device, a second cuppen_solve re-runs the whole
 * decomposition without any host<->device traffic of the inputs (one-GPU handles replay a CUDA graph). */
int cuppen_solve(cuppen_handle h);

/* ---- results (host pointers) ---------------------------------------------------------------- */
int cuppen_get_eigenvalues(cuppen_handle h, double* lambda_ascending /* n */);
/* residual of the eigenvector of the idx[i]-th smallest eigenvalue (0-based); idx==NULL: all n. */
int cuppen_get_residuals(cuppen_handle h, const int* idx, int cnt, double* out);
int cuppen_get_merge_stats(cuppen_handle h, cuppen_merge_stat* out, int capacity, int* count);
int cuppen_get_timers(cuppen_handle h, cuppen_timers* out);
/* rows [*row0, *row0+*rows) of V held by this rank, columns in ascending-lambda order,
 * column-major with leading dimension ld (>= *rows).  V may be NULL to query the row range. */
int cuppen_local_rows(cuppen_handle h, int* row0, int* rows);
/* global row index of every local row (length *rows of cuppen_local_rows): with several GPUs a rank
 * holds one slice of every subtree of the divide tree's top levels, not one contiguous range */
int cuppen_local_row_map(cuppen_handle h, int* global_rows);
int cuppen_copy_eigenvectors(cuppen_handle h, double* V, long ld);
/* the listed columns only (idx[i] = 0-based rank in ascending-lambda order): local rows x cnt, column-major, ld >= rows.
 * What the reference's -eFILE hands to its residual loop one vector at a time (src/filehandling.c:339-345,508). */
int cuppen_copy_eigenvector_columns(cuppen_handle h, const int* idx, int cnt, double* V, long ld);
/* Selected-eigenvector mode (handle created with CUPPEN_FLAG_SELECT): replaces determineEigenvectorsToCompute's
 * EVToCompute list (src/filehandling.h:10-24,69) and the per-index loop of writeResults (src/filehandling.c:339-345).
 * idx[i] = 0-based rank in ascending-lambda order; takes effect at the next cuppen_solve.  Afterwards
 * cuppen_get_residuals answers for the selected ranks (idx == NULL: length-n array, NaN where not selected) and
 * cuppen_copy_selected_eigenvectors returns the n x cnt matrix (column t = eigenvector of rank idx[t], column-major, ld >= n). */
int cuppen_select_eigenvectors(cuppen_handle h, const int* idx, int cnt);
int cuppen_copy_selected_eigenvectors(cuppen_handle h, double* V, long ld);
/* max_ij |(V^T V - I)_ij| of the computed eigenvector matrix, evaluated on the GPU by a DMMA Gram kernel that never
 * writes the product (BASELINE's orthogonality criterion; no counterpart in the reference, which cannot emit V).
 * CUPPEN_FLAG_VECTORS.  seconds (may be NULL): device time of the check.  Several ranks: every rank calls; the row slices
 * are all-gathered and every rank evaluates its share of the tiles of the Gram triangle (seconds then includes the gather). */
int cuppen_orthogonality(cuppen_handle h, double* max_abs_dev, double* seconds);
/* Eigenvector file (binary): "CUPPENV1" | int64 n | int64 ncols | int64 rank[ncols] | double lambda[ncols] |
 * double V[ncols][n].  All n vectors (CUPPEN_FLAG_VECTORS, ascending lambda) or the selected ones (CUPPEN_FLAG_SELECT).
 * Several ranks: every rank calls (the row slices are gathered panel by panel), rank 0 writes; filename may be NULL elsewhere. */
int cuppen_write_eigenvectors(cuppen_handle h, const char* filename);
const char* cuppen_last_error(void);

/* ---- dense front end (no counterpart in the reference, whose input is tridiagonal) ----------
 * Full eigendecomposition A = Z diag(W) Z^T of a dense symmetric matrix on one GPU: blocked Householder
 * tridiagonalisation (dsytrd / dlatrd structure, symv-bound), the tridiagonal path above under the accurate rule, and the
 * back-transformation Z = Q V through compact-WY block reflectors (dormtr / dlarft structure, DMMA GEMMs).
 * A: host, column-major, symmetric, ld = lda (the lower triangle is read); W[n] ascending; Z: n x n column-major, ld = ldz,
 * or NULL for eigenvalues only; tm may be NULL. */
typedef struct {
    double tridiagonalise_s;      /* device time of the Householder reduction */
    double tridiagonal_solve_s;   /* wall time of the tridiagonal eigenproblem (handle creation + solve) */
    double tridiagonal_device_s;  /* device time of the tridiagonal solve alone */
    double backtransform_s;       /* device time of the block-reflector back-transformation */
} cuppen_dense_timers;
int cuppen_dense_eigh(int n, const double* A, long lda, double* W, double* Z, long ldz, int device, cuppen_dense_timers* tm);

/* ---- host-side helpers of the CLI (no GPU involved) ------------------------------------------- */
int cuppen_scheme(int scheme, int n, double* D, double* E);
int cuppen_read_mtx(const char* filename, double** D, double** E, int* n);   /* callee allocates (malloc) */
int cuppen_read_ev_file(const char* filename, int n, int** indices, int* count);
int cuppen_write_results(const char* filename, int n, const double* lambda, const double* resid,
                         int all_vectors, const int* indices, int count);

#ifdef __cplusplus
}
#endif
#endif
