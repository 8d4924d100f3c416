/*
 * TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product.
 * The reference includes "mkl.h" for exactly two routines
 * (/root/reference/src/main.c:460 LAPACKE_dsteqr, /root/reference/src/eigenvalues.c:141,281
 * cblas_dnrm2).  Intel MKL is not in this image; they are mapped either to the
 * OpenBLAS that scipy bundles (symbols carry a scipy_ prefix, LP64) or, when
 * that library is absent, to the plain-C stand-ins in shim/lapack_standin.c.
 */
#ifndef CUPPEN_ORACLE_MKL_SHIM_H
#define CUPPEN_ORACLE_MKL_SHIM_H
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#ifdef CUPPEN_SHIM_USE_SCIPY_OPENBLAS
int scipy_LAPACKE_dsteqr(int layout, char compz, int n, double *d, double *e, double *z, int ldz);
double scipy_cblas_dnrm2(int n, const double *x, int incx);
#define LAPACKE_dsteqr scipy_LAPACKE_dsteqr
#define cblas_dnrm2 scipy_cblas_dnrm2
#else
int LAPACKE_dsteqr(int layout, char compz, int n, double *d, double *e, double *z, int ldz);
double cblas_dnrm2(int n, const double *x, int incx);
#endif
#endif
