/*
 * TEST INFRASTRUCTURE ONLY (oracle build) -- used when scipy's OpenBLAS is absent.
 * LAPACKE_dsteqr / cblas_dnrm2 call sites: /root/reference/src/main.c:460,
 * /root/reference/src/eigenvalues.c:141,281.
 */
#include <math.h>
#include "../tridiag_ql.h"
#include "mkl.h"
#ifndef CUPPEN_SHIM_USE_SCIPY_OPENBLAS
int LAPACKE_dsteqr(int layout, char compz, int n, double *d, double *e, double *z, int ldz) {
    (void)compz;
    return cuppen_oracle_tql2(n, d, e, z, ldz, layout == LAPACK_ROW_MAJOR);
}
double cblas_dnrm2(int n, const double *x, int incx) {
    double scale = 0.0, ssq = 1.0;
    for (int i = 0; i < n; ++i) {
        double a = fabs(x[(long)i * incx]);
        if (a != 0.0) {
            if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    }
    return scale * sqrt(ssq);
}
#endif
