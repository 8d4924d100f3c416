/*
 * TEST INFRASTRUCTURE ONLY -- oracle leaf solver, never linked into the product.
 *
 * Stand-in for the third-party routine the reference calls on every leaf:
 *   LAPACKE_dsteqr(LAPACK_ROW_MAJOR,'I',nl,D,E,Q,nl)   /root/reference/src/main.c:460
 * (Intel MKL, version unpinned, not vendored under /root/reference).  This is a
 * restatement of the published implicit-shift QL iteration (EISPACK TQL2 /
 * Bowdler-Martin-Reinsch-Wilkinson 1968) with the eigenpairs sorted ascending
 * as dsteqr returns them.  Mathematically the leaf decomposition is unique up
 * to eigenvector sign, which the merge does not see (only z_i^2 and products
 * of paired rows matter), so any accurate solver is an admissible stand-in.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "tridiag_ql.h"

/* d[n], e[n-1] in; d ascending eigenvalues out; z: n x n with ld `ldz`,
 * eigenvector j in column j (row-major: z[r*ldz+j]) or, if !row_major,
 * column-major (z[r + j*ldz]).  Returns 0 or the index of a non-converged value. */
int cuppen_oracle_tql2(int n, double *d, const double *e_in, double *z, int ldz, int row_major)
{
    if (n <= 0) return 0;
    double *e = (double *)calloc((size_t)n, sizeof(double));
    /* work on a dense row-major copy, transpose at the end if required */
    double *q = (double *)calloc((size_t)n * n, sizeof(double));
    int i, k, l, m, iter;
    for (i = 0; i < n - 1; ++i) e[i] = e_in[i];
    e[n - 1] = 0.0;
    for (i = 0; i < n; ++i) q[(size_t)i * n + i] = 1.0;
    const double eps = 2.220446049250313e-16;

    for (l = 0; l < n; ++l) {
        iter = 0;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 90) { free(e); free(q); return l + 1; }
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    double b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (k = 0; k < n; ++k) {
                        double *row = q + (size_t)k * n;
                        f = row[i + 1];
                        row[i + 1] = s * row[i] + c * f;
                        row[i] = c * row[i] - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
            }
        } while (m != l);
    }
    /* selection sort ascending, swapping columns */
    for (i = 0; i < n - 1; ++i) {
        int kmin = i; double p = d[i];
        for (k = i + 1; k < n; ++k) if (d[k] < p) { kmin = k; p = d[k]; }
        if (kmin != i) {
            d[kmin] = d[i]; d[i] = p;
            for (k = 0; k < n; ++k) {
                double t = q[(size_t)k * n + i];
                q[(size_t)k * n + i] = q[(size_t)k * n + kmin];
                q[(size_t)k * n + kmin] = t;
            }
        }
    }
    for (i = 0; i < n; ++i)
        for (k = 0; k < n; ++k) {
            if (row_major) z[(size_t)i * ldz + k] = q[(size_t)i * n + k];
            else z[(size_t)i + (size_t)k * ldz] = q[(size_t)i * n + k];
        }
    free(e); free(q);
    return 0;
}
