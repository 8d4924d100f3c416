/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into / imported by the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker.
 *
 * Plain-C restatement of the reference's algorithm for the hot path (Cuppen's divide and
 * conquer for a symmetric tridiagonal T), following SURVEY.md Appendix A.  Every function
 * cites the reference lines it restates (paths under /root/reference).  Parity status:
 * PINNED -- tests/test_oracle.py checks this restatement against (i) the tinyL goldens and the
 * scheme-2 closed form, (ii) committed outputs of the reference itself (oracle/_ref/cuppens_ref,
 * the unmodified reference sources compiled with the shims in oracle/shim/) in tests/golden/.
 *
 * The one third-party routine on the path, LAPACKE_dsteqr (Intel MKL, unpinned, not vendored;
 * src/main.c:460), is replaced by the implicit QL iteration in tridiag_ql.c.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tridiag_ql.h"

typedef struct { double e; int i; } DiagElem;           /* src/helper.h:100-104 */

/* stable merge sort on (e) -- glibc's qsort is a stable merge sort for these sizes, and the
 * reference comparator returns 0 on ties (src/helper.c:95-103) */
static void sort_diag(DiagElem *a, int n) {
    if (n < 2) return;
    DiagElem *tmp = (DiagElem *)malloc((size_t)n * sizeof(DiagElem));
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (a[j].e < a[i].e) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, (size_t)n * sizeof(DiagElem));
    }
    free(tmp);
}

/* src/eigenvalues.c:8-17 */
static double secular(double lambda, double roh, const double *z, const double *D, int n, const int *G) {
    double sum = 0;
    for (int i = 0; i < n; ++i)
        if (G[i] == -1) sum += z[i] * z[i] / (D[i] - lambda);
    return 1 + roh * sum;
}

/* plain dnrm2 stand-in (src/eigenvalues.c:141,281 call cblas_dnrm2) */
static double nrm2(int n, const double *x) {
    double scale = 0.0, ssq = 1.0;
    for (int i = 0; i < n; ++i) {
        double a = fabs(x[i]);
        if (a != 0.0) {
            if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    }
    return scale * sqrt(ssq);
}

/* getEigenVector, src/eigenvalues.c:291-358 */
static void get_eigenvector(int n, const double *D, const double *z, const double *L, const double *N,
                            const double *C, const double *S, const int *G, const int *P, int numGR,
                            double *ev, int i) {
    int j;
    if (G[i] != -1) {
        for (j = 0; j < n; ++j) ev[j] = (j == i) ? 1 : 0;
    } else {
        for (j = 0; j < n; ++j) {
            if (G[j] < -1) ev[j] = 0;
            else ev[j] = z[j] / ((D[j] - L[i]) * N[i]);
        }
    }
    for (j = numGR - 1; j >= 0; --j) {
        int a = P[j], b = G[a];
        double c = C[j], s = S[j];
        double ti = c * ev[a] + s * ev[b];
        double tj = -s * ev[a] + c * ev[b];
        ev[a] = ti; ev[b] = tj;
    }
}

/*
 * One rank-one update merge: computeEigenvalues (src/eigenvalues.c:19-255) followed by
 * computeNormalizationFactors (src/eigenvalues.c:257-289).
 * In/out: D[m], z[m] (mutated by the Givens sweep exactly like the reference).
 * Out: G[m], L[m], N[m], P[m], C[m], S[m]; returns numGR.
 */
int cuppen_oracle_merge(int n, double *D, double *z, double roh,
                        int *G, double *L, double *N, int *P, double *C, double *S) {
    int i, numGR = 0;
    for (i = 0; i < n; ++i) G[i] = -1;                                  /* :58-60 */
    const double eps = 1e-6;                                            /* :72 */
    for (i = 0; i < n; ++i) if (fabs(z[i]) < eps) G[i] = -2;            /* :75-81 */
    DiagElem *SD = (DiagElem *)malloc((size_t)n * sizeof(DiagElem));
    for (i = 0; i < n; ++i) { SD[i].e = D[i]; SD[i].i = i; }            /* :83-88 */
    sort_diag(SD, n);
    for (i = 0; i < n - 1; ++i) {                                       /* :98-135 */
        if (G[SD[i].i] != -2) {
            int nx = i + 1;
            while (G[SD[nx].i] == -2) { if (++nx == n) break; }
            if (nx >= n) { G[SD[i].i] = -1; continue; }
            if (fabs(SD[nx].e - SD[i].e) < 1e-5) {
                int a = SD[i].i, b = SD[nx].i;
                double r = sqrt(z[a] * z[a] + z[b] * z[b]);
                double c = z[b] / r, s = z[a] / r;
                C[numGR] = c; S[numGR] = s;
                G[a] = b; z[b] = r; z[a] = 0; P[numGR] = a;
                double ti = c * c * SD[i].e + s * s * SD[nx].e;
                double tj = s * s * SD[i].e + c * c * SD[nx].e;
                SD[i].e = ti; SD[nx].e = tj; D[a] = ti; D[b] = tj;
                numGR++;
            }
        }
    }
    double normZ = nrm2(n, z);                                          /* :141 */
    const long maxIter = 10000;                                         /* :146 */
    /* the first live position (for roh<0) is the only sequential dependency of the loop at :161 */
    int firstLive = -1;
    for (i = 0; i < n; ++i) if (G[SD[i].i] == -1) { firstLive = i; break; }
#pragma omp parallel for schedule(dynamic, 16)
    for (i = 0; i < n; ++i) {                                           /* :161-247 */
        double lambda = 0, a, b, fa, fl;
        int ind = SD[i].i;
        double di = SD[i].e;
        if (G[ind] != -1) { L[ind] = di; continue; }
        if (roh < 0) {                                                  /* :174-189 */
            if (i == firstLive) {
                a = di - normZ;
                int j = 0;
                while (secular(a, roh, z, D, n, G) < 0) { a -= normZ; if (++j >= 100) break; }
            } else {
                int p = i - 1;
                while (G[SD[p].i] != -1) p--;
                a = SD[p].e;
            }
            b = di;
        } else {                                                        /* :190-208 */
            a = di;
            int p = i + 1;
            while (p < n && G[SD[p].i] != -1) p++;
            if (p >= n) {
                b = di + normZ;
                int j = 0;
                while (secular(b, roh, z, D, n, G) < 0) { b += normZ; if (++j >= 100) break; }
            } else b = SD[p].e;
        }
        long j = 0;
        while (++j < maxIter) {                                         /* :210-243 */
            lambda = (a + b) / 2;
            fa = secular(a, roh, z, D, n, G);
            fl = secular(lambda, roh, z, D, n, G);
            if (fa == INFINITY || fa == -INFINITY) fa = (roh > 0 ? -INFINITY : INFINITY);
            if (fl == 0 || (b - a) / 2 < 1e-14) break;
            if ((fa >= 0 && fl >= 0) || (fa < 0 && fl < 0)) a = lambda; else b = lambda;
        }
        L[ind] = lambda;
    }
    free(SD);
    /* computeNormalizationFactors, :257-289 (getEigenVector with N==1, then dnrm2) */
    double *ones = (double *)malloc((size_t)n * sizeof(double));
    for (i = 0; i < n; ++i) ones[i] = 1;
#pragma omp parallel
    {
        double *ev = (double *)malloc((size_t)n * sizeof(double));
#pragma omp for schedule(dynamic, 16)
        for (i = 0; i < n; ++i) {
            if (G[i] != -1) N[i] = 1;
            else { get_eigenvector(n, D, z, L, ones, C, S, G, P, numGR, ev, i); N[i] = nrm2(n, ev); }
        }
        free(ev);
    }
    free(ones);
    return numGR;
}

/* ---- tree (src/backtransformation.c:28-114) ------------------------------------------- */
typedef struct Node {
    int o, n, numLeaves;
    int left, right, parent;      /* indices into the per-stage arrays; left==right: pass-through */
    double beta, theta;
    double *L;                    /* eigenvalues, original-index order */
    double *Q;                    /* n x n column-major eigenvector matrix of this node's T */
} Node;

typedef struct { int n; Node *s; } Stage;

/*
 * Full solve with the reference's P-leaf tree.
 *  D[n], E[n-1]      input (not modified)
 *  P                 number of reference ranks / leaves
 *  lambda[n]         ascending eigenvalues (order of the output file, src/filehandling.c:315-321)
 *  resid[n]          ||T x - lambda x||_2 per line (src/filehandling.c:511-531), may be NULL
 *  V                 n x n column-major, column k = eigenvector of lambda[k], may be NULL
 *  stats             per merge, leaf level first, left to right: {offset, m, zdefl, givens} (may be NULL)
 *  rhos              per merge beta*theta (may be NULL);  *nmerges receives the count
 * returns 0, or 4 when n < P (src/main.c:324-327).
 */
int cuppen_oracle_solve(int n, const double *Din, const double *Ein, int P,
                        double *lambda, double *resid, double *V,
                        int *stats, double *rhos, int *nmerges) {
    if (n / P == 0) return 4;
    int depth = 1, maxMod = 1;
    while (maxMod < P) { maxMod *= 2; depth++; }                        /* src/main.c:274-281 */
    Stage *st = (Stage *)calloc((size_t)depth, sizeof(Stage));
    int s, j;
    for (s = 0; s < depth; ++s) {                                       /* backtransformation.c:36-79 */
        int h = 1 << (depth - 1 - s);
        int cnt = (P - 1) / h + 1;
        st[s].n = cnt;
        st[s].s = (Node *)calloc((size_t)cnt, sizeof(Node));
        for (j = 0; j < cnt; ++j) { st[s].s[j].left = st[s].s[j].right = st[s].s[j].parent = -1; }
        if (s > 0)
            for (j = 0; j < cnt; ++j) {
                Node *par = &st[s - 1].s[j / 2];
                st[s].s[j].parent = j / 2;
                if (j % 2 == 0) { par->left = j; if (j == cnt - 1) par->right = j; }
                else par->right = j;
            }
    }
    int leafSize = n / P, rem = n % P, off = 0;                         /* backtransformation.c:85-96 */
    for (j = 0; j < P; ++j) {
        Node *nd = &st[depth - 1].s[j];
        nd->n = leafSize + (j < rem ? 1 : 0); nd->o = off; off += nd->n; nd->numLeaves = 1;
    }
    for (s = depth - 2; s >= 0; --s) {                                  /* :97-110 */
        off = 0;
        for (j = 0; j < st[s].n; ++j) {
            Node *nd = &st[s].s[j];
            Node *l = &st[s + 1].s[nd->left], *r = &st[s + 1].s[nd->right];
            if (nd->left == nd->right) { nd->n = l->n; nd->numLeaves = l->numLeaves; }
            else { nd->n = l->n + r->n; nd->numLeaves = l->numLeaves + r->numLeaves; }
            nd->o = off; off += nd->n;
        }
    }
    double *D = (double *)malloc((size_t)n * sizeof(double));
    memcpy(D, Din, (size_t)n * sizeof(double));
    /* divide, top-down on the already modified D (src/main.c:339-421) */
    for (s = 0; s < depth - 1; ++s)
        for (j = 0; j < st[s].n; ++j) {
            Node *nd = &st[s].s[j];
            if (nd->left == nd->right) continue;
            int n1 = st[s + 1].s[nd->left].n;
            int g = nd->o + n1;                         /* global index of the first row of T2 */
            nd->beta = Ein[g - 1];
            double dl = D[g - 1], df = D[g];
            if ((dl > 0 && df > 0) || (dl < 0 && df < 0)) {             /* :370-375 */
                nd->theta = ((dl * (-nd->beta)) < 0) ? -1 : 1;
            } else {                                                    /* :376-389 */
                if (fabs(nd->beta) < fabs(df)) nd->theta = 1000 * nd->beta;
                else nd->theta = nd->beta / 1000;
            }
            D[g - 1] -= nd->theta * nd->beta;                           /* :392-394 */
            D[g] -= 1.0 / nd->theta * nd->beta;
        }
    /* leaves (src/main.c:460-474) */
    int rc = 0;
    for (j = 0; j < P; ++j) {
        Node *nd = &st[depth - 1].s[j];
        nd->L = (double *)malloc((size_t)nd->n * sizeof(double));
        nd->Q = (double *)malloc((size_t)nd->n * nd->n * sizeof(double));
        memcpy(nd->L, D + nd->o, (size_t)nd->n * sizeof(double));
        if (cuppen_oracle_tql2(nd->n, nd->L, Ein + nd->o, nd->Q, nd->n, 0) != 0) rc = -1;
    }
    /* conquer, bottom-up (src/main.c:495-664), with Q materialised: Q_parent = diag(Q1,Q2) U,
     * which is what writeResults evaluates lazily row by row (src/filehandling.c:376-507) */
    int nm = 0;
    for (s = depth - 2; s >= 0; --s)
        for (j = 0; j < st[s].n; ++j) {
            Node *nd = &st[s].s[j];
            Node *l = &st[s + 1].s[nd->left], *r = &st[s + 1].s[nd->right];
            if (nd->left == nd->right) { nd->L = l->L; nd->Q = l->Q; l->L = NULL; l->Q = NULL; continue; }
            int n1 = l->n, n2 = r->n, m = n1 + n2, i, k;
            double *Dm = (double *)malloc((size_t)m * sizeof(double));
            double *z = (double *)malloc((size_t)m * sizeof(double));
            memcpy(Dm, l->L, (size_t)n1 * sizeof(double));              /* :534-538 */
            memcpy(Dm + n1, r->L, (size_t)n2 * sizeof(double));
            for (i = 0; i < n1; ++i) z[i] = l->Q[(size_t)(n1 - 1) + (size_t)i * n1];     /* helper.c:36-50 */
            for (i = 0; i < n2; ++i) z[n1 + i] = r->Q[(size_t)0 + (size_t)i * n2] / nd->theta;
            int *G = (int *)malloc((size_t)m * sizeof(int)), *Pp = (int *)malloc((size_t)m * sizeof(int));
            double *L = (double *)malloc((size_t)m * sizeof(double)), *N = (double *)malloc((size_t)m * sizeof(double));
            double *C = (double *)malloc((size_t)m * sizeof(double)), *S = (double *)malloc((size_t)m * sizeof(double));
            double roh = nd->beta * nd->theta;                          /* eigenvalues.c:54 */
            int numGR = cuppen_oracle_merge(m, Dm, z, roh, G, L, N, Pp, C, S);
            if (stats) {
                int zd = 0; for (i = 0; i < m; ++i) if (G[i] == -2) zd++;
                stats[4 * nm] = nd->o; stats[4 * nm + 1] = m; stats[4 * nm + 2] = zd; stats[4 * nm + 3] = numGR;
            }
            if (rhos) rhos[nm] = roh;
            nm++;
            double *Q = (double *)calloc((size_t)m * m, sizeof(double));
#pragma omp parallel
            {
                double *ev = (double *)malloc((size_t)m * sizeof(double));
#pragma omp for schedule(dynamic, 8)
                for (i = 0; i < m; ++i) {
                    get_eigenvector(m, Dm, z, L, N, C, S, G, Pp, numGR, ev, i);
                    double *col = Q + (size_t)i * m;
                    for (k = 0; k < n1; ++k) {
                        double u = ev[k]; if (u == 0) continue;
                        const double *qc = l->Q + (size_t)k * n1;
                        for (int rr = 0; rr < n1; ++rr) col[rr] += qc[rr] * u;
                    }
                    for (k = 0; k < n2; ++k) {
                        double u = ev[n1 + k]; if (u == 0) continue;
                        const double *qc = r->Q + (size_t)k * n2;
                        for (int rr = 0; rr < n2; ++rr) col[n1 + rr] += qc[rr] * u;
                    }
                }
                free(ev);
            }
            free(Dm); free(z); free(G); free(Pp); free(N); free(C); free(S);
            free(l->L); free(l->Q); free(r->L); free(r->Q); l->L = l->Q = r->L = r->Q = NULL;
            nd->L = L; nd->Q = Q;
        }
    if (nmerges) *nmerges = nm;
    /* output ordering + residual (src/filehandling.c:315-321, 511-531) */
    Node *root = &st[0].s[0];
    DiagElem *SL = (DiagElem *)malloc((size_t)n * sizeof(DiagElem));
    for (j = 0; j < n; ++j) { SL[j].e = root->L[j]; SL[j].i = j; }
    sort_diag(SL, n);
#pragma omp parallel for schedule(static)
    for (j = 0; j < n; ++j) {
        const double *xi = root->Q + (size_t)SL[j].i * n;
        double lam = SL[j].e;
        lambda[j] = lam;
        if (V) memcpy(V + (size_t)j * n, xi, (size_t)n * sizeof(double));
        if (resid) {
            double norm = 0;
            for (int k = 0; k < n; ++k) {
                double x;
                if (n == 1) x = Din[0] * xi[0];
                else if (k == 0) x = Din[0] * xi[0] + Ein[0] * xi[1];
                else if (k == n - 1) x = Ein[n - 2] * xi[n - 2] + Din[n - 1] * xi[n - 1];
                else x = Ein[k - 1] * xi[k - 1] + Din[k] * xi[k] + Ein[k] * xi[k + 1];
                x -= lam * xi[k];
                norm += x * x;
            }
            resid[j] = sqrt(norm);
        }
    }
    free(SL); free(root->L); free(root->Q); free(D);
    for (s = 0; s < depth; ++s) free(st[s].s);
    free(st);
    return rc;
}

/* src/helper.c:7-33 */
void cuppen_oracle_scheme(int scheme, int n, double *D, double *E) {
    if (scheme == 1) {
        double sp = (100.0 - 1.0) / (n - 1);
        for (int i = 0; i < n - 1; ++i) { E[i] = -1; D[i] = 1.0 + i * sp; }
        D[n - 1] = 1.0 + (n - 1) * sp;
    } else {
        for (int i = 0; i < n - 1; ++i) { E[i] = -1; D[i] = 2; }
        D[n - 1] = 2.0;
    }
}
