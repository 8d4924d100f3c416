/*
 * TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product.
 *
 * Observation hooks around two reference functions, added WITHOUT touching the
 * reference sources: oracle/Makefile compiles /root/reference/src/eigenvalues.c
 * with -DcomputeEigenvalues=cuppen_ref_computeEigenvalues and
 * -DcomputeNormalizationFactors=cuppen_ref_computeNormalizationFactors, so the
 * calls in main.c (/root/reference/src/main.c:560,575) land here first.
 *
 *   CUPPEN_ORACLE_STATS=<file>  append one line per merge:
 *        "<offset> <m> <zdefl> <givens> <rho %.17g>"   
 *   CUPPEN_ORACLE_DUMP=<dir>    write <dir>/merge_<offset>_<m>.bin holding the
 *        merge inputs (D, z before deflation) and outputs (D, z after, G, L, N, P, C, S).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "backtransformation.h"
#include "eigenvalues.h"

void cuppen_ref_computeEigenvalues(EVRepNode *node, MPIHandle mpiHandle);
void cuppen_ref_computeNormalizationFactors(EVRepNode *node);

static double *g_preD = NULL, *g_preZ = NULL;

void computeEigenvalues(EVRepNode *node, MPIHandle mpiHandle) {
    if (mpiHandle.taskid == node->taskid && getenv("CUPPEN_ORACLE_DUMP")) {
        int n = node->n;
        g_preD = malloc(n * sizeof(double)); g_preZ = malloc(n * sizeof(double));
        memcpy(g_preD, node->D, n * sizeof(double));
        memcpy(g_preZ, node->z, n * sizeof(double));
    }
    cuppen_ref_computeEigenvalues(node, mpiHandle);
}

void computeNormalizationFactors(EVRepNode *node) {
    int n = node->n, zdefl = 0, i;
    for (i = 0; i < n; ++i) if (node->G[i] == -2) zdefl++;
    const char *stats = getenv("CUPPEN_ORACLE_STATS");
    if (stats) {
        FILE *f = fopen(stats, "a");
        if (f) {
            fprintf(f, "%d %d %d %d %.17g\n", node->o, n, zdefl, node->numGR, node->beta * node->theta);
            fclose(f);
        }
    }
    cuppen_ref_computeNormalizationFactors(node);
    const char *dump = getenv("CUPPEN_ORACLE_DUMP");
    if (dump && g_preD) {
        char path[4096];
        snprintf(path, sizeof path, "%s/merge_%d_%d.bin", dump, node->o, n);
        FILE *f = fopen(path, "wb");
        if (f) {
            double rho = node->beta * node->theta;
            int hdr[4] = { node->o, n, zdefl, node->numGR };
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(&rho, sizeof(double), 1, f);
            fwrite(g_preD, sizeof(double), n, f);
            fwrite(g_preZ, sizeof(double), n, f);
            fwrite(node->D, sizeof(double), n, f);
            fwrite(node->z, sizeof(double), n, f);
            fwrite(node->G, sizeof(int), n, f);
            fwrite(node->L, sizeof(double), n, f);
            fwrite(node->N, sizeof(double), n, f);
            fwrite(node->P, sizeof(int), n, f);
            fwrite(node->C, sizeof(double), n, f);
            fwrite(node->S, sizeof(double), n, f);
            fclose(f);
        }
        free(g_preD); free(g_preZ); g_preD = g_preZ = NULL;
    }
}
