/*
 * Host-side (no GPU) pieces of the `cuppens` boundary, plain C:
 *   cuppen_scheme          createMatrixScheme1/2          /root/reference/src/helper.c:7-33
 *   cuppen_read_mtx        readSymmTriadiagonalMatrixFromSparseMTX  /root/reference/src/filehandling.c:76-153
 *                          (banner / size parsing of the vendored NIST mmio it calls:
 *                           /root/reference/lib/mmio.c:95-216; the unused rest of mmio is out of scope)
 *   cuppen_read_ev_file    determineEigenvectorsToCompute /root/reference/src/filehandling.c:165-239
 *   cuppen_write_results   output loop of writeResults    /root/reference/src/filehandling.c:332-345,537,544
 * Same accept/reject behaviour and the same diagnostics as the reference.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/cuppen_b200.h"

int cuppen_scheme(int scheme, int n, double* D, double* E) {
    int i;
    if (n < 1 || !D || (n > 1 && !E) || scheme < 1 || scheme > 2) return CUPPEN_ERR_ARG;
    if (scheme == 1) {
        double diagSpacing = (100.0 - 1.0) / (n - 1);
        for (i = 0; i < n - 1; ++i) { E[i] = -1; D[i] = 1.0 + i * diagSpacing; }
        D[n - 1] = 1.0 + (n - 1) * diagSpacing;
    } else {
        for (i = 0; i < n - 1; ++i) { E[i] = -1; D[i] = 2; }
        D[n - 1] = 2.0;
    }
    return CUPPEN_OK;
}

static void lower(char* s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); }

/* Matrix Market banner: "%%MatrixMarket matrix <coordinate|array> <field> <symmetry>" */
static int parse_banner(FILE* f, char type[4][64]) {
    char line[1025], banner[64];
    if (!fgets(line, sizeof line, f)) return -1;
    if (sscanf(line, "%63s %63s %63s %63s %63s", banner, type[0], type[1], type[2], type[3]) != 5) return -1;
    lower(type[0]); lower(type[1]); lower(type[2]); lower(type[3]);
    if (strncmp(banner, "%%MatrixMarket", 14) != 0) return -1;
    if (strcmp(type[0], "matrix") != 0) return -1;
    if (strcmp(type[1], "coordinate") != 0 && strcmp(type[1], "array") != 0) return -1;
    if (strcmp(type[2], "real") != 0 && strcmp(type[2], "complex") != 0 && strcmp(type[2], "pattern") != 0 &&
        strcmp(type[2], "integer") != 0) return -1;
    if (strcmp(type[3], "general") != 0 && strcmp(type[3], "symmetric") != 0 && strcmp(type[3], "hermitian") != 0 &&
        strcmp(type[3], "skew-symmetric") != 0) return -1;
    return 0;
}

static int parse_size(FILE* f, int* R, int* C, int* NNZ) {
    char line[1025];
    do {
        if (!fgets(line, sizeof line, f)) return -1;
    } while (line[0] == '%');
    if (sscanf(line, "%d %d %d", R, C, NNZ) == 3) return 0;
    for (;;) {
        int got = fscanf(f, "%d %d %d", R, C, NNZ);
        if (got == EOF) return -1;
        if (got == 3) return 0;
    }
}

int cuppen_read_mtx(const char* filename, double** D, double** E, int* n) {
    FILE* f;
    char type[4][64];
    int R, C, NNZ, i;
    if ((f = fopen(filename, "r")) == NULL) {
        fprintf(stderr, "Could not open file\n");
        return CUPPEN_ERR_IO;
    }
    if (parse_banner(f, type) != 0) {
        printf("Could not process Matrix Market banner.\n");
        fclose(f);
        return CUPPEN_ERR_IO;
    }
    if (!(strcmp(type[1], "coordinate") == 0 && strcmp(type[2], "real") == 0 && strcmp(type[3], "general") == 0)) {
        printf("Sorry, this application does not support ");
        printf("Market Market type: [%s %s %s %s]\n", type[0], type[1], type[2], type[3]);
        fclose(f);
        return CUPPEN_ERR_IO;
    }
    if (parse_size(f, &R, &C, &NNZ) != 0) { fclose(f); return CUPPEN_ERR_IO; }
    if (R != C) {
        printf("Matrix is not square\n");
        fclose(f);
        return CUPPEN_ERR_IO;
    }
    if (R < 1) { fclose(f); return CUPPEN_ERR_IO; }
    *n = R;
    *D = (double*)malloc((size_t)R * sizeof(double));
    *E = (double*)malloc((size_t)(R > 1 ? R - 1 : 1) * sizeof(double));
    /* the reference leaves E uninitialised and compares against it when a super-diagonal entry
     * comes first (filehandling.c:137-146); NaN makes that case a deterministic "not symmetric" */
    for (i = 0; i < R; ++i) (*D)[i] = 0.0;
    for (i = 0; i < R - 1; ++i) (*E)[i] = NAN;
    for (i = 0; i < NNZ; ++i) {
        int r, c;
        double v;
        if (fscanf(f, "%d %d %lg\n", &r, &c, &v) != 3) break;
        if (r - c > 1 || c - r > 1) {
            printf("Matrix is not tridiagonal\n");
            goto fail;
        }
        if (r < 1 || c < 1 || r > R || c > R) goto fail;
        if (r == c) (*D)[r - 1] = v;
        else if (c == r + 1) {
            if ((*E)[r - 1] != v) {
                printf("Matrix is not symmetric\n");
                goto fail;
            }
        } else (*E)[c - 1] = v;
    }
    /* a sub-diagonal entry that the file never listed is a zero of the matrix (the reference would read
     * uninitialised memory here): report it as the zero it is, so that the caller's `E[i] != 0` check
     * (src/main.c:196-200) fires instead of a NaN reaching the solver */
    for (i = 0; i < R - 1; ++i)
        if ((*E)[i] != (*E)[i]) (*E)[i] = 0.0;
    fclose(f);
    return CUPPEN_OK;
fail:
    fclose(f);
    free(*D); free(*E);
    *D = *E = NULL;
    return CUPPEN_ERR_IO;
}

static int int_cmp(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

int cuppen_read_ev_file(const char* filename, int n, int** indices, int* count) {
    FILE* f;
    char* line = NULL;
    size_t len = 0;
    int numLines = 0, cap = 16, j = 0;
    int* idx;
    if ((f = fopen(filename, "r")) == NULL) {
        fprintf(stderr, "Could not open file: %s\n", filename);
        return CUPPEN_ERR_IO;
    }
    idx = (int*)malloc((size_t)cap * sizeof(int));
    while (getline(&line, &len, f) != -1) {
        int curr = atoi(line);
        if (curr == 0 || curr > n) {
            size_t L = strlen(line);
            if (L > 0) line[L - 1] = '\0';
            printf("WARNING: Line %d (\"%s\") in file %s will be ignored. No valid eigenvector index for given problem.\n",
                   numLines, line, filename);
        } else {
            numLines++;
            if (curr > 0) {                /* negative values are counted but never stored, as in the reference */
                if (j == cap) { cap *= 2; idx = (int*)realloc(idx, (size_t)cap * sizeof(int)); }
                idx[j++] = curr - 1;
            }
        }
    }
    fclose(f);
    free(line);
    qsort(idx, (size_t)j, sizeof(int), int_cmp);
    *indices = idx;
    *count = j;
    return CUPPEN_OK;
}

int cuppen_write_results(const char* filename, int n, const double* lambda, const double* resid, int all_vectors,
                         const int* indices, int count) {
    FILE* f;
    int iter, iterEV = 0;
    if ((f = fopen(filename, "w")) == NULL) {
        fprintf(stderr, "Could not open file\n");
        return CUPPEN_ERR_IO;
    }
    for (iter = 0; iter < n; ++iter) {
        int computeCurrEV = 0;
        if (all_vectors) computeCurrEV = 1;
        else if (count > 0) {
            while (iterEV < count && indices[iterEV] < iter) iterEV++;
            if (iterEV < count && indices[iterEV] == iter) computeCurrEV = 1;
        }
        if (computeCurrEV && resid) fprintf(f, "%20.19g %20.19g\n", lambda[iter], resid[iter]);
        else fprintf(f, "%20.19g\n", lambda[iter]);
    }
    fclose(f);
    return CUPPEN_OK;
}
