/*
 * cuppens -- drop-in command line front end, host side in plain C over the C ABI of
 * libcuppen_b200.so.  Mirrors the reference driver /root/reference/src/main.c:
 *   options  getopt "hi:n:s:e::" (main.c:99-137), defaults -s 1 -n 1000 (main.c:92-93),
 *            at most one positional argument = output file (main.c:140-145)
 *   stdout   the progress / timer lines of main.c:149-163,193-194,224,241,329,435,485,675-678,
 *            683-684,695 and filehandling.c:566-568
 *   exit     1 bad option, 2 unreadable input, 3 output / -e file, 4 "Leaf Size is too small"
 * Additions that do not collide with the reference's options:
 *   -p P     number of reference MPI tasks whose divide tree is reproduced (what `mpirun -n P`
 *            was; default 1 or $CUPPENS_NUMTASKS)
 *   -g G     number of B200s, any count up to 8 (one process per GPU is forked; the O(n) vectors and the row
 *            redistribution travel through peer memory over NVLink, NCCL ships the IPC handles)
 *   -v FILE  write the computed eigenvectors (all with -e, the selected ones with -eFILE) to FILE
 *            (binary CUPPENV1 layout, include/cuppen_b200.h) -- the reference cannot emit V
 *   -c       print max|V^T V - I| of the computed eigenvectors (evaluated on the GPUs; needs -e)
 * With -eFILE and few requested indices (count <= n/16) the library's selected-eigenvector
 * mode is used: no n x n matrix is formed (filehandling.c:339-345 computes one vector at a time too).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include "cuppen_b200.h"

static void showHelp(void) {
    printf("\n");
    printf("USAGE cuppens [options] [outputfile]\n");
    printf("\n");
    printf("The program can compute all the eigenpairs of a matrix on a parallel machine\n");
    printf("by using cuppens algorithm\n");
    printf("The results can be written into an outputfile, if specified.\n");
    printf("\n");
    printf("OPTIONS\n");
    printf(" -h\n");
    printf("    Show help.\n");
    printf(" -i FILENAME\n");
    printf("    The name of a file which contains a tridiagonal matrix in mtx format.\n");
    printf("    The eigenvalues of this matrix will then be computed.\n");
    printf(" -s NUM\n");
    printf("    If you want to compute the eigenvalues of a predefined matrix, you may\n");
    printf("    use this option to define the scheme of the matrix.\n");
    printf("    1 - Matrix will have the tridiagonal form [-1,d_i,-1] where the diagonal\n");
    printf("        elements will be evenly spaced in the interval [1,100] \n");
    printf("    2 - Eigenvalue i has the form: 2 + 2*cos((PI*i)/(n+1)) \n");
    printf("        Poisson-matrix (tridiagonal form of [-1,2-1])\n");
    printf("    If option i is used, then this option will be ignored.\n");
    printf(" -n NUM\n");
    printf("    Specify the dimension of the matrix chosen with option -s.\n");
    printf(" -e(FILENAME)\n");
    printf("    Without this option, no eigenvectors are computed, just the eigenvalues.\n");
    printf("    If you just specify the flag -e, then all eigenvectors will be computed.\n");
    printf("    If you specify additionally a filename, then it will read the indices\n");
    printf("    of the eigenvectors to compute from this file (each line one index).\n");
    printf("    Note, there is no blank between the option and the filename.\n");
    printf(" -p NUM\n");
    printf("    Number of tasks of the original MPI program whose divide tree and deflation\n");
    printf("    rules are reproduced (default 1: accurate tolerances on all levels).\n");
    printf(" -g NUM\n");
    printf("    Number of GPUs (any count up to 8; one process per GPU).\n");
    printf(" -v FILENAME\n");
    printf("    Write the computed eigenvectors to this file (binary; needs -e).\n");
    printf(" -c\n");
    printf("    Check the orthogonality of the computed eigenvectors (needs -e).\n");
    printf("\n");
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void xwrite(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n) { ssize_t w = write(fd, c, n); if (w <= 0) _exit(90); c += w; n -= (size_t)w; }
}
static void xread(int fd, void* p, size_t n) {
    char* c = (char*)p;
    while (n) { ssize_t r = read(fd, c, n); if (r <= 0) _exit(91); c += r; n -= (size_t)r; }
}

int main(int argc, char** argv) {
    int n = 1000, usedScheme = 1, computeEV = 0, writeOutput = 0, numtasks = 1, gpus = 1;
    char *inputfile = NULL, *outputfile = NULL, *evFile = NULL, *vecFile = NULL;
    int checkOrth = 0;
    double *D = NULL, *E = NULL;
    int c, i, rc;
    const char* envp = getenv("CUPPENS_NUMTASKS");
    if (envp && atoi(envp) > 0) numtasks = atoi(envp);

    if (argc == 1) { showHelp(); return 0; }
    opterr = 0;
    while ((c = getopt(argc, argv, "hi:n:s:e::p:g:v:c")) != -1) switch (c) {
        case 'h': showHelp(); return 0;
        case 'i': inputfile = optarg; break;
        case 's':
            usedScheme = atoi(optarg);
            if (usedScheme < 1 || usedScheme > 2) { fprintf(stderr, "Invalid argument for option -s. See help.\n"); return 1; }
            break;
        case 'n':
            n = atoi(optarg);
            if (n < 1) { fprintf(stderr, "Invalid argument for option -n. See help.\n"); return 1; }
            break;
        case 'e': computeEV = 1; if (optarg) evFile = optarg; break;
        case 'p':
            numtasks = atoi(optarg);
            if (numtasks < 1) { fprintf(stderr, "Invalid argument for option -p. See help.\n"); return 1; }
            break;
        case 'v': vecFile = optarg; break;
        case 'c': checkOrth = 1; break;
        case 'g':
            gpus = atoi(optarg);
            if (gpus < 1 || gpus > 16) { fprintf(stderr, "Invalid argument for option -g. See help.\n"); return 1; }
            break;
        case '?':
            if (isprint(optopt)) fprintf(stderr, "Unknown option `-%c'.\n", optopt);
            else fprintf(stderr, "Unknown option character `\\x%x'.\n", optopt);
            return 1;
        default: return 1;
    }
    if (argc - optind > 1) { fprintf(stderr, "Invalid number of positional arguments. See help.\n"); return 1; }
    outputfile = argv[optind];

    if (inputfile != NULL) printf("Input file: %s\n", inputfile);
    else printf("Use a matrix of scheme %d with dimension %d\n", usedScheme, n);
    if (computeEV) {
        if (evFile != NULL) printf("Compute the eigenvectors defined in: %s\n", evFile);
        else printf("Program will compute all eigenvectors\n");
    }
    if (outputfile != NULL) { writeOutput = 1; printf("Output file: %s\n", outputfile); }

    if (inputfile != NULL) {
        if (cuppen_read_mtx(inputfile, &D, &E, &n) != 0) return 2;
    } else {
        D = (double*)malloc((size_t)n * sizeof(double));
        E = (double*)malloc((size_t)(n > 1 ? n - 1 : 1) * sizeof(double));
        cuppen_scheme(usedScheme, n, D, E);
    }
    printf("\n");
    printf("Number of MPI tasks is: %d\n", numtasks);
    /* entries that are not finite (a literal nan / inf in the input file) cannot be decomposed */
    for (i = 0; i < n; ++i)
        if (!isfinite(D[i]) || (i < n - 1 && !isfinite(E[i]))) {
            printf("Matrix contains an entry that is not finite\n");
            return 2;
        }
    /* the reference asserts on zero entries (main.c:196-200) */
    for (i = 0; i < n; ++i)
        if (D[i] == 0 || (i < n - 1 && E[i] == 0)) {
            fprintf(stderr, "cuppens: Assertion `%s[i] != 0' failed.\n", D[i] == 0 ? "D" : "E");
            abort();
        }
    if (n / numtasks == 0) { fprintf(stderr, "Leaf Size is too small! Reduce number of tasks.\n"); return 4; }
    fflush(stdout);

    /* one process per GPU: fork before any CUDA call; rank 0 is this process */
    int rank = 0;
    int up[16][2], down[16][2];
    pid_t kids[16];
    unsigned char id[CUPPEN_NCCL_ID_BYTES];
    if (gpus > 16) { fprintf(stderr, "Invalid argument for option -g. See help.\n"); return 1; }
    for (i = 1; i < gpus; ++i) { if (pipe(up[i]) || pipe(down[i])) return 1; }
    for (i = 1; i < gpus; ++i) {
        pid_t p = fork();
        if (p < 0) return 1;
        if (p == 0) { rank = i; break; }
        kids[i] = p;
    }
    char hostname[256];
    if (gethostname(hostname, sizeof hostname) != 0) strcpy(hostname, "localhost");
    if (rank == 0)
        for (i = 0; i < gpus; ++i)
            printf("   Task %d is running on node %s, which has %ld available processors.\n", i, hostname,
                   sysconf(_SC_NPROCESSORS_ONLN));

    double tic = now_s();
    if (rank == 0) printf("Start divide phase ...\n");
    cuppen_handle h = NULL;
    int vectors = (computeEV && writeOutput) ? CUPPEN_FLAG_VECTORS : 0;
    /* -eFILE: look at the index file now (quietly -- its WARNING lines belong to the write phase, where the
     * reference parses it, filehandling.c:339) to decide between the full and the selected-eigenvector mode */
    int selectMode = 0, selCount = 0;
    int* selIdx = NULL;
    if (vectors && evFile != NULL) {
        fflush(stdout);
        int keep = dup(1), nul = open("/dev/null", O_WRONLY);
        if (keep >= 0 && nul >= 0) {
            dup2(nul, 1);
            int ok = cuppen_read_ev_file(evFile, n, &selIdx, &selCount);
            fflush(stdout);
            dup2(keep, 1);
            if (ok == 0 && selCount == 0) vectors = 0;                       /* nothing valid requested */
            else if (ok == 0 && selCount <= n / 16 && !checkOrth) { selectMode = 1; vectors = 0; }
        }
        if (keep >= 0) close(keep);
        if (nul >= 0) close(nul);
    }
    const int createFlags = selectMode ? CUPPEN_FLAG_SELECT : vectors;
    if (gpus > 1) {
        if (rank == 0) {
            if (cuppen_nccl_unique_id(id) != 0) { fprintf(stderr, "%s\n", cuppen_last_error()); return 5; }
            for (i = 1; i < gpus; ++i) xwrite(down[i][1], id, sizeof id);
        } else xread(down[rank][0], id, sizeof id);
        rc = cuppen_create_nccl(&h, n, numtasks, createFlags, rank, rank, gpus, id);
    } else rc = cuppen_create(&h, n, numtasks, createFlags, 0);
    if (rc == CUPPEN_ERR_LEAF) { fprintf(stderr, "Leaf Size is too small! Reduce number of tasks.\n"); return 4; }
    if (rc != 0) { fprintf(stderr, "cuppens: %s\n", cuppen_last_error()); return 5; }
    if (rank == 0) printf("Average leaf size will be %.1lf\n", n * 1.0 / numtasks);
    rc = cuppen_set_tridiagonal(h, D, E);
    if (rc == 0 && selectMode) rc = cuppen_select_eigenvectors(h, selIdx, selCount);
    if (rc != 0) { fprintf(stderr, "cuppens: %s\n", cuppen_last_error()); return 5; }
    if (rank == 0) {
        printf("Apply QR algorithm on leaves ...\n");
        if (numtasks > 1) printf("Start Conquer Phase ...\n");
        fflush(stdout);
    }
    rc = cuppen_solve(h);
    if (rc != 0) { fprintf(stderr, "cuppens: %s\n", cuppen_last_error()); return 5; }
    double toc = now_s();

    if (rank != 0) {
        /* the eigenvector file is written by rank 0 from row slices that every rank contributes */
        if (vectors && vecFile != NULL) cuppen_write_eigenvectors(h, NULL);
        if (vectors && checkOrth) { double dev = 0, sec = 0; cuppen_orthogonality(h, &dev, &sec); }
        cuppen_destroy(h);
        _exit(0);
    }

    cuppen_timers tm;
    cuppen_get_timers(h, &tm);
    double elapsed = toc - tic;
    /* with -e the back-transformation runs inside the conquer loop; report it separately as the reference does */
    double evalTime = (vectors || selectMode) ? elapsed - tm.backtransform_s : elapsed;
    if (evalTime <= 0) evalTime = elapsed;
    printf("\n");
    printf("Required time to compute all eigenvalues: %f seconds\n", evalTime);
    printf("Required time for root finding: %f seconds; fraction: %.1f%%\n", tm.root_finding_s, 100 * tm.root_finding_s / evalTime);
    printf("Required time for eigenvector extraction from U_i's: %f seconds; fraction: %.1f%%\n",
           tm.ev_extract_s - tm.backtransform_ev_s, 100 * (tm.ev_extract_s - tm.backtransform_ev_s) / evalTime);

    if (writeOutput) {
        printf("\n");
        printf("Write results to file ...\n");
        double* lambda = (double*)malloc((size_t)n * sizeof(double));
        double* resid = NULL;
        int* indices = NULL;
        int count = 0;
        cuppen_get_eigenvalues(h, lambda);
        FILE* probe = fopen(outputfile, "w");
        if (probe == NULL) { fprintf(stderr, "Could not open file\n"); return 3; }
        fclose(probe);
        if (computeEV && evFile != NULL) {
            if (cuppen_read_ev_file(evFile, n, &indices, &count) != 0) return 3;
        }
        if (vectors || selectMode) {
            resid = (double*)malloc((size_t)n * sizeof(double));
            if (cuppen_get_residuals(h, NULL, n, resid) != 0) { fprintf(stderr, "cuppens: %s\n", cuppen_last_error()); return 5; }
        }
        if (cuppen_write_results(outputfile, n, lambda, resid, computeEV && evFile == NULL, indices, count) != 0) return 3;
        if (computeEV && (evFile == NULL || count > 0)) {
            printf("\n");
            printf("Required time for backtransformation: %f seconds\n", tm.backtransform_s);
            printf("Required time eigenvector extraction from U_i's within backtransformation: %f seconds; fraction: %.1f%%\n",
                   tm.backtransform_ev_s, tm.backtransform_s > 0 ? 100 * tm.backtransform_ev_s / tm.backtransform_s : 0.0);
        }
        free(lambda); free(resid); free(indices); free(selIdx);
    }
    if ((vectors || selectMode) && vecFile != NULL) {
        if (cuppen_write_eigenvectors(h, vecFile) != 0) { fprintf(stderr, "cuppens: %s\n", cuppen_last_error()); return 3; }
        printf("\nEigenvectors written to: %s\n", vecFile);
    }
    if (vectors && checkOrth) {
        double dev = 0, sec = 0;
        if (cuppen_orthogonality(h, &dev, &sec) != 0) { fprintf(stderr, "cuppens: %s\n", cuppen_last_error()); return 5; }
        printf("\nOrthogonality max|V^T V - I|: %.3e (checked in %f seconds)\n", dev, sec);
    }
    cuppen_destroy(h);
    for (i = 1; i < gpus; ++i) { int st; waitpid(kids[i], &st, 0); }
    printf("\nProgram finished successfully!\n");
    free(D); free(E);
    return 0;
}
