/*
 * TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product.
 * fork()+socketpair implementation of the MPI subset declared in shim/mpi.h.
 * Blocking tagged point-to-point with an unexpected-message queue; Bcast and
 * Barrier are built on top of it.  Rank 0 is the process the user started.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#define SHIM_MAXP 64
#define TAG_BCAST (-101)
#define TAG_BARRIER (-102)

typedef struct Msg { int tag; int nbytes; char *data; struct Msg *next; } Msg;

static int g_np = 1, g_rank = 0;
static int g_fd[SHIM_MAXP];          /* socket to each peer */
static Msg *g_queue[SHIM_MAXP];      /* unexpected messages per source */
static pid_t *g_pids = NULL;         /* shared pid table */

static void xwrite(int fd, const void *p, size_t n) {
    const char *c = (const char *)p;
    while (n) {
        ssize_t w = write(fd, c, n);
        if (w < 0) { if (errno == EINTR) continue; perror("mpishim write"); _exit(99); }
        c += w; n -= (size_t)w;
    }
}
static void xread(int fd, void *p, size_t n) {
    char *c = (char *)p;
    while (n) {
        ssize_t r = read(fd, c, n);
        if (r < 0) { if (errno == EINTR) continue; perror("mpishim read"); _exit(99); }
        if (r == 0) { fprintf(stderr, "mpishim: rank %d: peer closed\n", g_rank); _exit(98); }
        c += r; n -= (size_t)r;
    }
}

int MPI_Init(int *argc, char ***argv) {
    (void)argc; (void)argv;
    const char *e = getenv("MPISHIM_NP");
    g_np = e ? atoi(e) : 1;
    if (g_np < 1 || g_np > SHIM_MAXP) { fprintf(stderr, "mpishim: bad MPISHIM_NP\n"); exit(97); }
    g_pids = mmap(NULL, sizeof(pid_t) * SHIM_MAXP, PROT_READ | PROT_WRITE,
                  MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset(g_pids, 0, sizeof(pid_t) * SHIM_MAXP);
    g_pids[0] = getpid();
    static int sv[SHIM_MAXP][SHIM_MAXP][2];
    for (int i = 0; i < g_np; ++i)
        for (int j = i + 1; j < g_np; ++j)
            if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv[i][j]) != 0) { perror("socketpair"); exit(96); }
    fflush(stdout); fflush(stderr);
    g_rank = 0;
    for (int r = 1; r < g_np; ++r) {
        pid_t p = fork();
        if (p < 0) { perror("fork"); exit(95); }
        if (p == 0) { g_rank = r; break; }
        g_pids[r] = p;
    }
    for (int i = 0; i < g_np; ++i)
        for (int j = i + 1; j < g_np; ++j) {
            if (i == g_rank) { g_fd[j] = sv[i][j][0]; close(sv[i][j][1]); }
            else if (j == g_rank) { g_fd[i] = sv[i][j][1]; close(sv[i][j][0]); }
            else { close(sv[i][j][0]); close(sv[i][j][1]); }
        }
    return MPI_SUCCESS;
}

int MPI_Finalize(void) {
    fflush(stdout); fflush(stderr);
    if (g_rank == 0) {
        for (int r = 1; r < g_np; ++r) { int st; if (g_pids[r] > 0) waitpid(g_pids[r], &st, 0); }
    } else {
        _exit(0);   /* children never return into the caller's atexit chain twice */
    }
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode) {
    (void)comm;
    fflush(stdout); fflush(stderr);
    for (int r = 0; r < g_np; ++r)
        if (r != g_rank && g_pids && g_pids[r] > 0) kill(g_pids[r], SIGKILL);
    _exit(errorcode);
}

int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = g_np; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Get_processor_name(char *name, int *resultlen) {
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) strcpy(name, "localhost");
    name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
    *resultlen = (int)strlen(name);
    return MPI_SUCCESS;
}

static void send_raw(int dest, int tag, const void *buf, int nbytes) {
    int hdr[2] = { tag, nbytes };
    xwrite(g_fd[dest], hdr, sizeof hdr);
    if (nbytes) xwrite(g_fd[dest], buf, (size_t)nbytes);
}
static void recv_raw(int src, int tag, void *buf, int nbytes) {
    Msg **pp = &g_queue[src];
    for (; *pp; pp = &(*pp)->next)
        if ((*pp)->tag == tag) {
            Msg *m = *pp; *pp = m->next;
            if (m->nbytes != nbytes) { fprintf(stderr, "mpishim: size mismatch tag %d\n", tag); _exit(94); }
            memcpy(buf, m->data, (size_t)nbytes); free(m->data); free(m); return;
        }
    for (;;) {
        int hdr[2];
        xread(g_fd[src], hdr, sizeof hdr);
        if (hdr[0] == tag) {
            if (hdr[1] != nbytes) { fprintf(stderr, "mpishim: size mismatch tag %d (%d vs %d)\n", tag, hdr[1], nbytes); _exit(94); }
            if (nbytes) xread(g_fd[src], buf, (size_t)nbytes);
            return;
        }
        Msg *m = malloc(sizeof *m);
        m->tag = hdr[0]; m->nbytes = hdr[1]; m->data = malloc(hdr[1] ? (size_t)hdr[1] : 1); m->next = NULL;
        if (hdr[1]) xread(g_fd[src], m->data, (size_t)hdr[1]);
        Msg **q = &g_queue[src]; while (*q) q = &(*q)->next; *q = m;
    }
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) {
    (void)comm; send_raw(dest, tag, buf, count * dt); return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st) {
    (void)comm; recv_raw(src, tag, buf, count * dt);
    if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
    return MPI_SUCCESS;
}
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm) {
    (void)comm;
    if (g_rank == root) { for (int r = 0; r < g_np; ++r) if (r != root) send_raw(r, TAG_BCAST, buf, count * dt); }
    else recv_raw(root, TAG_BCAST, buf, count * dt);
    return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm comm) {
    (void)comm; char c = 0;
    if (g_rank == 0) {
        for (int r = 1; r < g_np; ++r) recv_raw(r, TAG_BARRIER, &c, 1);
        for (int r = 1; r < g_np; ++r) send_raw(r, TAG_BARRIER, &c, 1);
    } else { send_raw(0, TAG_BARRIER, &c, 1); recv_raw(0, TAG_BARRIER, &c, 1); }
    return MPI_SUCCESS;
}
