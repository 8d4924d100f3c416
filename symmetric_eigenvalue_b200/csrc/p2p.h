// Peer-memory exchange between the GPUs of one box (NVLink 5 / NVSwitch): the O(n) vectors of the cooperative
// merges and the one-off row redistribution travel as plain loads / stores on peer pointers issued by OUR kernels,
// ordered by a flag barrier in peer memory -- no collective library call inside a solve, so the whole multi-GPU
// decomposition is a fixed kernel sequence that a CUDA graph can replay.  Replaces the blocking MPI_Send/Recv/Bcast
// traffic of the reference (/root/reference/src/main.c:397-417,504-542, src/filehandling.c:347-348,415-437).
//
// Every rank allocates one "symmetric heap" with the same layout (lam, first/last rows, (tau, origin), halo rows,
// residual partial sums, barrier flags) and maps every peer's heap through CUDA IPC; the address of a peer's copy of a
// heap buffer is peer_base + (local pointer - local base).  NCCL is only the bootstrap that ships the IPC handles.
//   push  a producer kernel stores its results into every peer's copy as well as its own (secular roots, boundary
//         rows, halo rows, residual partial sums);
//   pull  the row redistribution reads the peers' Q blocks directly into the slice layout (no staging copies);
//   p2p_barrier_kernel  release-store of a monotonically increasing epoch into every peer's flag array, acquire-spin
//         on the own array; a watchdog turns a lost peer into an error code instead of a hung GPU.
#ifndef CUPPEN_P2P_H
#define CUPPEN_P2P_H

#include "platform.h"

namespace cuppen {

enum { P2P_MAX = 8 };

struct SymHeap {
    char* base[P2P_MAX];     // heap of every rank as mapped into this process; base[me] is the local one
    int me = 0, G = 0;       // G == 0: no peers (one GPU, or the collective back end)
    template <class T>
    CUPPEN_HD T* at(int r, T* local) const { return (T*)(base[r] + ((char*)local - base[me])); }
};

// transition to the cooperative phase: replicate my subtree's eigenvalues and boundary rows on every peer
struct PushSubtreeVectors {
    SymHeap H;
    double *lam, *frow, *lrow;
    int lo;
    CUPPEN_HD void operator()(long i) const {
        const long g = lo + i;
        const double a = lam[g], b = frow[g], c = lrow[g];
        for (int r = 0; r < H.G; ++r) {
            if (r == H.me) continue;
            H.at(r, lam)[g] = a; H.at(r, frow)[g] = b; H.at(r, lrow)[g] = c;
        }
    }
};

// residual partial sums of this rank's row slices -> slot `me` of every rank's [G][n] table
struct PushResidualPartials {
    SymHeap H;
    const double* res2;
    double* part;            // heap: [G][n]
    int n;
    CUPPEN_HD void operator()(long c) const {
        const double v = res2[c];
        for (int r = 0; r < H.G; ++r) H.at(r, part)[(long)H.me * n + c] = v;
    }
};
struct SumResidualPartials {
    const double* part;
    double* res2;
    int n, G;
    CUPPEN_HD void operator()(long c) const {
        double s = 0;
        for (int r = 0; r < G; ++r) s += part[(long)r * n + c];      // fixed order: identical on every rank
        res2[c] = s;
    }
};

// Halo rows of the tridiagonal stencil in the slice layout: my first row of slice s is the row below the last row of
// (rank me-1, slice s) -- or of (rank G-1, slice s-1) when me == 0 --, my last row is the row above the first row of
// (me+1, s) or (0, s+1).  Each value is stored straight into the halo table of the rank that needs it.
struct HaloCtx {
    SymHeap H;
    const double* Q;
    long ldq;
    const int* perm;
    double* halo_lo;         // heap: [S][n]  row above the first row of my slice s, per output column
    double* halo_hi;         // heap: [S][n]  row below the last row of my slice s
    int n, S;                // S subtrees, one slice of each per rank
    int crow0[P2P_MAX + 1];
};
struct PushHaloRows {
    HaloCtx h;
    CUPPEN_HD void operator()(long t) const {
        const int G = h.H.G, me = h.H.me;
        const int s = (int)(t / h.n), c = (int)(t % h.n);
        const long col = (long)h.perm[c] * h.ldq;
        // my first row of slice s -> "hi" halo of the slice that ends just above it
        int rt = me - 1, st = s;
        if (me == 0) { rt = G - 1; st = s - 1; }
        if (st >= 0) h.H.at(rt, h.halo_hi)[(long)st * h.n + c] = h.Q[col + h.crow0[s]];
        // my last row of slice s -> "lo" halo of the slice that starts just below it
        rt = me + 1; st = s;
        if (me == G - 1) { rt = 0; st = s + 1; }
        if (st < h.S) h.H.at(rt, h.halo_lo)[(long)st * h.n + c] = h.Q[col + h.crow0[s + 1] - 1];
    }
};

#if CUPPEN_CUDA
// One block, one thread per peer.  `flags` (heap, [P2P_MAX]) is written by the peers, `epoch` is this rank's barrier
// counter (device memory, so that a replayed CUDA graph keeps counting).  All ranks run the same barrier sequence.
#define P2P_TIMEOUT_NS 30000000000ull      // a peer that has not arrived after 30 s of device time is lost
__device__ __forceinline__ unsigned long long p2p_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void p2p_barrier_kernel(SymHeap H, unsigned* flags, unsigned* epoch, int* fail) {
    __shared__ unsigned ep_s;
    if (threadIdx.x == 0) { ep_s = *epoch + 1; *epoch = ep_s; }
    __syncthreads();
    const int r = threadIdx.x;
    if (r >= H.G) return;
    const unsigned ep = ep_s;
    __threadfence_system();
    unsigned* remote = H.at(r, flags) + H.me;
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" :: "l"(remote), "r"(ep) : "memory");
    const unsigned* mine = flags + r;
    unsigned v = 0;
    unsigned spins = 0;
    const unsigned long long t0 = p2p_now_ns();
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - ep) >= 0) break;
        if ((++spins & 1023u) == 0 && (*(volatile int*)fail || p2p_now_ns() - t0 > P2P_TIMEOUT_NS)) { atomicCAS(fail, 0, 1 + r); break; }
        __nanosleep(spins < 64 ? 32 : 256);
    }
    __threadfence_system();
}

// Row redistribution (layout L -> slice layout C) as one pull kernel: column g of subtree s comes from the Q block of
// the rank that owns subtree s, rows [slo[s], slo[s] + len), and lands at local rows crow0[s].. of the destination buffer.
struct PullCtx {
    const double* src[P2P_MAX];   // Q buffer of the rank that owns subtree s (layout L)
    long src_ld[P2P_MAX];         // ... and its leading dimension (the ranks' row counts differ when G does not divide n)
    double* dst;
    long ldq;
    int S;                        // subtrees (>= ranks)
    int sub_off[P2P_MAX + 1];     // global column range of subtree s
    int slo[P2P_MAX];             // owner-local first row of the slice of subtree s this rank takes
    int crow0[P2P_MAX + 1];       // local row range of slice s in layout C
};
__global__ void __launch_bounds__(256) p2p_pull_rows_kernel(PullCtx p) {
    const int g = blockIdx.x;
    int s = 0;
    while (s + 1 < p.S && g >= p.sub_off[s + 1]) ++s;
    const int len = p.crow0[s + 1] - p.crow0[s];
    const double* src = p.src[s] + (long)g * p.src_ld[s] + p.slo[s];
    double* dst = p.dst + (long)g * p.ldq + p.crow0[s];
    for (int r = blockIdx.y * 256 + threadIdx.x; r < len; r += gridDim.y * 256) dst[r] = src[r];
}
#endif  // CUPPEN_CUDA

}  // namespace cuppen
#endif
