// Named CUDA kernels of the hot path (one `__global__` each, so that they show up under their
// own names in ncu) and their TEST-ONLY host twins (CUPPEN_HOST_EMULATION, see platform.h):
//
//   leaf_ql_kernel      K0  warp-per-leaf implicit QL  (replaces LAPACKE_dsteqr, src/main.c:460)
//   secular_kernel      K3  warp-per-root secular solver (replaces the bisection, src/eigenvalues.c:161-247)
//   pack_kernel         K5a rotated / deflated column moves (inverse Givens of getEigenVector,
//                           src/eigenvalues.c:343-357, applied to the columns of Q instead of per vector)
//   ugen_kernel         K5b eigenvector matrix of the rank-one update (src/eigenvalues.c:316-321 with
//                           the Loewner z-hat and cancellation-free differences)
//   dgemm_dmma_kernel   K6  FP64 tensor-core GEMM  Q' = diag(Q1,Q2) U   (see gemm_dmma.h)
//   residual_kernel     K8  ||T x - lambda x||_2 per column (src/filehandling.c:511-531)
//   gather_cols_kernel      column permutation into ascending-lambda order (src/filehandling.c:315-321)
#ifndef CUPPEN_MATRIX_STAGES_H
#define CUPPEN_MATRIX_STAGES_H

#include "merge_stages.h"
#include "secular_core.h"
#include "gemm_dmma.h"
#include "p2p.h"
#if CUPPEN_CUDA
#include <cooperative_groups.h>
#endif

namespace cuppen {

enum { K_PAD = 32 };            // the GEMM reads K in multiples of this; A is zero-padded up to it
enum { LEAF_MAX = 32 };         // largest leaf handled by one warp

struct LeafDesc { int off, n; };

// Row support of a column of the block-diagonal eigenvector matrix, in GLOBAL rows.  A column is non-zero only inside
// the block of the last merge at which it was not z-deflated (a leaf block if never): the leaf kernel writes the leaf
// spans, pack_kernel widens the span of every column that a merge rewrites (live roots and rotated columns), z-deflated
// columns keep theirs.  The residual kernel reads only the rows lo-1 .. hi of a column -- on heavily deflating matrices
// (Wilkinson n=16384: a few hundred live columns at the top merges) that is a small part of the n rows.
struct RowSpan { int lo, hi; };

struct MatCtx {
    int n;                // global size
    long ldq;             // leading dimension of the Q buffers and of Apack (local rows, padded)
    const double* Qold;   // children (block diagonal), column-major, local rows
    double* Qnew;         // parents; the solver works in place (Qnew == Qold): every kernel reads a column before writing it
    double* Apack;        // packed live columns (K order), same shape as Q plus K_PAD columns
    double* B;            // U arena, row-major [n + pad][ldb]
    long ldb;
    RowSpan* span;        // per storage column: global rows outside [lo, hi) are exactly zero (nullptr: not tracked)
};

// GEMM work list of one level and one panel, built on the device from the merge descriptors so that
// the host never has to read the deflation counts back between the stages of a level.
struct WorkCtx {
    const MergeDesc* desc;
    int nd;                 // merges of this level
    int p0, width;          // panel of root columns [p0, p0+width)
    int BM, BN;             // CTA tile of the GEMM kernel that will consume the list
    long ldq, ldb;
    double* Apack;
    double* B;
    double* Qnext;
    const int* lidx;
    GemmProblem* probs;     // [2*nd]
    GemmTile* tiles;
    int* ntiles;            // [0] tile count, [1] number of problems with an odd first row (the TMA kernels fetch those from one row earlier)
    int tile_cap;
    int* fail;              // set to 1 when the level needs more than tile_cap tiles (the solve then reports an error)
    int supercol_mb;        // L2 budget (MiB) of the B panel of one super-column (work_supercol)
    int split_grid;         // > 0: persistent CTAs of the consuming kernel, which accepts half tiles (tensor-map TMA GEMM): the
                            // tiles of an under-filled last wave are emitted as two 64-column halves each (build_gemm_work_body)
};

// Tile order of one problem: super-columns of `nsw` n-tiles, all m-tiles inside a super-column, n fastest -- so B is
// read from DRAM about once and A once per super-column instead of B once per m-tile.  nsw: the B panel (K x nsw*BN
// doubles) takes at most `supercol_mb` MiB of the L2.  Measured with ncu on the top merge of GOE n=16384 (4.3 GB of
// operand bytes, 81 ms, DMMA-bound either way; profiles/README.md, r02 calls I and J): DRAM read + write
//   48 MiB, no hints 24.7 + 1.7 GB | 96 MiB, no hints 27.0 + 1.7 (the panel does not stay resident: two L2 partitions, C
//   write-allocates) | 48 MiB, B evict_last + streaming C stores 22.1 + 1.7 | 72 MiB with the same hints 20.9 + 1.7
//   | A evict_first on top 32.3 + 1.7 (the A strip is shared by the n-tiles in flight and must survive them).
// The hints cost 0.1 % of GEMM time in every run (112.81 -> 112.91 ms per solve; the kernel is DMMA-bound, DRAM at 0.3 TB/s),
// so the default stays 48 MiB without hints; CUPPEN_GEMM_HINT=1 CUPPEN_SUPERCOL_MB=72 selects the low-traffic variant.
// The floor of this order is |B| + |C| + (N / nsw BN) |A| = 17 GB at 48 MiB: A is streamed once per super-column.
CUPPEN_HD int work_supercol(const WorkCtx& w, const GemmProblem& Pb) {
    const int ntn = (Pb.N + w.BN - 1) / w.BN;
    long nsw = ((long)w.supercol_mb << 20) / ((long)(Pb.K > 0 ? Pb.K : 1) * 8 * w.BN);
    if (nsw < 1) nsw = 1;
    return nsw < ntn ? (int)nsw : ntn;
}
// the q-th tile of problem p in that order (closed form, so that the tiles of a problem can be written in parallel)
CUPPEN_HD GemmTile work_tile_at(const WorkCtx& w, const GemmProblem& Pb, int p, int q) {
    const int ntm = (Pb.M + w.BM - 1) / w.BM, ntn = (Pb.N + w.BN - 1) / w.BN;
    const int nsw = work_supercol(w, Pb);
    const int full = ntm * nsw;                      // tiles of a full super-column (only the last one can be narrower)
    const int sc = q / full, rem = q - sc * full;
    const int ns = sc * nsw;
    const int wd = (ntn - ns) < nsw ? (ntn - ns) : nsw;
    const int mt = rem / wd, nt = ns + (rem - mt * wd);
    return GemmTile{p, mt * w.BM, nt * w.BN};
}
// Wave quantisation: the persistent kernel takes the tiles round-robin, `split_grid` at a time; when the last wave fills at
// most half of the CTAs, its R tiles are emitted as 2R half tiles (64 columns each), so that the wave takes about half a
// tile time (8 GPUs, top merge of GOE n=16384: 1696 tiles per rank = 11 waves of 148 + 68).  A level with fewer tiles
// than half the CTAs is one short wave: every tile is split.  Returns R (0: no split).
CUPPEN_HD int work_tail_tiles(const WorkCtx& w, int run) {
    int tail = 0;
    if (w.split_grid > 0) {
        tail = run % w.split_grid;
        if (2 * tail > w.split_grid) tail = 0;
    }
    return tail;
}
// entry t of the emitted list (run + tail entries): which tile of the plain order it comes from, and which half
CUPPEN_HD int work_entry_source(int run, int tail, int t, int* half) {
    const int whole = run - tail;
    if (t < whole) { *half = -1; return t; }
    *half = (t - whole) & 1;
    return whole + ((t - whole) >> 1);
}
CUPPEN_HD GemmTile work_entry_tile(GemmTile T, int half) {
    if (half >= 0) { T.prob |= GEMM_TILE_HALF; T.n0 += half * 64; }
    return T;
}

CUPPEN_HD int work_fill_problem(const WorkCtx& w, int p, GemmProblem& Pb) {
    const MergeDesc& D = w.desc[p >> 1];
    const int half = p & 1;
    Pb.M = 0; Pb.N = 0; Pb.K = 0; Pb.c_sub = 0;
    if (D.k <= w.p0) return 0;
    const int hs = half ? D.off + D.n1 : D.off;                 // arena row of the half's first pole
    const int rs = half ? D.lsplit : D.lr0, re = half ? D.lr1 : D.lsplit;   // local rows of the half
    if (re <= rs) return 0;
    Pb.M = re - rs;
    Pb.N = (D.k - w.p0) < w.width ? (D.k - w.p0) : w.width;
    Pb.K = half ? D.kbot : D.ktop;
    Pb.A = w.Apack + rs + (long)D.off * w.ldq; Pb.lda = w.ldq;
    Pb.B = w.B + (long)hs * w.ldb; Pb.ldb = w.ldb;
    Pb.C = w.Qnext + rs + (long)D.off * w.ldq; Pb.ldc = w.ldq;
    Pb.colidx = w.lidx + D.off + w.p0;
    Pb.a_row0 = rs; Pb.a_col0 = D.off; Pb.b_row0 = hs; Pb.b_col0 = 0;
    return ((Pb.M + w.BM - 1) / w.BM) * ((Pb.N + w.BN - 1) / w.BN);
}

#if CUPPEN_CUDA
// one block (any size); work_off: 2*nd ints of shared memory (tile offsets).  Runs as the extra last block of ugen_kernel:
// the work list only needs the descriptors, which are final once the vector front end of the level is done.
__device__ __forceinline__ void build_gemm_work_body(const WorkCtx& w, int* work_off) {
    __shared__ int misaligned;
    const int np = 2 * w.nd;
    if (threadIdx.x == 0) misaligned = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        GemmProblem Pb;
        work_off[p] = work_fill_problem(w, p, Pb);
        w.probs[p] = Pb;
        if (Pb.M > 0 && (Pb.a_row0 & 1)) atomicAdd(&misaligned, 1);
    }
    __syncthreads();
    __shared__ int s_run, s_tail;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int p = 0; p < np; ++p) { int c = work_off[p]; work_off[p] = run; run += c; }
        int tail = work_tail_tiles(w, run);
        if (run + tail > w.tile_cap) { *w.fail = 1; tail = 0; if (run > w.tile_cap) run = w.tile_cap; }
        s_run = run; s_tail = tail;
        w.ntiles[0] = run + tail;
        w.ntiles[1] = misaligned;
    }
    __syncthreads();
    // every thread writes total / blockDim tiles: the problem of a tile index is found by bisection in the prefix sums
    // (one thread per problem took ~45 us per launch at the top levels -- thousands of tiles in two problems --, a
    // loop over the problems ~150 us at the bottom levels -- hundreds of problems; profiles/README.md)
    const int run = s_run, tail = s_tail;
    for (int t = threadIdx.x; t < run + tail; t += blockDim.x) {
        int half;
        const int src = work_entry_source(run, tail, t, &half);           // tile of the plain order that entry t comes from
        int lo = 0, hi = np - 1;                     // last p with work_off[p] <= src (empty problems share their successor's offset)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (work_off[mid] <= src) lo = mid; else hi = mid - 1;
        }
        const GemmProblem& Pb = w.probs[lo];
        w.tiles[t] = work_entry_tile(work_tile_at(w, Pb, lo, src - work_off[lo]), half);
    }
}
__global__ void __launch_bounds__(1024) build_gemm_work_kernel(WorkCtx w) {
    extern __shared__ int work_off_smem[];
    build_gemm_work_body(w, work_off_smem);
}
#endif

#if CUPPEN_CUDA
// ------------------------------------------------------------------------------------------------
// K0: one warp per leaf, lane r owns row r of Q (kept in shared memory).  The scalar QL recurrence
// is executed redundantly by all lanes and every lane stores the identical d/e values, so the warp
// needs no synchronisation inside a sweep; the next pole pair is prefetched ahead of the dependent
// sqrt/divide chain.  The leaf is scaled by a power of two first so that f*f+g*g cannot overflow.
__global__ void __launch_bounds__(128) leaf_ql_kernel(const LeafDesc* __restrict__ leaves, int nleaves,
                                                      const double* __restrict__ Dm, const double* __restrict__ E,
                                                      double* __restrict__ lam, double* __restrict__ frow,
                                                      double* __restrict__ lrow, double* __restrict__ Q, long ldq,
                                                      int R0, int* __restrict__ fail, int compact, RowSpan* __restrict__ span) {
    __shared__ double sq[4][LEAF_MAX][LEAF_MAX + 1];
    __shared__ double sd[4][LEAF_MAX];
    __shared__ double se[4][LEAF_MAX + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * 4 + w;
    if (leaf >= nleaves) return;
    const int off = leaves[leaf].off, nl = leaves[leaf].n;
    double (*q)[LEAF_MAX + 1] = sq[w];
    double* d = sd[w];
    double* e = se[w];
    for (int c = 0; c < nl; ++c) q[lane][c] = (lane == c) ? 1.0 : 0.0;
    double dv = (lane < nl) ? Dm[off + lane] : 0.0;
    double ev = (lane < nl - 1) ? E[off + lane] : 0.0;
    double mx = fmax(fabs(dv), fabs(ev));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int ex = 0;
    if (mx > 0.0) frexp(mx, &ex);
    const double scl = ldexp(1.0, -ex), unscl = ldexp(1.0, ex);
    if (lane < nl) d[lane] = dv * scl;
    if (lane < nl) e[lane] = ev * scl;
    __syncwarp();
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < nl; ++l) {
        int iter = 0, m;
        do {
            __syncwarp();      // all lanes have finished the previous sweep before anybody rescans d/e
            for (m = l; m < nl - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 90) { if (lane == 0) atomicExch(fail, 1 + off + l); break; }
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = sqrt(g * g + 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
                double s = 1.0, c = 1.0, p = 0.0;
                double ei = e[m - 1], di = d[m - 1], dip1 = d[m];
                __syncwarp();  // the shift above read d/e: no lane may start storing before all have read
                bool early = false;
                for (int i = m - 1; i >= l; --i) {
                    const double e_nx = (i > l) ? e[i - 1] : 0.0, d_nx = (i > l) ? d[i - 1] : 0.0;
                    const double f = s * ei, b = c * ei;
                    const double h2 = f * f + g * g;
                    if (h2 == 0.0) {
                        e[i + 1] = 0.0;
                        d[i + 1] = dip1 - p;
                        e[m] = 0.0;
                        early = true;
                        break;
                    }
                    // one reciprocal square root instead of a square root followed by a division: the scalar
                    // recurrence is the serial chain that bounds the kernel
                    const double rinv = rsqrt(h2);
                    r = h2 * rinv;
                    e[i + 1] = r;
                    s = f * rinv;
                    c = g * rinv;
                    g = dip1 - p;
                    r = (di - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    const double f2 = q[lane][i + 1], q0 = q[lane][i];
                    q[lane][i + 1] = s * q0 + c * f2;
                    q[lane][i] = c * q0 - s * f2;
                    dip1 = di; di = d_nx; ei = e_nx;
                }
                if (!early) { d[l] = dip1 - p; e[l] = g; e[m] = 0.0; }   // dip1 holds the old d[l] here
            }
        } while (m != l);
    }
    __syncwarp();
    // ascending order (selection sort; swaps are per-row so every lane swaps its own entries)
    for (int i = 0; i < nl - 1; ++i) {
        int kmin = i;
        double p = d[i];
        const double di0 = p;
        for (int k2 = i + 1; k2 < nl; ++k2) if (d[k2] < p) { kmin = k2; p = d[k2]; }
        __syncwarp();
        if (kmin != i) {
            d[kmin] = di0; d[i] = p;
            double t = q[lane][i]; q[lane][i] = q[lane][kmin]; q[lane][kmin] = t;
        }
        __syncwarp();
    }
    __syncwarp();
    if (lane < nl) {
        lam[off + lane] = d[lane] * unscl;
        frow[off + lane] = q[0][lane];
        lrow[off + lane] = q[nl - 1][lane];
        if (span != nullptr) span[off + lane] = RowSpan{off, off + nl};
    }
    const int grow = off + lane;          // global row of this lane
    if (Q != nullptr && lane < nl)
        for (int c = 0; c < nl; ++c) Q[(long)(grow - R0) + (long)(compact ? c : off + c) * ldq] = q[lane][c];   // compact: (row, c) at row + c*ldq
}

// K3: one warp per secular root.  The poles and weights of the merge are staged in shared memory: all of
// them when they fit (k <= kcap <= SEC_SMEM_K; 64 KB per CTA; the 64 registers per thread allow two CTAs = 32 warps
// per SM -- with the whole 227 KB per CTA the 16 warps of the single resident CTA left the FP64 pipe 32 % active, ncu;
// forcing three CTAs with __launch_bounds__(512, 3) spills and is 12 % slower, profiles/r02_session2_ab_runs.txt call R),
// otherwise chunk by chunk through a CTA-collective evaluator --
// the SEC_WARPS roots of a CTA then iterate in lockstep (a warp whose root has converged keeps taking part in
// the chunk loads and barriers until the slowest root of the CTA is done), so every chunk is read from L2 once
// per CTA and evaluation instead of once per warp.
enum { SEC_SMEM_K = 4096, SEC_WARPS = 16 };

struct SecularStagedEval {
    const double* d;       // global poles / weights of the merge
    const double* w;
    int k, cap;
    double* sm;            // [2][cap] chunk buffer
    // CTA-collective; `math` = false for a drained warp (takes part in loads and barriers only)
    __device__ __forceinline__ SecularSums run(double dorg, double tau, int split, bool math) const {
        const int lane = threadIdx.x & 31;
        double psi = 0, dpsi = 0, phi = 0, dphi = 0;          // (err = |psi| + |phi|, see secular_eval)
        for (int c0 = 0; c0 < k; c0 += cap) {
            const int cnt = min(cap, k - c0);
            __syncthreads();                                   // the previous chunk has been consumed by every warp
            for (int t = threadIdx.x; t < cnt; t += blockDim.x) { sm[t] = d[c0 + t]; sm[cap + t] = w[c0 + t]; }
            __syncthreads();
            if (!math) continue;
            const int npsi = max(0, min(cnt, split + 1 - c0));  // poles of this chunk that belong to psi
            int j = lane;
#pragma unroll 4
            for (; j < npsi; j += 32) {
                const double t = (sm[j] - dorg) - tau;
                const double inv = CUPPEN_RCP(t);
                const double r = sm[cap + j] * inv;
                psi += r; dpsi += r * inv;
            }
#pragma unroll 4
            for (; j < cnt; j += 32) {
                const double t = (sm[j] - dorg) - tau;
                const double inv = CUPPEN_RCP(t);
                const double r = sm[cap + j] * inv;
                phi += r; dphi += r * inv;
            }
        }
        SecularSums s;
        WarpLanes L;
        s.psi = L.sum(psi); s.dpsi = L.sum(dpsi); s.phi = L.sum(phi); s.dphi = L.sum(dphi); s.err = fabs(s.psi) + fabs(s.phi);
        return s;
    }
    // an active warp announces itself (the drained ones are waiting in the same vote), then evaluates
    __device__ __forceinline__ SecularSums operator()(double dorg, double tau, int split) const {
        __syncthreads_or(1);
        return run(dorg, tau, split, true);
    }
    __device__ __forceinline__ void drain() const {
        while (__syncthreads_or(0)) run(0.0, 0.0, 0, false);
    }
};

// several GPUs, peer-memory back end: the root is stored into every peer's (origin, tau) arrays as well (H.G > 0)
__device__ __forceinline__ void secular_store(const LevelCtx& c, const SymHeap& H, long g, const SecularRoot& r, double dorg) {
    c.org[g] = r.origin; c.tau[g] = r.tau; c.dorgv[g] = dorg;
    for (int p = 0; p < H.G; ++p) {
        if (p == H.me) continue;
        H.at(p, c.org)[g] = r.origin; H.at(p, c.tau)[g] = r.tau; H.at(p, c.dorgv)[g] = dorg;
    }
}

__global__ void __launch_bounds__(SEC_WARPS * 32) secular_kernel(LevelCtx c, int kcap, int part, int nparts, SymHeap H) {
    extern __shared__ double sec_smem[];
    const int id = blockIdx.y;
    const MergeDesc& D = c.desc[id];
    const int k = D.k;
    // roots of this merge handled by this rank: contiguous range [i0,i1) (multi-GPU root split)
    const int per = (k + nparts - 1) / nparts;
    const int i0 = part * per, i1 = min(k, i0 + per);
    const int i = i0 + blockIdx.x * SEC_WARPS + (threadIdx.x >> 5);
    if (i0 + (int)blockIdx.x * SEC_WARPS >= i1) return;
    const double* dl = c.dl + D.off;
    const double* wl = c.wl + D.off;
    WarpLanes L;
    if (k <= kcap) {
        for (int t = threadIdx.x; t < k; t += blockDim.x) { sec_smem[t] = dl[t]; sec_smem[kcap + t] = wl[t]; }
        __syncthreads();
        if (i >= i1) return;
        SecularRoot r = secular_solve(L, k, sec_smem, sec_smem + kcap, fabs(D.rho), D.sumw, i);
        if (L.lane() == 0) secular_store(c, H, D.off + i, r, sec_smem[r.origin]);
        return;
    }
    const SecularStagedEval ev{dl, wl, k, kcap, sec_smem};
    if (i < i1) {
        SecularRoot r = secular_solve_ev(ev, k, dl, wl, fabs(D.rho), D.sumw, i);
        if (L.lane() == 0) secular_store(c, H, D.off + i, r, dl[r.origin]);
    }
    ev.drain();
}

// Compact as a scan: the canonical index of a live element is the number of live elements before it in the
// sorted z-live list -- a prefix sum, not the O(m^2) count of the per-warp Compact functor (which stays in use
// for the small merges of fused_front_kernel and in the host test build).  One thread-block CLUSTER per merge (1 CTA
// below 4096 entries, up to 8 at 16384 and more): every CTA takes a contiguous range of the sorted list -- its share of the
// Givens sweep, then a reduction pass for the totals (k is needed up front: rho < 0 problems are stored reflected;
// the CTAs exchange their partial totals through distributed shared memory), then a tiled exclusive scan of the three
// flags (live, top-supported, bottom-supported) that starts from the totals of the CTAs before it.  With one CTA per
// merge every pass walked m / 1024 elements per thread with two dependent L2 round trips each: 154 us at m = 16384,
// 92 us at 8192 (profiles/r02_ncu_launches_goe_n16384.csv) -- on every rank, the stage is replicated.
enum { CS_THREADS = 1024, CS_MAX_CLUSTER = 8 };
inline int compact_cluster_size(int maxm) { return maxm >= 8192 ? 8 : maxm >= 4096 ? 4 : maxm >= 2048 ? 2 : 1; }
// cluster barrier, or the block barrier of the one-CTA instantiation (launched without the cluster attribute: the three
// cluster barriers and the cluster launch cost a few microseconds per level on the small merges)
template <bool CLUSTER>
__device__ __forceinline__ void cs_sync() {
    if (CLUSTER) cooperative_groups::this_cluster().sync();
    else __syncthreads();
}
template <bool CLUSTER>
__global__ void __launch_bounds__(CS_THREADS) compact_scan_kernel(LevelCtx c) {
    namespace cg = cooperative_groups;
    const int csize = CLUSTER ? (int)cg::this_cluster().num_blocks() : 1, crank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    MergeDesc& D = c.desc[blockIdx.x / csize];
    // the Givens sweep of this merge first (segment heads walk their segments, also into the ranges of the other CTAs;
    // was a launch of its own)
    {
        const int per = (D.m + csize - 1) / csize;
        const int p1 = min(D.m, (crank + 1) * per);
        for (int p = crank * per + threadIdx.x; p < p1; p += CS_THREADS) GivensSweep{c}(D.off + p);
    }
    if (CLUSTER) __threadfence();
    cs_sync<CLUSTER>();
    const int off = D.off, nl = D.nlive1, n1 = D.n1;
    const int* ls = c.lsort + off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int perq = ((nl + csize - 1) / csize + CS_THREADS - 1) / CS_THREADS * CS_THREADS;     // whole tiles per CTA
    const int q0 = min(nl, crank * perq), q1 = min(nl, q0 + perq);
    __shared__ int s_cnt[3][32];
    __shared__ double s_sw[32];
    __shared__ int s_part[3];          // this CTA's totals (read by the other CTAs of the cluster)
    __shared__ double s_partw;
    __shared__ int s_base[3], s_tot[3];
    // ---- pass 1: totals of the own range ----------------------------------------------------------------
    int k = 0, kt = 0, kb = 0;
    double sw = 0;
    for (int q = q0 + tid; q < q1; q += CS_THREADS) {
        const int eq = ls[q];
        if (c.G[off + eq] != -1) continue;
        const int sp = c.sup[off + eq];
        const double zq = c.zn[off + eq];
        sw += zq * zq;
        k++; kt += (sp & SUP_TOP) ? 1 : 0; kb += (sp & SUP_BOT) ? 1 : 0;
    }
    k = __reduce_add_sync(0xffffffffu, k); kt = __reduce_add_sync(0xffffffffu, kt); kb = __reduce_add_sync(0xffffffffu, kb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
    if (lane == 0) { s_cnt[0][warp] = k; s_cnt[1][warp] = kt; s_cnt[2][warp] = kb; s_sw[warp] = sw; }
    __syncthreads();
    if (warp == 0) {
        int a = s_cnt[0][lane], b = s_cnt[1][lane], d3 = s_cnt[2][lane];
        double w = s_sw[lane];
        a = __reduce_add_sync(0xffffffffu, a); b = __reduce_add_sync(0xffffffffu, b); d3 = __reduce_add_sync(0xffffffffu, d3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) { s_part[0] = a; s_part[1] = b; s_part[2] = d3; s_partw = w; }
    }
    cs_sync<CLUSTER>();                                    // every CTA's partial totals are in its shared memory
    if (tid == 0) {
        int base[3] = {0, 0, 0}, tot[3] = {0, 0, 0};
        double w = 0;
        for (int r = 0; r < csize; ++r) {                  // fixed order: the sum of z^2 does not depend on timing
            const int* rp = CLUSTER ? cg::this_cluster().map_shared_rank(s_part, r) : s_part;
            const double* rw = CLUSTER ? cg::this_cluster().map_shared_rank(&s_partw, r) : &s_partw;
#pragma unroll
            for (int v = 0; v < 3; ++v) { const int x = rp[v]; if (r < crank) base[v] += x; tot[v] += x; }
            w += *rw;
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) { s_base[v] = base[v]; s_tot[v] = tot[v]; }
        if (crank == 0) { D.k = tot[0]; D.ktop = tot[1]; D.kbot = tot[2]; D.sumw = w; }
    }
    cs_sync<CLUSTER>();                                    // remote reads done (no CTA may exit before), s_base / s_tot visible
    const int ktot = s_tot[0];
    const bool neg = D.rho < 0;
    // ---- pass 2: tiled exclusive scan of the own range --------------------------------------------------
    int carry0 = s_base[0], carry1 = s_base[1], carry2 = s_base[2];
    for (int base = q0; base < q1; base += CS_THREADS) {
        const int q = base + tid;
        int e = -1, sp = 0, f0 = 0, f1 = 0, f2 = 0;
        if (q < q1) {
            e = ls[q];
            if (c.G[off + e] == -1) { sp = c.sup[off + e]; f0 = 1; f1 = (sp & SUP_TOP) ? 1 : 0; f2 = (sp & SUP_BOT) ? 1 : 0; }
        }
        // warp-level inclusive scans
        int i0 = f0, i1 = f1, i2 = f2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o), t2 = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) { i0 += t0; i1 += t1; i2 += t2; }
        }
        __syncthreads();                                   // previous tile's s_cnt reads are done
        if (lane == 31) { s_cnt[0][warp] = i0; s_cnt[1][warp] = i1; s_cnt[2][warp] = i2; }
        __syncthreads();
        int w0 = 0, w1 = 0, w2 = 0, t0 = 0, t1 = 0, t2 = 0;       // offsets of this warp, totals of the tile
#pragma unroll 8
        for (int w = 0; w < CS_THREADS / 32; ++w) {
            const int a0 = s_cnt[0][w], a1 = s_cnt[1][w], a2 = s_cnt[2][w];
            if (w < warp) { w0 += a0; w1 += a1; w2 += a2; }
            t0 += a0; t1 += a1; t2 += a2;
        }
        if (f0) {
            const int cnt = carry0 + w0 + i0 - 1, tcnt = carry1 + w1 + i1 - f1, bcnt = carry2 + w2 + i2 - f2;
            const int ci = neg ? (ktot - 1 - cnt) : cnt;
            const double dv = c.dn[off + e], zv = c.zn[off + e];
            c.dl[off + ci] = neg ? -dv : dv;
            c.zl[off + ci] = zv;
            c.wl[off + ci] = zv * zv;
            c.lidx[off + ci] = e;
            if (f1) { c.tpos[off + e] = tcnt; c.toplist[off + tcnt] = ci; }
            if (f2) { c.bpos[off + e] = bcnt; c.botlist[off + n1 + bcnt] = ci; }
        }
        carry0 += t0; carry1 += t1; carry2 += t2;
    }
}
// one cluster of compact_cluster_size(maxm) CTAs per merge
inline void launch_compact_scan(Stream st, int merges, int maxm, LevelCtx c) {
    const int cs = compact_cluster_size(maxm);
    if (cs == 1) {
        compact_scan_kernel<false><<<(unsigned)merges, CS_THREADS, 0, st>>>(c);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(merges * cs), 1, 1);
    cfg.blockDim = dim3(CS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, compact_scan_kernel<true>, c));
}

// ---- the O(k^2) stages as tiled kernels -----------------------------------------------------------------
// Loewner, Norms and RowGemv (merge_stages.h) are Cauchy-like sums/products: every output (a pole or a root)
// visits every item of the other kind with one fp64 division.  The per-warp functors stream the k-vectors
// through L1/L2 once per output; here a CTA owns TL_TJ outputs, stages the other side in shared memory TL_RC
// items at a time (L2 traffic / TL_TJ, broadcast LDS in the inner loop, FP64-pipe bound) and splits every
// chunk over TL_SL slices of threads (thread = (output, slice)) so that small problems still fill the SMs and no
// thread walks more than k / TL_SL items; the slices are combined through shared memory in a fixed order.
// grid.x = tile of outputs, grid.y = merge.  The functors stay in use inside fused_front_kernel (m <= 128) and
// in the host test build.
enum { TL_TJ = 32, TL_SL = 16, TL_THREADS = TL_TJ * TL_SL, TL_RC = 512, TL_SUB = TL_RC / TL_SL, TILED_MIN_M = 2048 };

// zhat_j = sign(z_j) sqrt( |prod_i (lambda_i - d_j) / prod_{i != j} (d_i - d_j)| / |rho| )
// (part, nparts, H): several GPUs with peer memory -- this rank computes the outputs [part*per, (part+1)*per) of every
// merge, per = ceil(k / nparts) rounded up to the tile, and stores them into every rank's copy of the vector
CUPPEN_D int tl_share(int k, int part, int nparts, int* hi) {
    const int per = ((k + nparts - 1) / nparts + TL_TJ - 1) / TL_TJ * TL_TJ;
    const int lo = part * per;
    *hi = (lo + per < k) ? lo + per : k;
    return lo;
}
__global__ void __launch_bounds__(TL_THREADS) loewner_tiled_kernel(LevelCtx c, int part, int nparts, SymHeap H) {
    __shared__ double s_lam_org[TL_RC], s_tau[TL_RC], s_dl[TL_RC];
    __shared__ double s_part[TL_SL][TL_TJ];
    const MergeDesc& D = c.desc[blockIdx.y];
    const int k = D.k, off = D.off;
    int jhi;
    const int j0 = tl_share(k, part, nparts, &jhi) + blockIdx.x * TL_TJ;
    if (j0 >= jhi) return;
    const int out = threadIdx.x & (TL_TJ - 1), slice = threadIdx.x / TL_TJ;
    const int j = j0 + out;
    const double* dl = c.dl + off;
    const double dj = dl[j < k ? j : k - 1];
    double prod = 1.0;
    for (int i0 = 0; i0 < k; i0 += TL_RC) {
        const int cnt = min((int)TL_RC, k - i0);
        for (int t = threadIdx.x; t < cnt; t += TL_THREADS) {
            s_lam_org[t] = dl[c.org[off + i0 + t]];
            s_tau[t] = c.tau[off + i0 + t];
            s_dl[t] = dl[i0 + t];
        }
        __syncthreads();
        const int t1 = min(cnt, (slice + 1) * TL_SUB);
#pragma unroll 4
        for (int t = slice * TL_SUB; t < t1; ++t) {
            const double num = (s_lam_org[t] - dj) + s_tau[t];
            const double den = s_dl[t] - dj;
            prod *= (i0 + t == j) ? num : num * CUPPEN_RCP(den);
        }
        __syncthreads();
    }
    s_part[slice][out] = prod;
    __syncthreads();
    if (slice == 0 && j < k) {
        double p = s_part[0][out];
#pragma unroll
        for (int q = 1; q < TL_SL; ++q) p *= s_part[q][out];
        const double zh = sqrt(fabs(p) / fabs(D.rho));
        const double v = (c.zl[off + j] < 0) ? -zh : zh;
        c.zhat[off + j] = v;
        for (int r = 0; r < H.G; ++r) if (r != H.me) H.at(r, c.zhat)[off + j] = v;
    }
}

// N_i = || zhat / (d - lambda_i) ||_2
__global__ void __launch_bounds__(TL_THREADS) norms_tiled_kernel(LevelCtx c, int part, int nparts, SymHeap H) {
    __shared__ double s_dl[TL_RC], s_zh[TL_RC];
    __shared__ double s_part[TL_SL][TL_TJ];
    const MergeDesc& D = c.desc[blockIdx.y];
    const int k = D.k, off = D.off;
    int ihi;
    const int i0 = tl_share(k, part, nparts, &ihi) + blockIdx.x * TL_TJ;
    if (i0 >= ihi) return;
    const int out = threadIdx.x & (TL_TJ - 1), slice = threadIdx.x / TL_TJ;
    const int i = i0 + out;
    const double* dl = c.dl + off;
    const int ii = i < k ? i : k - 1;
    const double dorg = dl[c.org[off + ii]], t = c.tau[off + ii];
    double s = 0.0;
    for (int q0 = 0; q0 < k; q0 += TL_RC) {
        const int cnt = min((int)TL_RC, k - q0);
        for (int q = threadIdx.x; q < cnt; q += TL_THREADS) { s_dl[q] = dl[q0 + q]; s_zh[q] = c.zhat[off + q0 + q]; }
        __syncthreads();
        const int q1 = min(cnt, (slice + 1) * TL_SUB);
#pragma unroll 4
        for (int q = slice * TL_SUB; q < q1; ++q) {
            const double u = s_zh[q] * CUPPEN_RCP((s_dl[q] - dorg) - t);
            s = fma(u, u, s);
        }
        __syncthreads();
    }
    s_part[slice][out] = s;
    __syncthreads();
    if (slice == 0 && i < k) {
        double a = s_part[0][out];
#pragma unroll
        for (int q = 1; q < TL_SL; ++q) a += s_part[q][out];
        const double v = sqrt(a);
        c.nrm[off + i] = v;
        for (int r = 0; r < H.G; ++r) if (r != H.me) H.at(r, c.nrm)[off + i] = v;
    }
}

// boundary rows of the merged node (eigenvalue-only mode, src/main.c:613-639): first row over the top-supported
// live columns, last row over the bottom-supported ones
__global__ void __launch_bounds__(TL_THREADS) rowgemv_tiled_kernel(LevelCtx c, RowCtx r) {
    __shared__ double s_dl[TL_RC], s_w[TL_RC];
    __shared__ double s_part[2][TL_SL][TL_TJ];
    const MergeDesc& D = c.desc[blockIdx.y];
    const int k = D.k, off = D.off;
    const int i0 = blockIdx.x * TL_TJ;
    if (i0 >= k) return;
    const int out = threadIdx.x & (TL_TJ - 1), slice = threadIdx.x / TL_TJ;
    const int i = i0 + out;
    const double* dl = c.dl + off;
    const double* zh = c.zhat + off;
    const int ii = i < k ? i : k - 1;
    const double dorg = dl[c.org[off + ii]], t = c.tau[off + ii];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int kh = half ? D.kbot : D.ktop;
        const int* list = half ? c.botlist + off + D.n1 : c.toplist + off;
        const double* pk = half ? r.lpack + off + D.n1 : r.fpack + off;
        double s = 0.0;
        for (int q0 = 0; q0 < kh; q0 += TL_RC) {
            const int cnt = min((int)TL_RC, kh - q0);
            for (int q = threadIdx.x; q < cnt; q += TL_THREADS) {
                const int jq = list[q0 + q];
                s_dl[q] = dl[jq];
                s_w[q] = pk[q0 + q] * zh[jq];
            }
            __syncthreads();
            const int q1 = min(cnt, (slice + 1) * TL_SUB);
#pragma unroll 4
            for (int q = slice * TL_SUB; q < q1; ++q) s = fma(s_w[q], CUPPEN_RCP((s_dl[q] - dorg) - t), s);
            __syncthreads();
        }
        s_part[half][slice][out] = s;
    }
    __syncthreads();
    if (slice == 0 && i < k) {
        double a = s_part[0][0][out], b = s_part[1][0][out];
#pragma unroll
        for (int q = 1; q < TL_SL; ++q) { a += s_part[0][q][out]; b += s_part[1][q][out]; }
        const double rn = 1.0 / c.nrm[off + i];
        r.frow_new[off + c.lidx[off + i]] = a * rn;
        r.lrow_new[off + c.lidx[off + i]] = b * rn;
    }
}

// MergeTol (merge_stages.h) with one CTA per merge instead of one warp: the warp walked m / 32 dependent pairs of loads
// (20 us at m = 2048 for a stage that every accurate-rule level waits on)
__global__ void __launch_bounds__(256) merge_tol_kernel(LevelCtx c) {
    MergeDesc& D = c.desc[blockIdx.x];
    if (D.mode == MODE_REFERENCE) return;
    __shared__ double s_d[8], s_z[8];
    double dmax = 0, zmax = 0;
    const double* dd = c.d + D.off;
    const double* zz = c.z + D.off;
    for (int t = threadIdx.x; t < D.m; t += 256) { dmax = fmax(dmax, fabs(dd[t])); zmax = fmax(zmax, fabs(zz[t])); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o)); zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o)); }
    if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = dmax; s_z[threadIdx.x >> 5] = zmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) { dmax = fmax(dmax, s_d[w]); zmax = fmax(zmax, s_z[w]); }
        D.tol = 8.0 * 2.220446049250313e-16 * fmax(dmax, D.sigma * zmax);
    }
}

// RankLive (stable enumeration sort, merge_stages.h) in the same tiled shape: a thread owns one
// element and one slice of every staged chunk of keys; z-deflated entries are staged as NaN (never "before").
// One DSETP per pair: the tie-break by index is decided per CHUNK -- keys staged from indices below the CTA's outputs
// count with `<=`, keys from above with `<`, only the chunk that contains the outputs pays the full (d, index)
// comparison -- and the live total is counted while staging, not per pair (round 1 issued three DSETPs per pair and
// was bound by them, VERDICT weak #4).
__global__ void __launch_bounds__(TL_THREADS) rank_tiled_kernel(LevelCtx c) {
    __shared__ double s_d[TL_RC];
    __shared__ int s_cnt[TL_SL][TL_TJ], s_tot[TL_THREADS / 32];
    MergeDesc& D = c.desc[blockIdx.y];
    const int m = D.m, off = D.off;
    const int j0 = blockIdx.x * TL_TJ;
    if (j0 >= m) return;
    const int out = threadIdx.x & (TL_TJ - 1), slice = threadIdx.x / TL_TJ;
    const int j = j0 + out;
    const double dj = (j < m) ? c.d[off + j] : 0.0;
    // a CTA whose outputs are all z-deflated has nothing to rank (heavily deflating matrices: most CTAs of the upper
    // levels); block 0 always runs, it counts the live entries of the merge
    const bool live_out = (j < m) && c.G[off + j] != -2;
    if (!__syncthreads_or(live_out ? 1 : 0) && blockIdx.x != 0) return;
    int cnt = 0, tot = 0;
    for (int t0 = 0; t0 < m; t0 += TL_RC) {
        const int n_here = min((int)TL_RC, m - t0);
        for (int t = threadIdx.x; t < n_here; t += TL_THREADS) {
            const bool live = c.G[off + t0 + t] != -2;
            s_d[t] = live ? c.d[off + t0 + t] : __longlong_as_double(0x7ff8000000000000LL);
            tot += live ? 1 : 0;
        }
        __syncthreads();
        const int t1 = min(n_here, (slice + 1) * TL_SUB);
        if (t0 + n_here <= j0) {                     // every staged index is below every output of this CTA
#pragma unroll 4
            for (int t = slice * TL_SUB; t < t1; ++t) cnt += (s_d[t] <= dj) ? 1 : 0;
        } else if (t0 >= j0 + TL_TJ) {               // every staged index is above
#pragma unroll 4
            for (int t = slice * TL_SUB; t < t1; ++t) cnt += (s_d[t] < dj) ? 1 : 0;
        } else {
#pragma unroll 4
            for (int t = slice * TL_SUB; t < t1; ++t) {
                const double dt = s_d[t];
                cnt += ((dt < dj) || (dt == dj && t0 + t < j)) ? 1 : 0;
            }
        }
        __syncthreads();
    }
    s_cnt[slice][out] = cnt;
    if (blockIdx.x == 0) {                           // live total of the merge: block 0 staged every key exactly once
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = tot;
    }
    __syncthreads();
    if (slice == 0 && j < m) {
#pragma unroll
        for (int q = 1; q < TL_SL; ++q) cnt += s_cnt[q][out];
        if (live_out) c.lsort[off + cnt] = j;
        if (j == 0) {
            int total = 0;
#pragma unroll
            for (int w = 0; w < TL_THREADS / 32; ++w) total += s_tot[w];
            D.nlive1 = total;
        }
    }
}

// Fused front end for the small merges at the bottom of the tree (m <= FUSE_MAXM): ONE CTA per merge runs every vector
// stage (z assembly ... new eigenvalues, and in eigenvalue-only mode the boundary-row update) back to back, with all 23
// per-level vectors of the merge held in SHARED memory: the stage functors of merge_stages.h run unchanged on a LevelCtx
// whose vector pointers are redirected to shared memory (biased by -off, so that the global index g = off + j still
// works), __syncthreads between the stages, and the vectors are copied out once at the end for the matrix kernels
// (pack / U / GEMM) and the selected-eigenvector mode.  Round 1 ran these stages on global memory -- a cluster of 4 CTAs
// per merge with cluster barriers, ~12 dependent round trips to L2 per level: 32 / 50 / 108 us for the levels m = 32, 64,
// 128 of n = 4096; the shared-memory version takes 26 / 38 / 71 us (profiles/r02_ncu_launches_s1_n4096_fused_smem_512.csv).
// One CTA per merge stops paying at m = 256 (124 us against ~80 us for the separate kernels, whose O(m^2) stages spread
// over many CTAs) and m = 512 (186 us), so the fused path ends at m = 128.
enum { FUSE_MAXM = 128, FUSE_THREADS = 1024, FUSE_NVEC_D = 12, FUSE_NVEC_I = 11 };
inline size_t fused_front_smem_bytes(int mcap) { return (size_t)mcap * (FUSE_NVEC_D * sizeof(double) + FUSE_NVEC_I * sizeof(int)); }

__global__ void __launch_bounds__(FUSE_THREADS) fused_front_kernel(LevelCtx c, RowCtx rc, int rows_mode, int mcap) {
    extern __shared__ __align__(16) unsigned char fuse_smem[];
    const int id = blockIdx.x;
    const int off = c.desc[id].off, m = c.desc[id].m;
    const int tid = threadIdx.x, warp = tid >> 5, nthreads = blockDim.x, nwarps = blockDim.x / 32;
    double* sd = reinterpret_cast<double*>(fuse_smem);
    int* si = reinterpret_cast<int*>(fuse_smem + (size_t)mcap * FUSE_NVEC_D * sizeof(double));
    // global homes of the vectors and their shared-memory stand-ins, in the same order
    double* gd[FUSE_NVEC_D] = {c.d, c.z, c.dn, c.zn, c.gc, c.gs, c.dl, c.wl, c.zl, c.tau, c.zhat, c.nrm};
    int* gi[FUSE_NVEC_I] = {c.G, c.lsort, c.head, c.sup, c.prev, c.tpos, c.bpos, c.lidx, c.org, c.toplist, c.botlist};
    LevelCtx s = c;
    {
        double** pd[FUSE_NVEC_D] = {&s.d, &s.z, &s.dn, &s.zn, &s.gc, &s.gs, &s.dl, &s.wl, &s.zl, &s.tau, &s.zhat, &s.nrm};
        int** pi[FUSE_NVEC_I] = {&s.G, &s.lsort, &s.head, &s.sup, &s.prev, &s.tpos, &s.bpos, &s.lidx, &s.org, &s.toplist, &s.botlist};
#pragma unroll
        for (int v = 0; v < FUSE_NVEC_D; ++v) *pd[v] = sd + (size_t)v * mcap - off;
#pragma unroll
        for (int v = 0; v < FUSE_NVEC_I; ++v) *pi[v] = si + (size_t)v * mcap - off;
    }
    // (vectors that a stage reads before any stage of this level wrote them -- lsort / lidx / toplist / botlist beyond the
    // live counts, gc / gs of unrotated entries -- keep whatever the copy-out of an earlier level left in global memory:
    // start from the global contents so that the copy-out does not publish uninitialised shared memory)
    for (int j = tid; j < m; j += nthreads) {
#pragma unroll
        for (int v = 0; v < FUSE_NVEC_D; ++v) sd[(size_t)v * mcap + j] = gd[v][off + j];
#pragma unroll
        for (int v = 0; v < FUSE_NVEC_I; ++v) si[(size_t)v * mcap + j] = gi[v][off + j];
    }
    __syncthreads();
    const WarpLanes L;
    for (int g = off + tid; g < off + m; g += nthreads) ZAssemble{s}(g);
    __syncthreads();
    if (warp == 0) MergeTol{s}(id, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += nthreads) FlagDeflate{s}(g);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) RankLive{s}(g, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += nthreads) GivensSweep{s}(g);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) Compact{s}(g, L);
    __syncthreads();
    {
        const MergeDesc& D = c.desc[id];
        const int k = D.k;
        for (int i = warp; i < k; i += nwarps) {
            SecularRoot r = secular_solve(L, k, s.dl + off, s.wl + off, fabs(D.rho), D.sumw, i);
            if (L.lane() == 0) { s.org[off + i] = r.origin; s.tau[off + i] = r.tau; c.dorgv[off + i] = s.dl[off + r.origin]; }
        }
    }
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) Loewner{s}(g, L);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) Norms{s}(g, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += nthreads) NewLambda{s}(g);
    if (rows_mode) {
        __syncthreads();
        for (int g = off + tid; g < off + m; g += nthreads) RowPack{s, rc}(g);
        __syncthreads();
        for (int g = off + warp; g < off + m; g += nwarps) RowGemv{s, rc}(g, L);
        __syncthreads();
        for (int g = off + tid; g < off + m; g += nthreads) RowCommit{s, rc, const_cast<double*>(c.frow), const_cast<double*>(c.lrow)}(g);
    }
    __syncthreads();
    for (int j = tid; j < m; j += nthreads) {
#pragma unroll
        for (int v = 0; v < FUSE_NVEC_D; ++v) gd[v][off + j] = sd[(size_t)v * mcap + j];
#pragma unroll
        for (int v = 0; v < FUSE_NVEC_I; ++v) gi[v][off + j] = si[(size_t)v * mcap + j];
    }
}

// few merges: all 1024 threads on each (latency); more merges than SMs: smaller CTAs so that several are resident per SM
inline void launch_fused_front(Stream st, int num_sms, int merges, int maxm, LevelCtx c, RowCtx rc, int rows_mode) {
    const int mcap = (maxm + 31) / 32 * 32;
    const int threads = merges <= num_sms ? FUSE_THREADS : (merges <= 2 * num_sms ? 512 : 256);
    fused_front_kernel<<<(unsigned)merges, threads, fused_front_smem_bytes(mcap), st>>>(c, rc, rows_mode, mcap);
    CUDA_CHECK(cudaGetLastError());
}

// K5a: walk every rotation chain once per row.  grid.x = global column, grid.y = row chunk.
// Columns without a rotation (z-deflated, or live and unrotated: the common cases) take a fast path
// that issues the loads of all PACK_ROWS rows before the stores.
enum { PACK_THREADS = 128, PACK_ROWS = 8 };
__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(LevelCtx c, MatCtx M) {
    const int g = blockIdx.x;
    // The blocks are short and bound by their dependent L2 round trips: node_of[g] -> descriptor -> G[g] / head[g] were
    // three of them (the compiler sinks loads below the early exits).  G[g] and head[g] depend on g only, and the
    // descriptor fields sit in three different 32-byte sectors: the two exit tests below are written so that they
    // consume everything that can be requested at that point -- two round trips.  (G >= -2, head in {0, 1} and the
    // descriptor fields are >= 0, so the extra terms never fire.  L1 prefetches -- CCTL.PF1 -- instead: 4x slower.)
    const int Gg = c.G[g];
    const int headg = c.head[g];
    const int id = c.node_of[g];
    if ((id | (Gg + 2) | headg) < 0) return;
    const MergeDesc& D = c.desc[id];
    const int off = D.off, Dm = D.m, Dn1 = D.n1, Dkt = D.ktop, Dkb = D.kbot, Dlr1 = D.lr1, Dlsplit = D.lsplit;
    const int rbase = D.lr0 + blockIdx.y * PACK_ROWS * PACK_THREADS + threadIdx.x;      // local row
    if (((Dlr1 - 1 - rbase) | off | Dm | Dn1 | Dkt | Dkb | Dlsplit) < 0) return;      // rbase >= lr1: no rows of the node here
    const int e = g - off;
    const bool zdefl = (Gg == -2);
    const bool etop = e < Dn1;
    // (the zdefl / head early return was moved below the tail zeroing)
    {
        // zero the K tail of Apack: columns [kh, round_up(kh, K_PAD)) of each half -- the GEMM reads K in multiples
        // of K_PAD.  Column off+e is handled by this block; a tail that runs past the node's last column (m not a
        // multiple of K_PAD) is finished by the block of the last column.  (Was a separate launch, pack_tail_kernel.)
        const int kt = Dkt, kb = Dkb;
        const int kt_end = (kt + K_PAD - 1) / K_PAD * K_PAD, kb_end = (kb + K_PAD - 1) / K_PAD * K_PAD;
        const int last = (e == Dm - 1) ? max(kt_end, kb_end) : e + 1;
        for (int kk = e; kk < last; ++kk) {
            if (!((kk >= kt && kk < kt_end) || (kk >= kb && kk < kb_end))) continue;
#pragma unroll
            for (int t = 0; t < PACK_ROWS; ++t) {
                const int r = rbase + t * PACK_THREADS;
                if (r >= Dlr1) break;
                const bool rtop = r < Dlsplit;
                const int kh = rtop ? kt : kb, kend = rtop ? kt_end : kb_end;
                if (kk >= kh && kk < kend) M.Apack[(long)r + (long)(off + kk) * M.ldq] = 0.0;
            }
        }
    }
    // row support (RowSpan): every column that this merge rewrites -- a root column of the GEMM or a rotated column --
    // spans the parent block from now on.  (Placed here, after the descriptor fields have been consumed: a store at the
    // top of the kernel made the warp wait for G[g] before it issued the other descriptor loads -- one more dependent L2
    // round trip per block, 1.08 -> 1.41 ms of pack time at n = 16384, gpurun_out r02 call K.)
    if (M.span != nullptr && !zdefl && blockIdx.y == 0 && threadIdx.x == 0) M.span[g] = RowSpan{off, off + Dm};
    if (!zdefl && !headg) return;
    if (zdefl && M.Qnew == M.Qold) {
        // in place: a z-deflated column keeps its own-half rows where they are; only the other half's rows of the
        // parent block are new and must read zero (they may hold the previous solve's V)
        double* dst = M.Qnew + (long)g * M.ldq;
#pragma unroll
        for (int t = 0; t < PACK_ROWS; ++t) {
            const int r = rbase + t * PACK_THREADS;
            if (r < Dlr1 && ((r < Dlsplit) != etop)) dst[r] = 0.0;
        }
        return;
    }
    if (zdefl || Gg == -1) {
        // plain column move: own-half rows from the child, zeros in the other half (deflated columns
        // go to Q', live ones to their K slot of Apack -- only over the rows of their own half)
        double* dst;
        bool write_other;
        if (zdefl) { dst = M.Qnew + (long)g * M.ldq; write_other = true; }
        else {
            const int pos = etop ? c.tpos[g] : c.bpos[g];
            dst = M.Apack + (long)(off + pos) * M.ldq;
            write_other = false;
        }
        const double* src = M.Qold + (long)g * M.ldq;
        double v[PACK_ROWS];
#pragma unroll
        for (int t = 0; t < PACK_ROWS; ++t) {
            const int r = rbase + t * PACK_THREADS;
            v[t] = (r < Dlr1 && ((r < Dlsplit) == etop)) ? src[r] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < PACK_ROWS; ++t) {
            const int r = rbase + t * PACK_THREADS;
            if (r < Dlr1 && (write_other || ((r < Dlsplit) == etop))) dst[r] = v[t];
        }
        return;
    }
    for (int t = 0; t < PACK_ROWS; ++t) {
        const int r = rbase + t * PACK_THREADS;
        if (r >= Dlr1) return;
        const long rl = r;
        const bool rtop = r < Dlsplit;
        int a = e;
        double carry = ((a < Dn1) == rtop) ? M.Qold[rl + (long)(off + a) * M.ldq] : 0.0;
        int b;
        while ((b = c.G[off + a]) >= 0) {
            const double cs = c.gc[off + a], sn = c.gs[off + a];
            const double x = ((b < Dn1) == rtop) ? M.Qold[rl + (long)(off + b) * M.ldq] : 0.0;
            M.Qnew[rl + (long)(off + a) * M.ldq] = cs * carry - sn * x;
            carry = sn * carry + cs * x;
            a = b;
        }
        const int pos = rtop ? c.tpos[off + a] : c.bpos[off + a];
        if (pos >= 0) M.Apack[rl + (long)(off + pos) * M.ldq] = carry;
    }
}

// K5b: B[row = arena row of pole j][col = root i - p0] = zhat_j / (((d_j - d_org(i)) - tau_i) N_i)
// A block owns UG_ROWS consecutive arena rows and strides over 256-wide column chunks: a thread keeps the three
// per-root constants (origin pole, tau, norm) of its column in registers and walks down the rows, whose per-pole
// constants sit in shared memory -- one division and 8 stored bytes per element, UG_ROWS independent divisions in
// flight per thread, rows written as coalesced 2 KB segments.  (The first version spent four dependent L2 loads per
// element -- dl[org[i]], tau[i], nrm[i] -- on one row per block: 2 TB/s; profiles/README.md.)
enum { UG_ROWS = 8, UG_COLS = 4 };
__global__ void __launch_bounds__(256) ugen_kernel(LevelCtx c, MatCtx M, int p0, int width, int new_lambda, WorkCtx w) {
    extern __shared__ int ugen_dyn_smem[];
    __shared__ double s_dj[UG_ROWS], s_zj[UG_ROWS];
    __shared__ int s_id[UG_ROWS];
    if (blockIdx.x == gridDim.x - 1) {               // the extra block: GEMM work list of this level and panel (was a launch of its own)
        if (blockIdx.y == 0) build_gemm_work_body(w, ugen_dyn_smem);
        return;
    }
    const int row0 = blockIdx.x * UG_ROWS;
    if (threadIdx.x < UG_ROWS) {
        const int row = row0 + threadIdx.x;
        int id = -1;
        if (row < c.n) {
            // the new eigenvalues of the level (NewLambda, one thread per index) ride along with the first panel: nothing
            // between here and the next level's z assembly reads lam
            if (new_lambda && blockIdx.y == 0) NewLambda{c}(row);
            // (the K-list entries depend on the row only: requested together with node_of[row] -- the test below consumes
            // them, or the compiler sinks the loads behind the descriptor -- three dependent round trips instead of four)
            const int jt = c.toplist[row], jb = c.botlist[row];
            id = c.node_of[row];
            if (id >= 0 && (jt | jb) != (int)0x80000000) {       // (list entries are indices or stale non-negative values: always true)
                const MergeDesc& D = c.desc[id];
                const bool top = row < D.off + D.n1;
                const int jj = row - (top ? D.off : D.off + D.n1);
                const int kh = top ? D.ktop : D.kbot;
                if (jj >= kh) id = -1;
                else {
                    const int j = top ? jt : jb;
                    s_dj[threadIdx.x] = c.dl[D.off + j];
                    s_zj[threadIdx.x] = c.zhat[D.off + j];
                }
            }
        }
        s_id[threadIdx.x] = id;
    }
    __syncthreads();
    int cur = -1, off = 0, iend = 0;
    for (int r0 = 0; r0 < UG_ROWS; ++r0) {            // (usually one pass: all rows of a tile belong to one merge)
        const int id = s_id[r0];
        if (id < 0 || id == cur) continue;
        bool seen = false;
        for (int q = 0; q < r0; ++q) seen = seen || (s_id[q] == id);
        if (seen) continue;
        cur = id;
        const MergeDesc& D = c.desc[id];
        off = D.off;
        iend = min(D.k, p0 + width);
        // UG_COLS columns per thread and pass: their (origin pole, tau, norm) -- three independent coalesced loads each -- are
        // all issued before the first division, so a block pays the L2 latency once per pass instead of once per column
        // (with a dependent gather dl[org[i]] on top, the first version was latency-bound at 3.2 TB/s)
        const int stride = gridDim.y * 256;
        for (int ib = p0 + blockIdx.y * 256 + threadIdx.x; ib < iend; ib += UG_COLS * stride) {
            double dorg[UG_COLS], tau[UG_COLS], nrm[UG_COLS];
#pragma unroll
            for (int u = 0; u < UG_COLS; ++u) {
                const int i = ib + u * stride;
                const int ii = i < iend ? i : iend - 1;
                dorg[u] = c.dorgv[off + ii]; tau[u] = c.tau[off + ii]; nrm[u] = c.nrm[off + ii];
            }
#pragma unroll
            for (int u = 0; u < UG_COLS; ++u) {
                const int i = ib + u * stride;
                if (i >= iend) break;
#pragma unroll
                for (int r = 0; r < UG_ROWS; ++r) {
                    if (s_id[r] != id) continue;
                    const double den = ((s_dj[r] - dorg[u]) - tau[u]) * nrm[u];
                    double v = s_zj[r] * CUPPEN_RCP(den);
                    if (!(fabs(v) < 1.7e308)) v = ((s_zj[r] < 0) != (den < 0)) ? -1.7e308 : 1.7e308;   // a root on its pole: clamp, never inf / NaN
                    M.B[(long)(row0 + r) * M.ldb + (i - p0)] = v;
                }
            }
        }
    }
}

// K8: residuals of RES_NC output columns per block over one contiguous slice of rows: global rows
// [g0, g0+cnt) stored at local rows [l0, l0+cnt).  Output column c (ascending lambda) is storage column
// perm[c] of V.  A thread owns two consecutive rows per step.  Interior pairs take a branch-free fast path:
// the diagonal / off-diagonal entries are loaded once (16 bytes each) and reused for all RES_NC columns; per
// column one 16-byte load of (x[r], x[r+1]) plus the two neighbours x[r-1], x[r+2] (L1 hits: the adjacent
// threads' lines) and ten FP64 instructions -- few enough instructions per byte that the kernel is bound by HBM
// and not by the issue rate (the first version spent ~80 instructions per 16 bytes on boundary predicates).
// Pairs that touch the slice boundary (halo rows of the multi-GPU layout, first / last row of T, odd offsets)
// take a general per-row path.  `accumulate` adds to res2 (a rank holds several slices when the rows are
// distributed).  RES_NC and the minimum resident blocks per SM are template parameters: the variants were timed
// on the B200 (cuppen_selftest_residual, profiles/README.md) and launch_residual() picks the default.
enum { RES_DEFAULT_VARIANT = 25, RES_MAX_SLICES = 16 };
// the row slices a rank holds: global rows [g0, g0+cnt) stored at local rows [l0, l0+cnt), with the rows just above /
// below the slice (halo rows, per output column) -- one slice on one GPU, one per subtree in the multi-GPU slice layout.
// All slices are handled by ONE launch (a launch per slice left the 256-row slices of 8 GPUs with 2 % of the work of a
// one-GPU launch each, and the phase took longer on 4 GPUs than on one).
struct ResSlices {
    int ns;
    int g0[RES_MAX_SLICES], l0[RES_MAX_SLICES], cnt[RES_MAX_SLICES];
    const double* lo[RES_MAX_SLICES];
    const double* hi[RES_MAX_SLICES];
};
template <int RES_NC, int MINB>
__global__ void __launch_bounds__(256, MINB) residual_kernel(const double* __restrict__ V, long ldq, int n, ResSlices S,
                                                             const double* __restrict__ OD, const double* __restrict__ OE,
                                                             const double* __restrict__ lam_sorted, const int* __restrict__ perm,
                                                             double* __restrict__ res2, const RowSpan* __restrict__ span) {
    const int col0 = blockIdx.x * RES_NC;
    const int lane = threadIdx.x & 31;
    // per-column constants live in shared memory (broadcast reads) to keep the register budget for loads in flight
    __shared__ const double* xb[RES_NC];
    __shared__ double lambda[RES_NC];
    __shared__ int slo[RES_NC], shi[RES_NC];
    double acc[RES_NC];
    if (threadIdx.x < RES_NC) {
        const int col = min(col0 + (int)threadIdx.x, n - 1);      // columns past the end repeat the last one (not stored)
        xb[threadIdx.x] = V + (long)perm[col] * ldq;              // (ldq is even: the parity of an element's address is that of its row)
        lambda[threadIdx.x] = lam_sorted[col];
        slo[threadIdx.x] = span ? span[perm[col]].lo : 0;
        shi[threadIdx.x] = span ? span[perm[col]].hi : n;
    }
    __syncthreads();
    // rows that can contribute: the union of the block's column supports, widened by one row on either side (the rows
    // next to a support see it through the off-diagonal); everything else is (d - lambda) 0 + e 0 + e 0
    int blo = slo[0], bhi = shi[0];
#pragma unroll
    for (int c = 1; c < RES_NC; ++c) { blo = min(blo, slo[c]); bhi = max(bhi, shi[c]); }
    blo -= 1; bhi += 1;
#pragma unroll
    for (int c = 0; c < RES_NC; ++c) acc[c] = 0.0;
    for (int sl = 0; sl < S.ns; ++sl) {
        const int g0 = S.g0[sl], g1 = g0 + S.cnt[sl];
        const long shift = (long)S.l0[sl] - g0;                   // element of global row r = xb[c][r + shift]
        const bool vec_ok = ((shift & 1) == 0) && ((ldq & 1) == 0);   // 16-byte loads of (x[r], x[r+1]) with r even
        const double* halo_lo = S.lo[sl];
        const double* halo_hi = S.hi[sl];
        const int ra = max(g0, blo), rb = min(g1, bhi);
        for (int r = (ra & ~1) + 2 * (int)threadIdx.x; r < rb; r += 2 * 256) {
            if (vec_ok && r - 1 >= g0 && r + 2 < g1) {
                // interior pair: rows r-1 .. r+2 are all inside the slice (so 0 < r and r + 1 < n - 1)
                const double2 dv = *reinterpret_cast<const double2*>(OD + r);
                const double2 ev = *reinterpret_cast<const double2*>(OE + r);
                const double em = OE[r - 1];
#pragma unroll
                for (int c = 0; c < RES_NC; ++c) {
                    const double* xc = xb[c] + shift;
                    const double2 xv = *reinterpret_cast<const double2*>(xc + r);
                    const double xm = xc[r - 1], xp = xc[r + 2];
                    const double lam = lambda[c];
                    const double y0 = fma(dv.x - lam, xv.x, fma(em, xm, ev.x * xv.y));
                    const double y1 = fma(dv.y - lam, xv.y, fma(ev.x, xv.x, ev.y * xp));
                    acc[c] = fma(y0, y0, fma(y1, y1, acc[c]));
                }
            } else {
#pragma unroll 1
                for (int rr = max(r, g0); rr < min(r + 2, g1); ++rr) {
                    const double dr = OD[rr];
                    const double el = (rr > 0) ? OE[rr - 1] : 0.0, eu = (rr < n - 1) ? OE[rr] : 0.0;
#pragma unroll
                    for (int c = 0; c < RES_NC; ++c) {                  // (unrolled: acc[] must stay in registers)
                        const int col = min(col0 + c, n - 1);
                        const double* xc = xb[c] + shift;
                        const double xr = xc[rr];
                        double y = (dr - lambda[c]) * xr;
                        if (rr > 0) y = fma(el, (rr > g0) ? xc[rr - 1] : halo_lo[col], y);
                        if (rr < n - 1) y = fma(eu, (rr + 1 < g1) ? xc[rr + 1] : halo_hi[col], y);
                        acc[c] = fma(y, y, acc[c]);
                    }
                }
            }
        }
    }
    __shared__ double red[8][RES_NC];
#pragma unroll
    for (int c = 0; c < RES_NC; ++c) {
        double a = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[threadIdx.x >> 5][c] = a;
    }
    __syncthreads();
    if (threadIdx.x < RES_NC && col0 + (int)threadIdx.x < n) {
        double sum = 0;
        for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
        res2[col0 + threadIdx.x] = sum;
    }
}

// variant: 0 default, else NC*10 + MINB
inline void launch_residual(Stream st, int variant, const double* V, long ldq, int n, const ResSlices& S, const double* OD,
                            const double* OE, const double* lam_sorted, const int* perm, double* res2, const RowSpan* span = nullptr) {
#define CUPPEN_RES_CASE(NC, MB)                                                                                          \
    case NC * 10 + MB:                                                                                                   \
        residual_kernel<NC, MB><<<(unsigned)((n + NC - 1) / NC), 256, 0, st>>>(V, ldq, n, S, OD, OE, lam_sorted, perm, res2, span); \
        break;
    switch (variant == 0 ? RES_DEFAULT_VARIANT : variant) {
        CUPPEN_RES_CASE(1, 4) CUPPEN_RES_CASE(1, 6) CUPPEN_RES_CASE(2, 4) CUPPEN_RES_CASE(2, 5) CUPPEN_RES_CASE(4, 3)
        CUPPEN_RES_CASE(4, 4) CUPPEN_RES_CASE(8, 2) CUPPEN_RES_CASE(8, 3) CUPPEN_RES_CASE(4, 2) CUPPEN_RES_CASE(2, 3)
        default: CUPPEN_THROW(-1, "unknown residual kernel variant %d", variant);
    }
#undef CUPPEN_RES_CASE
    CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256) gather_cols_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                          long ldq, int rows, const int* __restrict__ perm) {
    const int cidx = blockIdx.x;
    const double* s = src + (long)perm[cidx] * ldq;
    double* d = dst + (long)cidx * ldq;
    for (int r = blockIdx.y * 256 + threadIdx.x; r < rows; r += gridDim.y * 256) d[r] = s[r];
}
// the listed columns only (ascending-lambda ranks idx[]; perm == nullptr: the storage order is the rank order)
__global__ void __launch_bounds__(256) gather_sel_cols_kernel(const double* __restrict__ src, double* __restrict__ dst, long ldq, int rows,
                                                              const int* __restrict__ perm, const int* __restrict__ idx) {
    const int c = blockIdx.x;
    const int col = perm ? perm[idx[c]] : idx[c];
    const double* s = src + (long)col * ldq;
    double* d = dst + (long)c * ldq;
    for (int r = blockIdx.y * 256 + threadIdx.x; r < rows; r += gridDim.y * 256) d[r] = s[r];
}
#endif  // CUPPEN_CUDA

}  // namespace cuppen
#endif
