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

namespace cuppen {

enum { K_PAD = 32 };            // the GEMM reads K in multiples of this; A is zero-padded up to it
enum { LEAF_MAX = 32 };         // largest leaf handled by one warp

struct LeafDesc { int off, n; };

struct MatCtx {
    int n;                // global size
    long ldq;             // leading dimension of the Q buffers and of Apack (local rows, padded)
    const double* Qold;   // children (block diagonal), column-major, local rows
    double* Qnew;         // parents
    double* Apack;        // packed live columns (K order), same shape as Q plus K_PAD columns
    double* B;            // U arena, row-major [n + pad][ldb]
    long ldb;
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
    int* ntiles;            // [0] tile count, [1] number of problems whose lines are not 16-byte aligned
    int tile_cap;
};

// Tile order of one problem: super-columns of `nsw` n-tiles whose B panel (K x nsw*BN doubles) fits
// in a third of the 126 MB L2, all m-tiles inside a super-column, n fastest -- so B is read from
// DRAM once and A once per super-column instead of B once per m-tile.
CUPPEN_HD int work_supercol(const WorkCtx& w, const GemmProblem& Pb) {
    const int ntn = (Pb.N + w.BN - 1) / w.BN;
    long nsw = (48L << 20) / ((long)(Pb.K > 0 ? Pb.K : 1) * 8 * w.BN);
    if (nsw < 1) nsw = 1;
    return nsw < ntn ? (int)nsw : ntn;
}
template <class Emit>
CUPPEN_HD void work_emit_tiles(const WorkCtx& w, const GemmProblem& Pb, int p, int t, Emit emit) {
    const int ntm = (Pb.M + w.BM - 1) / w.BM, ntn = (Pb.N + w.BN - 1) / w.BN;
    const int nsw = work_supercol(w, Pb);
    for (int ns = 0; ns < ntn; ns += nsw)
        for (int mt = 0; mt < ntm; ++mt)
            for (int nt = ns; nt < ns + nsw && nt < ntn; ++nt) {
                if (t < w.tile_cap) emit(t, GemmTile{p, mt * w.BM, nt * w.BN});
                ++t;
            }
}

CUPPEN_HD int work_fill_problem(const WorkCtx& w, int p, GemmProblem& Pb) {
    const MergeDesc& D = w.desc[p >> 1];
    const int half = p & 1;
    Pb.M = 0; Pb.N = 0; Pb.K = 0;
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
// one block; shared memory: 2*nd ints (tile offsets)
__global__ void __launch_bounds__(256) build_gemm_work_kernel(WorkCtx w) {
    extern __shared__ int work_off[];
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
    if (threadIdx.x == 0) {
        int run = 0;
        for (int p = 0; p < np; ++p) { int c = work_off[p]; work_off[p] = run; run += c; }
        w.ntiles[0] = run < w.tile_cap ? run : w.tile_cap;
        w.ntiles[1] = misaligned;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        const GemmProblem Pb = w.probs[p];
        if (Pb.M == 0) continue;
        GemmTile* tl = w.tiles;
        work_emit_tiles(w, Pb, p, work_off[p], [tl](int t, GemmTile T) { tl[t] = T; });
    }
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
                                                      int R0, int* __restrict__ fail, int compact) {
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
                    r = sqrt(h2);
                    const double rinv = 1.0 / r;
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
    }
    const int grow = off + lane;          // global row of this lane
    if (Q != nullptr && lane < nl)
        for (int c = 0; c < nl; ++c) Q[(long)(grow - R0) + (long)(compact ? c : off + c) * ldq] = q[lane][c];   // compact: (row, c) at row + c*ldq
}

// K3: one warp per secular root; poles and weights of the merge staged in shared memory when
// they fit (k <= SEC_SMEM_K), otherwise read through L1/L2.
enum { SEC_SMEM_K = 14336, SEC_WARPS = 16 };
__global__ void __launch_bounds__(SEC_WARPS * 32) secular_kernel(LevelCtx c, int kcap, int part, int nparts) {
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
    if (k <= kcap) {
        for (int t = threadIdx.x; t < k; t += blockDim.x) { sec_smem[t] = dl[t]; sec_smem[kcap + t] = wl[t]; }
        __syncthreads();
        dl = sec_smem;
        wl = sec_smem + kcap;
    }
    if (i >= i1) return;
    WarpLanes L;
    SecularRoot r = secular_solve(L, k, dl, wl, fabs(D.rho), D.sumw, i);
    if (L.lane() == 0) { c.org[D.off + i] = r.origin; c.tau[D.off + i] = r.tau; }
}

// Fused front end for the small merges at the bottom of the tree: one CTA per merge runs every vector
// stage (z assembly ... new eigenvalues, and in eigenvalue-only mode the boundary-row update) back
// to back with block barriers in between, instead of ~12 separate launches per level.
enum { FUSE_MAXM = 128, FUSE_THREADS = 512 };
__global__ void __launch_bounds__(FUSE_THREADS) fused_front_kernel(LevelCtx c, RowCtx rc, int rows_mode) {
    const int id = blockIdx.x;
    const int off = c.desc[id].off, m = c.desc[id].m;
    const int tid = threadIdx.x, warp = tid >> 5, nwarps = FUSE_THREADS / 32;
    const WarpLanes L;
    for (int g = off + tid; g < off + m; g += FUSE_THREADS) ZAssemble{c}(g);
    __syncthreads();
    if (warp == 0) MergeTol{c}(id, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += FUSE_THREADS) FlagDeflate{c}(g);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) RankLive{c}(g, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += FUSE_THREADS) GivensSweep{c}(g);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) Compact{c}(g, L);
    __syncthreads();
    {
        const MergeDesc& D = c.desc[id];
        const int k = D.k;
        for (int i = warp; i < k; i += nwarps) {
            SecularRoot r = secular_solve(L, k, c.dl + off, c.wl + off, fabs(D.rho), D.sumw, i);
            if (L.lane() == 0) { c.org[off + i] = r.origin; c.tau[off + i] = r.tau; }
        }
    }
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) Loewner{c}(g, L);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) Norms{c}(g, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += FUSE_THREADS) NewLambda{c}(g);
    if (!rows_mode) return;
    __syncthreads();
    for (int g = off + tid; g < off + m; g += FUSE_THREADS) RowPack{c, rc}(g);
    __syncthreads();
    for (int g = off + warp; g < off + m; g += nwarps) RowGemv{c, rc}(g, L);
    __syncthreads();
    for (int g = off + tid; g < off + m; g += FUSE_THREADS) RowCommit{c, rc, const_cast<double*>(c.frow), const_cast<double*>(c.lrow)}(g);
}

// K5a: walk every rotation chain once per row.  grid.x = global column, grid.y = row chunk.
// Columns without a rotation (z-deflated, or live and unrotated: the common cases) take a fast path
// that issues the loads of all PACK_ROWS rows before the stores.
enum { PACK_THREADS = 128, PACK_ROWS = 4 };
__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(LevelCtx c, MatCtx M) {
    const int g = blockIdx.x;
    const int id = c.node_of[g];
    if (id < 0) return;
    const MergeDesc& D = c.desc[id];
    const int off = D.off, e = g - off;
    const int Gg = c.G[g];
    const bool zdefl = (Gg == -2);
    if (!zdefl && !c.head[g]) return;
    const int rbase = D.lr0 + blockIdx.y * PACK_ROWS * PACK_THREADS + threadIdx.x;      // local row
    if (rbase >= D.lr1) return;
    const bool etop = e < D.n1;
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
            v[t] = (r < D.lr1 && ((r < D.lsplit) == etop)) ? src[r] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < PACK_ROWS; ++t) {
            const int r = rbase + t * PACK_THREADS;
            if (r < D.lr1 && (write_other || ((r < D.lsplit) == etop))) dst[r] = v[t];
        }
        return;
    }
    for (int t = 0; t < PACK_ROWS; ++t) {
        const int r = rbase + t * PACK_THREADS;
        if (r >= D.lr1) return;
        const long rl = r;
        const bool rtop = r < D.lsplit;
        int a = e;
        double carry = ((a < D.n1) == rtop) ? M.Qold[rl + (long)(off + a) * M.ldq] : 0.0;
        int b;
        while ((b = c.G[off + a]) >= 0) {
            const double cs = c.gc[off + a], sn = c.gs[off + a];
            const double x = ((b < D.n1) == rtop) ? M.Qold[rl + (long)(off + b) * M.ldq] : 0.0;
            M.Qnew[rl + (long)(off + a) * M.ldq] = cs * carry - sn * x;
            carry = sn * carry + cs * x;
            a = b;
        }
        const int pos = rtop ? c.tpos[off + a] : c.bpos[off + a];
        if (pos >= 0) M.Apack[rl + (long)(off + pos) * M.ldq] = carry;
    }
}

// zero the K tail of Apack: columns [kh, round_up(kh,K_PAD)) of each half.  grid.x = descriptor,
// grid.y = row chunk.
__global__ void __launch_bounds__(128) pack_tail_kernel(LevelCtx c, MatCtx M) {
    const MergeDesc& D = c.desc[blockIdx.x];
    const int r = D.lr0 + blockIdx.y * 128 + threadIdx.x;      // local row
    if (r >= D.lr1) return;
    const bool rtop = r < D.lsplit;
    const int kh = rtop ? D.ktop : D.kbot;
    const int kend = (kh + K_PAD - 1) / K_PAD * K_PAD;
    for (int kk = kh; kk < kend; ++kk) M.Apack[(long)r + (long)(D.off + kk) * M.ldq] = 0.0;
}

// K5b: B[row = arena row of pole j][col = root i - p0] = zhat_j / (((d_j - d_org(i)) - tau_i) N_i)
// grid.x = arena row (global index), grid.y = a few column lanes; a block strides over the 256-wide
// column chunks of its row (the live count k is only known on the device).
__global__ void __launch_bounds__(256) ugen_kernel(LevelCtx c, MatCtx M, int p0, int width) {
    const int row = blockIdx.x;
    const int id = c.node_of[row];
    if (id < 0) return;
    const MergeDesc& D = c.desc[id];
    const bool top = row < D.off + D.n1;
    const int jj = row - (top ? D.off : D.off + D.n1);
    const int kh = top ? D.ktop : D.kbot;
    if (jj >= kh) return;
    const int iend = min(D.k, p0 + width);
    const int j = top ? c.toplist[row] : c.botlist[row];
    const double* dl = c.dl + D.off;
    const double dj = dl[j], zj = c.zhat[D.off + j];
    double* out = M.B + (long)row * M.ldb - p0;
    for (int i = p0 + blockIdx.y * 256 + threadIdx.x; i < iend; i += gridDim.y * 256) {
        double den = ((dj - dl[c.org[D.off + i]]) - c.tau[D.off + i]) * c.nrm[D.off + i];
        if (den == 0.0) den = 4.9e-324;
        double v = zj / den;
        if (!(fabs(v) < 1.7e308)) v = (v > 0) ? 1.7e308 : -1.7e308;
        out[i] = v;
    }
}

// K8: one block per output column over one contiguous slice of rows: global rows [g0, g0+cnt) stored
// at local rows [l0, l0+cnt).  Output column c (ascending lambda) is storage column perm[c] of V.  Four
// independent rows per thread and iteration keep enough loads in flight to stream from HBM.
// `accumulate` adds to res2 (a rank holds several slices when the rows are distributed).
__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ V, long ldq, int n, int g0, int l0, int cnt,
                                                       const double* __restrict__ OD, const double* __restrict__ OE,
                                                       const double* __restrict__ lam_sorted, const int* __restrict__ perm,
                                                       const double* __restrict__ halo_lo, const double* __restrict__ halo_hi,
                                                       double* __restrict__ res2, int accumulate) {
    const int col = blockIdx.x;
    const double* x = V + (long)perm[col] * ldq + l0 - g0;      // x[r] = element of global row r
    const double lambda = lam_sorted[col];
    const int g1 = g0 + cnt;
    double acc = 0;
    for (int r0 = g0 + threadIdx.x; r0 < g1; r0 += 4 * 256) {
        double xm[4], xc[4], xp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * 256;
            xc[u] = (r < g1) ? x[r] : 0.0;
            xm[u] = (r < g1 && r > 0) ? ((r > g0) ? x[r - 1] : halo_lo[col]) : 0.0;
            xp[u] = (r < g1 && r < n - 1) ? ((r + 1 < g1) ? x[r + 1] : halo_hi[col]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * 256;
            if (r >= g1) continue;
            double y = OD[r] * xc[u] - lambda * xc[u];
            if (r > 0) y += OE[r - 1] * xm[u];
            if (r < n - 1) y += OE[r] * xp[u];
            acc += y * y;
        }
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        res2[col] = accumulate ? res2[col] + s : s;
    }
}

__global__ void __launch_bounds__(256) gather_cols_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                          long ldq, int rows, const int* __restrict__ perm) {
    const int cidx = blockIdx.x;
    const double* s = src + (long)perm[cidx] * ldq;
    double* d = dst + (long)cidx * ldq;
    for (int r = blockIdx.y * 256 + threadIdx.x; r < rows; r += gridDim.y * 256) d[r] = s[r];
}
#endif  // CUPPEN_CUDA

}  // namespace cuppen
#endif
