// Selected-eigenvector mode (the reference's `-eFILE`, /root/reference/src/filehandling.c:165-239,
// 339-345): eigenvectors of a few requested eigenvalues WITHOUT forming any n x n matrix.
//
// The reference back-transforms one eigenvector at a time, x = Q_leaves U_{d-2} ... U_1 U_0[:,i]
// (src/filehandling.c:332-508), re-deriving rows of the product lazily: O(n^3) per vector.  Here the
// solve runs in eigenvalue-only mode (boundary rows only, src/main.c:613-639) and keeps the O(n)
// vectors that define every level's U (poles, z-hat, (origin, tau), norms, Givens chains, index
// lists); the selected columns are then pushed through the tree from the root down to the leaves in
// *coefficient space*:
//
//     x_parent  (coefficients over the columns of the merged node's Q)
//       -> live roots:  gamma_j = zhat_j * sum_i  x[lidx_i] / (N_i ((d_j - d_org(i)) - tau_i))     (Cauchy-like product,
//                                                                                                   U generated on the fly)
//       -> rotation chains walked backwards (transpose of the forward walk of pack_kernel)
//       -> z-deflated columns copied
//     x_children (coefficients over the columns of diag(Q1, Q2))
//
// and finally x = Q_leaf x_leaf.  Cost 2 n^2 pole/root pairs per pass of up to SEL_NV vectors
// (one fp64 reciprocal + SEL_NV fma each) instead of (4/3) n^3 flop, memory O(n levels).
//
//   ApplyInit / ApplyPrep / ApplyChains / LeafApply   per-item functors (thread per global index)
//   cauchy_apply_kernel                               the hot kernel: thread per pole, roots staged in smem
//   SelResidual                                        ||T x - lambda x||_2 of the selected vectors
#ifndef CUPPEN_SELECT_STAGES_H
#define CUPPEN_SELECT_STAGES_H

#include "merge_stages.h"

namespace cuppen {

enum { SEL_NV = 8 };            // vectors pushed through the tree together (accumulators per thread)

struct SelCtx {
    int n;                 // global size; vector v of a pass lives at [v*n, (v+1)*n)
    int nv;                // vectors of this pass (<= SEL_NV)
    const int* sel;        // [nv] requested ranks (ascending-lambda order, 0-based)
    const int* perm;       // rank -> storage column of the root node
    double* X;             // coefficients over the current level's columns (in)
    double* Y;             // coefficients over the children's columns (out)
    double* XS;            // X gathered into canonical root order and divided by the column norm
    double* Gam;           // Cauchy product, canonical pole order
    double* dorg;          // [n] dl[org[i]] per canonical root
};

// X = unit vectors of the selected storage columns of the root
struct ApplyInit {
    SelCtx s;
    CUPPEN_HD void operator()(long g) const {
        for (int v = 0; v < s.nv; ++v) s.X[(long)v * s.n + g] = (s.perm[s.sel[v]] == (int)g) ? 1.0 : 0.0;
    }
};

// z-deflated columns pass through, untouched index ranges keep their coefficients, and the live
// roots' coefficients are gathered (scaled by 1/N_i) for the Cauchy product
struct ApplyPrep {
    LevelCtx c;
    SelCtx s;
    CUPPEN_HD void operator()(long g) const {
        const int id = c.node_of[g];
        if (id < 0) {
            for (int v = 0; v < s.nv; ++v) s.Y[(long)v * s.n + g] = s.X[(long)v * s.n + g];
            return;
        }
        const MergeDesc& D = c.desc[id];
        const int off = D.off, e = (int)g - off;
        if (c.G[g] == -2)
            for (int v = 0; v < s.nv; ++v) s.Y[(long)v * s.n + g] = s.X[(long)v * s.n + g];
        if (e < D.k) {
            const double rn = 1.0 / c.nrm[g];
            const int src = off + c.lidx[g];
            for (int v = 0; v < s.nv; ++v) s.XS[(long)v * s.n + g] = s.X[(long)v * s.n + src] * rn;
            s.dorg[g] = c.dl[off + c.org[g]];
        }
    }
};

// transpose of the forward chain walk (pack_kernel / getEigenVector's inverse rotations,
// src/eigenvalues.c:343-357): thread per canonical live pole j, whose chain ends at element lidx[j]
struct ApplyChains {
    LevelCtx c;
    SelCtx s;
    CUPPEN_HD void operator()(long g) const {
        const int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        const int off = D.off, j = (int)g - off;
        if (j >= D.k) return;
        double gam[SEL_NV];
#pragma unroll
        for (int v = 0; v < SEL_NV; ++v) gam[v] = (v < s.nv) ? s.Gam[(long)v * s.n + g] : 0.0;
        int a = c.lidx[g], p;
        while ((p = c.prev[off + a]) >= 0) {
            const double cs = c.gc[off + p], sn = c.gs[off + p];
#pragma unroll
            for (int v = 0; v < SEL_NV; ++v) {
                if (v >= s.nv) break;
                const double xa = s.X[(long)v * s.n + off + p];
                s.Y[(long)v * s.n + off + a] = gam[v] * cs - xa * sn;
                gam[v] = gam[v] * sn + xa * cs;
            }
            a = p;
        }
#pragma unroll
        for (int v = 0; v < SEL_NV; ++v)
            if (v < s.nv) s.Y[(long)v * s.n + off + a] = gam[v];
    }
};

// x[row] = sum_c Qleaf[row][c] * X[leaf offset + c]; Qleaf is stored column-of-leaf major: (row, c) at row + c*n
struct LeafApply {
    SelCtx s;
    const int* leaf_off;   // [n] first index of the leaf that owns the row
    const int* leaf_n;     // [n] its size
    const double* Qleaf;
    double* out;           // [nv][n]
    CUPPEN_HD void operator()(long g) const {
        const int off = leaf_off[g], nl = leaf_n[g];
        double acc[SEL_NV];
#pragma unroll
        for (int v = 0; v < SEL_NV; ++v) acc[v] = 0.0;
        for (int cidx = 0; cidx < nl; ++cidx) {
            const double q = Qleaf[g + (long)cidx * s.n];
#pragma unroll
            for (int v = 0; v < SEL_NV; ++v)
                if (v < s.nv) acc[v] = fma(q, s.X[(long)v * s.n + off + cidx], acc[v]);
        }
#pragma unroll
        for (int v = 0; v < SEL_NV; ++v)
            if (v < s.nv) out[(long)v * s.n + g] = acc[v];
    }
};

// squared residual of selected vector t (src/filehandling.c:511-531); one warp per vector
struct SelResidual {
    int n;
    const double* V;       // [cnt][n]
    const double* OD;
    const double* OE;
    const double* lam_sorted;
    const int* sel;
    double* res2;          // [cnt]
    template <class L>
    CUPPEN_HD void operator()(long t, const L& lanes) const {
        const double* x = V + t * n;
        const double lambda = lam_sorted[sel[t]];
        double acc = 0;
        for (int r = lanes.lane(); r < n; r += lanes.lanes()) {
            double y = OD[r] * x[r] - lambda * x[r];
            if (r > 0) y += OE[r - 1] * x[r - 1];
            if (r < n - 1) y += OE[r] * x[r + 1];
            acc += y * y;
        }
        acc = lanes.sum(acc);
        if (lanes.lane() == 0) res2[t] = acc;
    }
};

// ---- the Cauchy-like product ------------------------------------------------------------------------
// Gam[v][j] = zhat_j * sum_i XS[v][i] / ((dl_j - dorg_i) - tau_i)    for the live poles j of every merge
// of the level.  grid.x = tile of CA_TJ poles, grid.y = merge.  A thread owns one pole and SEL_NV
// accumulators; the roots are staged in shared memory CA_RC at a time ((dorg, tau) and the SEL_NV
// right-hand sides of a root are contiguous, so the inner loop is broadcast LDS.128 + one fp64
// reciprocal + SEL_NV DFMA per pair).  Chunks whose right-hand sides are all zero are skipped: at the
// root of the tree X is a set of unit vectors, so only the chunks holding a selected root do work.
enum { CA_TJ = 64, CA_SL = 8, CA_THREADS = CA_TJ * CA_SL, CA_RC = 512, CA_SUB = CA_RC / CA_SL };   // (32 x 16 measured 10-20 % slower)

CUPPEN_HD double cauchy_recip(double dj, double dorg, double tau) {
    double diff = (dj - dorg) - tau;
    if (diff == 0.0) diff = 1e-300;             // same guard as ugen_kernel: never 0/0
    return 1.0 / diff;
}

#if CUPPEN_CUDA
// thread = (pole of the CTA's tile, slice of every staged chunk); the CA_SL partial sums of a pole are combined
// through shared memory in a fixed order
__global__ void __launch_bounds__(CA_THREADS) cauchy_apply_kernel(LevelCtx c, SelCtx s) {
    __shared__ double2 sdt[CA_RC];                       // (dorg, tau)
    __shared__ __align__(16) double sx[CA_RC][SEL_NV];
    __shared__ double s_part[CA_SL][CA_TJ];
    const MergeDesc& D = c.desc[blockIdx.y];
    const int k = D.k, off = D.off;
    const int j0 = blockIdx.x * CA_TJ;
    if (j0 >= k) return;
    const int out = threadIdx.x & (CA_TJ - 1), slice = threadIdx.x / CA_TJ;
    const int j = j0 + out;
    const double dj = c.dl[off + (j < k ? j : k - 1)];
    double acc[SEL_NV];
#pragma unroll
    for (int v = 0; v < SEL_NV; ++v) acc[v] = 0.0;
    for (int i0 = 0; i0 < k; i0 += CA_RC) {
        const int cnt = min((int)CA_RC, k - i0);
        int nz = 0;
        for (int t = threadIdx.x; t < cnt; t += CA_THREADS) {
            sdt[t] = make_double2(s.dorg[off + i0 + t], c.tau[off + i0 + t]);
#pragma unroll
            for (int v = 0; v < SEL_NV; ++v) {
                const double x = (v < s.nv) ? s.XS[(long)v * s.n + off + i0 + t] : 0.0;
                sx[t][v] = x;
                nz |= (x != 0.0);
            }
        }
        nz = __syncthreads_or(nz);
        if (nz) {
            const int t1 = min(cnt, (slice + 1) * CA_SUB);
#pragma unroll 4
            for (int t = slice * CA_SUB; t < t1; ++t) {
                const double2 q = sdt[t];
                const double r = cauchy_recip(dj, q.x, q.y);
                const double2* xv = reinterpret_cast<const double2*>(sx[t]);
#pragma unroll
                for (int v = 0; v < SEL_NV / 2; ++v) {
                    const double2 xx = xv[v];
                    acc[2 * v] = fma(r, xx.x, acc[2 * v]);
                    acc[2 * v + 1] = fma(r, xx.y, acc[2 * v + 1]);
                }
            }
        }
        __syncthreads();
    }
    const double zh = c.zhat[off + (j < k ? j : k - 1)];
#pragma unroll
    for (int v = 0; v < SEL_NV; ++v) {
        if (v >= s.nv) break;                                   // block-uniform
        s_part[slice][out] = acc[v];
        __syncthreads();
        if (slice == 0 && j < k) {
            double a = s_part[0][out];
#pragma unroll
            for (int q = 1; q < CA_SL; ++q) a += s_part[q][out];
            s.Gam[(long)v * s.n + off + j] = zh * a;
        }
        __syncthreads();
    }
}
#else
// TEST-ONLY host twin (CUPPEN_HOST_EMULATION)
inline void cauchy_apply_host(LevelCtx c, SelCtx s, int ndesc) {
    for (int id = 0; id < ndesc; ++id) {
        const MergeDesc& D = c.desc[id];
        const int k = D.k, off = D.off;
        for (int j = 0; j < k; ++j) {
            const double dj = c.dl[off + j];
            double acc[SEL_NV] = {0};
            for (int i = 0; i < k; ++i) {
                const double r = cauchy_recip(dj, s.dorg[off + i], c.tau[off + i]);
                for (int v = 0; v < s.nv; ++v) acc[v] = fma(r, s.XS[(long)v * s.n + off + i], acc[v]);
            }
            for (int v = 0; v < s.nv; ++v) s.Gam[(long)v * s.n + off + j] = c.zhat[off + j] * acc[v];
        }
    }
}
#endif

// several ranks: gathered layout [rank][slot][n] (vector t of the caller's list sits at rank t % world, slot t / world)
// -> the caller's order
struct SelReorder {
    const double* Vgath;
    double* Vsel;
    const double* rgath;
    double* rsel;
    int n, world, per;
    CUPPEN_HD void operator()(long i) const {
        const long t = i / n, r = i - t * n;
        const long src = (t % world) * per + t / world;
        Vsel[i] = Vgath[src * n + r];
        if (r == 0) rsel[t] = rgath[src];
    }
};

}  // namespace cuppen
#endif
