// TEST-ONLY host twins of the named CUDA kernels in matrix_stages.h / gemm_dmma.h, compiled only
// with -DCUPPEN_HOST_EMULATION (tests/host/Makefile).  They let the container without a GPU
// exercise the orchestration in solver.cu end to end.  Never part of libcuppen_b200.so.
#ifndef CUPPEN_HOST_TWINS_H
#define CUPPEN_HOST_TWINS_H
#if !CUPPEN_CUDA

#include <algorithm>
#include "matrix_stages.h"

namespace cuppen {

inline int host_leaf_ql(int nl, double* d, double* e, double* q /* row-major nl x nl, identity in */) {
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < nl; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < nl - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 90) return 1 + l;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = sqrt(g * g + 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                bool early = false;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; early = true; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    for (int row = 0; row < nl; ++row) {
                        double f2 = q[row * nl + i + 1], q0 = q[row * nl + i];
                        q[row * nl + i + 1] = s * q0 + c * f2;
                        q[row * nl + i] = c * q0 - s * f2;
                    }
                }
                if (!early) { d[l] -= p; e[l] = g; e[m] = 0.0; }
            }
        } while (m != l);
    }
    for (int i = 0; i < nl - 1; ++i) {
        int kmin = i; double p = d[i];
        for (int k2 = i + 1; k2 < nl; ++k2) if (d[k2] < p) { kmin = k2; p = d[k2]; }
        if (kmin != i) {
            d[kmin] = d[i]; d[i] = p;
            for (int row = 0; row < nl; ++row) std::swap(q[row * nl + i], q[row * nl + kmin]);
        }
    }
    return 0;
}

inline void leaf_ql_host(const LeafDesc* leaves, int nleaves, const double* Dm, const double* E, double* lam,
                         double* frow, double* lrow, double* Q, long ldq, int R0, int* fail, int compact, RowSpan* span) {
    for (int leaf = 0; leaf < nleaves; ++leaf) {
        const int off = leaves[leaf].off, nl = leaves[leaf].n;
        std::vector<double> d(Dm + off, Dm + off + nl), e(nl, 0.0), q((size_t)nl * nl, 0.0);
        for (int i = 0; i < nl - 1; ++i) e[i] = E[off + i];
        for (int i = 0; i < nl; ++i) q[(size_t)i * nl + i] = 1.0;
        int rc = host_leaf_ql(nl, d.data(), e.data(), q.data());
        if (rc) *fail = off + rc;
        for (int c = 0; c < nl; ++c) {
            lam[off + c] = d[c];
            frow[off + c] = q[c];
            lrow[off + c] = q[(size_t)(nl - 1) * nl + c];
            if (span) span[off + c] = RowSpan{off, off + nl};
            if (Q) for (int r = 0; r < nl; ++r) Q[(long)(off + r - R0) + (long)(compact ? c : off + c) * ldq] = q[(size_t)r * nl + c];
        }
    }
}

inline void secular_host(LevelCtx c, int ndesc, int part, int nparts) {
    SerialLanes L;
    for (int id = 0; id < ndesc; ++id) {
        const MergeDesc& D = c.desc[id];
        const int k = D.k;
        const int per = (k + nparts - 1) / nparts;
        const int i0 = part * per, i1 = std::min(k, i0 + per);
        for (int i = i0; i < i1; ++i) {
            SecularRoot r = secular_solve(L, k, c.dl + D.off, c.wl + D.off, fabs(D.rho), D.sumw, i);
            c.org[D.off + i] = r.origin;
            c.dorgv[D.off + i] = c.dl[D.off + r.origin];
            c.tau[D.off + i] = r.tau;
        }
    }
}

inline void pack_host(LevelCtx c, MatCtx M) {
    for (int g = 0; g < c.n; ++g) {
        const int id = c.node_of[g];
        if (id < 0) continue;
        const MergeDesc& D = c.desc[id];
        const int off = D.off, e = g - off;
        const bool zdefl = c.G[g] == -2;
        if (M.span && !zdefl) M.span[g] = RowSpan{off, off + D.m};
        if (!zdefl && !c.head[g]) continue;
        for (int r = D.lr0; r < D.lr1; ++r) {
            const long rl = r;
            const bool rtop = r < D.lsplit;
            if (zdefl) {
                const bool mine = (e < D.n1) == rtop;
                M.Qnew[rl + (long)g * M.ldq] = mine ? M.Qold[rl + (long)g * M.ldq] : 0.0;
                continue;
            }
            int a = e, b;
            double carry = ((a < D.n1) == rtop) ? M.Qold[rl + (long)(off + a) * M.ldq] : 0.0;
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
}

inline void pack_tail_host(LevelCtx c, MatCtx M, int ndesc) {
    for (int id = 0; id < ndesc; ++id) {
        const MergeDesc& D = c.desc[id];
        for (int r = D.lr0; r < D.lr1; ++r) {
            const bool rtop = r < D.lsplit;
            const int kh = rtop ? D.ktop : D.kbot;
            const int kend = (kh + K_PAD - 1) / K_PAD * K_PAD;
            for (int kk = kh; kk < kend; ++kk) M.Apack[(long)r + (long)(D.off + kk) * M.ldq] = 0.0;
        }
    }
}

inline void ugen_host(LevelCtx c, MatCtx M, int p0, int width) {
    for (int row = 0; row < c.n; ++row) {
        const int id = c.node_of[row];
        if (id < 0) continue;
        const MergeDesc& D = c.desc[id];
        const bool top = row < D.off + D.n1;
        const int jj = row - (top ? D.off : D.off + D.n1);
        const int kh = top ? D.ktop : D.kbot;
        if (jj >= kh) continue;
        const int j = top ? c.toplist[row] : c.botlist[row];
        const double* dl = c.dl + D.off;
        const double dj = dl[j], zj = c.zhat[D.off + j];
        for (int i = p0; i < D.k && i < p0 + width; ++i) {
            double den = ((dj - dl[c.org[D.off + i]]) - c.tau[D.off + i]) * c.nrm[D.off + i];
            if (den == 0.0) den = 4.9e-324;
            double v = zj / den;
            if (!(fabs(v) < 1.7e308)) v = (v > 0) ? 1.7e308 : -1.7e308;
            M.B[(long)row * M.ldb + (i - p0)] = v;
        }
    }
}

inline void build_gemm_work_host(WorkCtx w) {
    // same list as build_gemm_work_body: the tiles of all problems in order, the last `tail` of them as two half tiles each
    const int np = 2 * w.nd;
    std::vector<int> off(np + 1, 0);
    int mis = 0;
    for (int p = 0; p < np; ++p) {
        GemmProblem Pb;
        off[p + 1] = off[p] + work_fill_problem(w, p, Pb);
        w.probs[p] = Pb;
        if (Pb.M > 0 && (Pb.a_row0 & 1)) mis++;
    }
    int run = off[np];
    int tail = work_tail_tiles(w, run);
    if (run + tail > w.tile_cap) { *w.fail = 1; tail = 0; if (run > w.tile_cap) run = w.tile_cap; }
    int p = 0;
    for (int t = 0; t < run + tail; ++t) {
        int half;
        const int src = work_entry_source(run, tail, t, &half);
        while (p > 0 && off[p] > src) --p;
        while (off[p + 1] <= src) ++p;
        w.tiles[t] = work_entry_tile(work_tile_at(w, w.probs[p], p, src - off[p]), half);
    }
    w.ntiles[0] = run + tail;
    w.ntiles[1] = mis;
}

// consumes the same (problem, tile) list as the device kernels
inline void gemm_host(const GemmProblem* probs, const GemmTile* tiles, const int* ntiles_ptr, int BM, int BN) {
    for (int t = 0; t < ntiles_ptr[0]; ++t) {
        const GemmProblem& P = probs[tiles[t].prob & GEMM_TILE_PROB_MASK];
        const int m0 = tiles[t].m0, n0 = tiles[t].n0;
        const int m1 = std::min(P.M, m0 + BM), n1 = std::min(P.N, n0 + ((tiles[t].prob & GEMM_TILE_HALF) ? 64 : BN));
        const int Kpad = (P.K + K_PAD - 1) / K_PAD * K_PAD;     // read the padded K like the device kernel
        for (int nn = n0; nn < n1; ++nn) {
            double* ccol = P.C + (long)P.colidx[nn] * P.ldc;
            for (int mm = m0; mm < m1; ++mm) ccol[mm] = 0.0;
            for (int kk = 0; kk < Kpad; ++kk) {
                const double b = P.B[(long)kk * P.ldb + nn];
                const double* acol = P.A + (long)kk * P.lda;
                for (int mm = m0; mm < m1; ++mm) ccol[mm] += acol[mm] * b;
            }
        }
    }
}

inline void residual_host(const double* V, long ldq, int n, int g0, int l0, int cnt, const double* OD, const double* OE,
                          const double* lam_sorted, const int* perm, const double* halo_lo, const double* halo_hi, double* res2,
                          int accumulate, const RowSpan* span) {
    const int g1 = g0 + cnt;
    for (int col = 0; col < n; ++col) {
        const double* x = V + (long)perm[col] * ldq + l0 - g0;
        const double lambda = lam_sorted[col];
        double acc = 0;
        const int ra = span ? std::max(g0, span[perm[col]].lo - 1) : g0, rb = span ? std::min(g1, span[perm[col]].hi + 1) : g1;
        if (span)      // (test build: the rows that the device kernel skips must hold exact zeros)
            for (int r = g0; r < g1; ++r)
                if ((r < span[perm[col]].lo || r >= span[perm[col]].hi) && x[r] != 0.0)
                    CUPPEN_THROW(CUPPEN_ERR_STATE, "column %d (storage %d) is non-zero at row %d outside its span [%d, %d)", col, perm[col], r,
                                 span[perm[col]].lo, span[perm[col]].hi);
        for (int r = ra; r < rb; ++r) {
            const double xc = x[r];
            double y = OD[r] * xc - lambda * xc;
            if (r > 0) y += OE[r - 1] * ((r > g0) ? x[r - 1] : halo_lo[col]);
            if (r < n - 1) y += OE[r] * ((r + 1 < g1) ? x[r + 1] : halo_hi[col]);
            acc += y * y;
        }
        res2[col] = accumulate ? res2[col] + acc : acc;
    }
}

inline void gather_cols_host(const double* src, double* dst, long ldq, int rows, const int* perm, int n) {
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < rows; ++r) dst[(long)c * ldq + r] = src[(long)perm[c] * ldq + r];
}

}  // namespace cuppen
#endif
#endif
