// The O(m) / O(m^2) "vector" stages of one level of merges (z assembly, deflation, Loewner
// vector, normalisation, new eigenvalues).  Every stage is a one-thread-per-item functor run by
// launch_items() over the global index range [0,n): all merges of a level own disjoint index
// ranges, `node_of[g]` says which MergeDesc (if any) index g belongs to.
//
// What they replace in the reference (paths under /root/reference):
//   ZAssemble      computeZ                         src/helper.c:36-50, D copy src/main.c:534-538
//   FlagDeflate    z-deflation                      src/eigenvalues.c:58-81
//   RankLive       qsort of {D[i],i}                src/eigenvalues.c:83-88 (+ src/helper.c:95-103)
//   GivensSweep    sequential Givens chain          src/eigenvalues.c:98-135
//   Compact        (implicit: G[] tests everywhere) src/eigenvalues.c:166-208
//   Loewner        -- absent in the reference (Gu/Eisenstat z-hat; SURVEY.md finding 4)
//   Norms          computeNormalizationFactors      src/eigenvalues.c:257-289
//   NewLambda      L[ind] = ...                     src/eigenvalues.c:170-171,244
#ifndef CUPPEN_MERGE_STAGES_H
#define CUPPEN_MERGE_STAGES_H

#include <math.h>
#include "platform.h"
#include "p2p.h"

namespace cuppen {

enum { MODE_ACCURATE = 0, MODE_REFERENCE = 1 };
enum { SUP_TOP = 1, SUP_BOT = 2 };

// One merge node of the current level (host fills the first block, device the second).
struct MergeDesc {
    int off, n1, n2, m;
    int mode;
    int pad0;
    double rho;      // rank-one weight as seen by the solver: beta*theta (reference rule) or 2*beta
    double theta;    // z = [last row of Q1 ; first row of Q2 / theta]
    double zscale;   // 1 (reference rule) or 1/sqrt(2) (accurate rule, LAPACK dlaed2 convention)
    // rows of this node's Q block held by this rank: local rows [lr0, lr1), the lower half starts at lsplit
    int lr0, lsplit, lr1;
    int own_first, own_last;   // this rank holds the node's first / last global row
    double sigma;              // max |entry| of T: LAPACK's dstedc scales T to unit norm before its
                               // tolerances apply; we keep T unscaled and scale the tolerance instead
    double dthr;               // reference rule: the absolute pole-gap threshold 1e-5 (src/eigenvalues.c:109) in the
                               // units of the uploaded matrix (1e-5 * s when T was scaled by the power of two s)
    // ---- written by the device ----
    int nlive1;      // entries that survive z-deflation
    int k;           // live entries after the Givens sweep = number of secular roots
    int ktop, kbot;  // live columns with support in the upper / lower half
    double tol;      // accurate rule: 8 eps max(|d|max,|z|max)
    double sumw;     // sum of z^2 over the live entries
};

struct LevelCtx {
    int n;                 // global problem size
    MergeDesc* desc;       // descriptors of this level
    const int* node_of;    // [n] descriptor id or -1
    double* lam;           // [n] eigenvalues of the current nodes (children in, parents out)
    const double* frow;    // [n] first row of each current node's Q
    const double* lrow;    // [n] last row
    const double* Qz;      // vector mode on one GPU: the children's Q itself (boundary rows are read from it, no
    long ldqz;             //   ExtractRows pass); nullptr: use frow / lrow
    double* d;             // [n] poles in child order
    double* z;             // [n]
    double* dn;            // [n] poles after the Givens sweep
    double* zn;            // [n] z after the sweep
    int* G;                // [n] -1 live, -2 z-deflated, >=0 local index of the rotation partner
    double* gc;            // [n] Givens cosine, keyed by the deflated index
    double* gs;            // [n] Givens sine
    int* lsort;            // [n] off+p -> local index of the p-th z-live entry in ascending d
    int* head;             // [n] 1 if the element starts a rotation chain (singletons included)
    int* sup;              // [n] support bits of the chain ending at this element
    int* prev;             // [n] local index of the element rotated into this one (G[prev] == this), -1 at a chain head
    int* tpos;             // [n] per element: position in the top K-list or -1
    int* bpos;             // [n] per element: position in the bottom K-list or -1
    // canonical live problem (rho>0, poles ascending), indexed off+ci
    double* dl;            // poles
    double* wl;            // z^2
    double* zl;            // z (signed)
    int* lidx;             // local index of the element
    int* org;              // secular root: origin pole (canonical index)
    double* tau;           // secular root: lambda = dl[org] + tau
    double* dorgv;         // dl[org] of every root, stored next to (org, tau) by the root finder so that the U kernel reads
                           // three independent coalesced vectors instead of a dependent gather per column
    double* zhat;          // Loewner vector
    double* nrm;           // column norms
    int* toplist;          // [n] off+t        -> canonical index of the t-th top-supported live column
    int* botlist;          // [n] off+n1+t     -> canonical index of the t-th bottom-supported one
};

CUPPEN_HD bool before(double da, int ia, double db, int ib) {
    return (da < db) || (da == db && ia < ib);
}

struct ZAssemble {
    LevelCtx c;
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        int j = (int)g - D.off;
        if (j == 0) {       // device-written fields of the descriptor start from zero on every solve
            MergeDesc& W = c.desc[id];
            W.nlive1 = 0; W.k = 0; W.ktop = 0; W.kbot = 0; W.sumw = 0.0;
        }
        c.d[g] = c.lam[g];
        double zv;
        if (c.Qz != nullptr)    // last row of the upper child / first row of the lower child, straight from Q
            zv = (j < D.n1) ? c.Qz[(long)(D.lsplit - 1) + g * c.ldqz] : c.Qz[(long)D.lsplit + g * c.ldqz] / D.theta;
        else
            zv = (j < D.n1) ? c.lrow[g] : c.frow[g] / D.theta;
        c.z[g] = zv * D.zscale;
        c.G[g] = -1;
        c.head[g] = 0;
        c.sup[g] = 0;
        c.prev[g] = -1;
        c.tpos[g] = -1;
        c.bpos[g] = -1;
    }
};

// accurate rule: deflation tolerance 8 eps max(|d|max, sigma |z|max) of every merge (one warp per merge;
// dlaed2's rule for a matrix that dstedc has scaled by 1/sigma)
struct MergeTol {
    LevelCtx c;
    template <class L>
    CUPPEN_HD void operator()(long id, const L& lanes) const {
        MergeDesc& D = c.desc[id];
        if (D.mode == MODE_REFERENCE) return;
        double dmax = 0, zmax = 0;
        const double* dd = c.d + D.off;
        const double* zz = c.z + D.off;
        for (int t = lanes.lane(); t < D.m; t += lanes.lanes()) {
            dmax = fmax(dmax, fabs(dd[t]));
            zmax = fmax(zmax, fabs(zz[t]));
        }
        dmax = lanes.max(dmax);
        zmax = lanes.max(zmax);
        if (lanes.lane() == 0) D.tol = 8.0 * 2.220446049250313e-16 * fmax(dmax, D.sigma * zmax);
    }
};

// z-deflation flags
struct FlagDeflate {
    LevelCtx c;
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        double zg = c.z[g];
        bool defl;
        if (D.mode == MODE_REFERENCE) defl = fabs(zg) < 1e-6;                       // src/eigenvalues.c:72,77
        else defl = fabs(D.rho) * fabs(zg) <= D.tol;
        if (defl) {
            c.G[g] = -2;
            c.dn[g] = c.d[g];
            c.zn[g] = zg;
        }
    }
};

// levels without accurate-rule merges (no tolerance reduction needed): z assembly and deflation flags in one pass
struct ZAssembleFlag {
    LevelCtx c;
    CUPPEN_HD void operator()(long g) const {
        ZAssemble{c}(g);
        FlagDeflate{c}(g);
    }
};

// stable enumeration sort of the z-live entries by (d, index); one warp per element
struct RankLive {
    LevelCtx c;
    template <class L>
    CUPPEN_HD void operator()(long g, const L& lanes) const {
        int id = c.node_of[g];
        if (id < 0) return;
        MergeDesc& D = c.desc[id];
        int j = (int)g - D.off;
        const double* dd = c.d + D.off;
        const int* GG = c.G + D.off;
        double dj = dd[j];
        bool live = GG[j] != -2;
        if (!live && j != 0) return;
        int cnt = 0, tot = 0;
        for (int t = lanes.lane(); t < D.m; t += lanes.lanes()) {
            if (GG[t] == -2) continue;
            tot++;
            if (before(dd[t], t, dj, j)) cnt++;
        }
        cnt = lanes.isum(cnt);
        tot = lanes.isum(tot);
        if (lanes.lane() == 0) {
            if (live) c.lsort[D.off + cnt] = j;
            if (j == 0) D.nlive1 = tot;
        }
    }
};

// can the rotation step (p-1 -> p) fire, whatever chain state p-1 is in?  false = certified no.
CUPPEN_HD bool step_may_fire(const MergeDesc& D, double dprev, double zprev, double dq, double zq) {
    if (D.mode == MODE_REFERENCE) return fabs(dq - dprev) < D.dthr;    // src/eigenvalues.c:109
    double t = dq - dprev;
    double y = fabs(zq), x1 = fabs(zprev), x2 = 1.0 + 1e-9;
    double g1 = x1 * y / (x1 * x1 + y * y), g2 = x2 * y / (x2 * x2 + y * y);
    return !(t * fmin(g1, g2) > D.tol * (1.0 + 1e-9));
}

struct GivensSweep {
    LevelCtx c;
    CUPPEN_HD void finalize(int off, int e, double dc, double zc, int sp) const {
        c.dn[off + e] = dc;
        c.zn[off + e] = zc;
        c.sup[off + e] = sp;
    }
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        int p = (int)g - D.off;
        if (p >= D.nlive1) return;
        const int off = D.off;
        const int* ls = c.lsort + off;
        const double* dd = c.d + off;
        const double* zz = c.z + off;
        int e = ls[p];
        if (p > 0) {
            int ep = ls[p - 1];
            if (step_may_fire(D, dd[ep], zz[ep], dd[e], zz[e])) return;      // not a segment head
        }
        double dc = dd[e], zc = zz[e];
        int sp = (e < D.n1) ? SUP_TOP : SUP_BOT;
        c.head[off + e] = 1;
        int eprev = e;
        for (int q = p + 1; q < D.nlive1; ++q) {
            int eq = ls[q];
            double dq = dd[eq], zq = zz[eq];
            if (!step_may_fire(D, dd[eprev], zz[eprev], dq, zq)) break;    // q heads the next segment
            bool fire;
            double r = sqrt(zc * zc + zq * zq);                             // src/eigenvalues.c:113
            double cs = zq / r, sn = zc / r;
            if (D.mode == MODE_REFERENCE) fire = fabs(dq - dc) < D.dthr;
            else fire = fabs((dq - dc) * cs * sn) <= D.tol;
            if (fire) {
                c.G[off + e] = eq;
                c.prev[off + eq] = e;
                c.gc[off + e] = cs;
                c.gs[off + e] = sn;
                double ti = cs * cs * dc + sn * sn * dq;                    // :126-127
                double tj = sn * sn * dc + cs * cs * dq;
                c.dn[off + e] = ti;
                c.zn[off + e] = 0.0;
                dc = tj;
                zc = r;
                sp |= (eq < D.n1) ? SUP_TOP : SUP_BOT;
            } else {
                finalize(off, e, dc, zc, sp);
                dc = dq;
                zc = zq;
                sp = (eq < D.n1) ? SUP_TOP : SUP_BOT;
                c.head[off + eq] = 1;
            }
            e = eq;
            eprev = eq;
        }
        finalize(off, e, dc, zc, sp);
    }
};

// build the canonical live problem (rho>0, ascending poles) and the per-half K lists; one warp per
// position of the z-live sorted list
struct Compact {
    LevelCtx c;
    template <class L>
    CUPPEN_HD void operator()(long g, const L& lanes) const {
        int id = c.node_of[g];
        if (id < 0) return;
        MergeDesc& D = c.desc[id];
        int p = (int)g - D.off;
        if (p >= D.nlive1) return;
        const int off = D.off;
        const int* ls = c.lsort + off;
        int e = ls[p];
        if (c.G[off + e] != -1) return;
        int cnt = 0, k = 0, tcnt = 0, bcnt = 0, kt = 0, kb = 0;
        double sw = 0;
        for (int q = lanes.lane(); q < D.nlive1; q += lanes.lanes()) {
            int eq = ls[q];
            if (c.G[off + eq] != -1) continue;
            int s = c.sup[off + eq];
            double zq = c.zn[off + eq];
            sw += zq * zq;
            k++;
            kt += (s & SUP_TOP) ? 1 : 0;
            kb += (s & SUP_BOT) ? 1 : 0;
            if (q < p) {
                cnt++;
                tcnt += (s & SUP_TOP) ? 1 : 0;
                bcnt += (s & SUP_BOT) ? 1 : 0;
            }
        }
        cnt = lanes.isum(cnt); k = lanes.isum(k); tcnt = lanes.isum(tcnt); bcnt = lanes.isum(bcnt);
        kt = lanes.isum(kt); kb = lanes.isum(kb);
        if (cnt == 0) sw = lanes.sum(sw);
        if (lanes.lane() != 0) return;
        const bool neg = D.rho < 0;
        int ci = neg ? (k - 1 - cnt) : cnt;
        double dv = c.dn[off + e], zv = c.zn[off + e];
        c.dl[off + ci] = neg ? -dv : dv;
        c.zl[off + ci] = zv;
        c.wl[off + ci] = zv * zv;
        c.lidx[off + ci] = e;
        int s = c.sup[off + e];
        if (s & SUP_TOP) { c.tpos[off + e] = tcnt; c.toplist[off + tcnt] = ci; }
        if (s & SUP_BOT) { c.bpos[off + e] = bcnt; c.botlist[off + D.n1 + bcnt] = ci; }
        if (cnt == 0) { D.k = k; D.ktop = kt; D.kbot = kb; D.sumw = sw; }
    }
};

// Gu/Eisenstat: z-hat such that the computed roots are the exact eigenvalues of D + rho zhat zhat^T
// (one warp per pole, the product over the roots is split across the lanes)
struct Loewner {
    LevelCtx c;
    template <class L>
    CUPPEN_HD void operator()(long g, const L& lanes) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        int j = (int)g - D.off;
        if (j >= D.k) return;
        const double* dl = c.dl + D.off;
        const double* tau = c.tau + D.off;
        const int* org = c.org + D.off;
        const double dj = dl[j];
        double prod = 1.0;
        for (int i = lanes.lane(); i < D.k; i += lanes.lanes()) {
            if (i == j) continue;
            prod *= CUPPEN_DIV((dl[org[i]] - dj) + tau[i], dl[i] - dj);
        }
        prod = lanes.prod(prod) * ((dl[org[j]] - dj) + tau[j]);
        double zh = sqrt(fabs(prod) / fabs(D.rho));
        if (lanes.lane() == 0) c.zhat[g] = (c.zl[g] < 0) ? -zh : zh;
    }
};

struct Norms {
    LevelCtx c;
    template <class L>
    CUPPEN_HD void operator()(long g, const L& lanes) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        int i = (int)g - D.off;
        if (i >= D.k) return;
        const double* dl = c.dl + D.off;
        const double* zh = c.zhat + D.off;
        const double dorg = dl[c.org[g]], t = c.tau[g];
        double s = 0;
        for (int j = lanes.lane(); j < D.k; j += lanes.lanes()) {
            double u = CUPPEN_DIV(zh[j], (dl[j] - dorg) - t);
            s += u * u;
        }
        s = lanes.sum(s);
        if (lanes.lane() == 0) c.nrm[g] = sqrt(s);
    }
};

// collective fall-back on several ranks: (org, tau) of the other ranks' root ranges arrive through an all-reduce; the
// origin poles that go with them are looked up locally (the peer-memory back end pushes dorgv next to org and tau)
struct FillDorg {
    LevelCtx c;
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        if ((int)g - D.off < D.k) c.dorgv[g] = c.dl[D.off + c.org[g]];
    }
};

struct NewLambda {
    LevelCtx c;
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        int j = (int)g - D.off;
        if (c.G[g] != -1) c.lam[g] = c.dn[g];             // deflated: src/eigenvalues.c:170-171
        if (j < D.k) {
            double v = c.dl[D.off + c.org[g]] + c.tau[g];
            c.lam[D.off + c.lidx[g]] = (D.rho < 0) ? -v : v;
        }
    }
};

// ---- K7: boundary-row propagation for the eigenvalue-only mode -------------------------------------
// The reference never forms Q above the leaves while computing eigenvalues: it carries only the
// first and last row of every node, Wf = Q1f U[0:n1,:], Wl = Q2l U[n1:,:] (src/main.c:613-639).
struct RowCtx {
    const double* frow_old;
    const double* lrow_old;
    double* frow_new;
    double* lrow_new;
    double* fpack;     // [n] off+t      : rotated first-row entries of the top-supported live columns
    double* lpack;     // [n] off+n1+t   : rotated last-row entries of the bottom-supported live columns
};

struct RowPack {
    LevelCtx c;
    RowCtx r;
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        const int off = D.off, e = (int)g - off, n1 = D.n1;
        if (c.G[g] == -2) {
            r.frow_new[g] = (e < n1) ? r.frow_old[g] : 0.0;
            r.lrow_new[g] = (e >= n1) ? r.lrow_old[g] : 0.0;
            return;
        }
        if (!c.head[g]) return;
        int a = e, b;
        double cf = (a < n1) ? r.frow_old[off + a] : 0.0;
        double cl = (a >= n1) ? r.lrow_old[off + a] : 0.0;
        while ((b = c.G[off + a]) >= 0) {
            const double cs = c.gc[off + a], sn = c.gs[off + a];
            const double xf = (b < n1) ? r.frow_old[off + b] : 0.0;
            const double xl = (b >= n1) ? r.lrow_old[off + b] : 0.0;
            r.frow_new[off + a] = cs * cf - sn * xf;
            r.lrow_new[off + a] = cs * cl - sn * xl;
            cf = sn * cf + cs * xf;
            cl = sn * cl + cs * xl;
            a = b;
        }
        if (c.tpos[off + a] >= 0) r.fpack[off + c.tpos[off + a]] = cf;
        if (c.bpos[off + a] >= 0) r.lpack[off + n1 + c.bpos[off + a]] = cl;
    }
};

struct RowGemv {
    LevelCtx c;
    RowCtx r;
    template <class L>
    CUPPEN_HD void operator()(long g, const L& lanes) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        const int off = D.off, i = (int)g - off;
        if (i >= D.k) return;
        const double* dl = c.dl + off;
        const double* zh = c.zhat + off;
        const double dorg = dl[c.org[g]], t = c.tau[g], nn = c.nrm[g];
        double sf = 0, sl = 0;
        for (int q = lanes.lane(); q < D.ktop; q += lanes.lanes()) {
            int j = c.toplist[off + q];
            sf += r.fpack[off + q] * CUPPEN_DIV(zh[j], ((dl[j] - dorg) - t) * nn);
        }
        for (int q = lanes.lane(); q < D.kbot; q += lanes.lanes()) {
            int j = c.botlist[off + D.n1 + q];
            sl += r.lpack[off + D.n1 + q] * CUPPEN_DIV(zh[j], ((dl[j] - dorg) - t) * nn);
        }
        sf = lanes.sum(sf);
        sl = lanes.sum(sl);
        if (lanes.lane() == 0) {
            r.frow_new[off + c.lidx[g]] = sf;
            r.lrow_new[off + c.lidx[g]] = sl;
        }
    }
};

// copy the new boundary rows of this level's nodes over the old ones (other index ranges keep theirs)
struct RowCommit {
    LevelCtx c;
    RowCtx r;
    double* frow;
    double* lrow;
    CUPPEN_HD void operator()(long g) const {
        if (c.node_of[g] < 0) return;
        frow[g] = r.frow_new[g];
        lrow[g] = r.lrow_new[g];
    }
};

// boundary rows of the freshly merged nodes, read back from the materialised Q (vector mode)
struct ExtractRows {
    LevelCtx c;
    const double* Q;
    long ldq;
    double* frow;
    double* lrow;
    SymHeap H;             // peer-memory back end: the owner of a boundary row stores it into every rank's copy
    CUPPEN_HD void operator()(long g) const {
        int id = c.node_of[g];
        if (id < 0) return;
        const MergeDesc& D = c.desc[id];
        if (D.lr1 <= D.lr0) return;
        if (D.own_first) {
            const double v = Q[(long)D.lr0 + g * ldq];
            frow[g] = v;
            for (int p = 0; p < H.G; ++p) if (p != H.me) H.at(p, frow)[g] = v;
        }
        if (D.own_last) {
            const double v = Q[(long)(D.lr1 - 1) + g * ldq];
            lrow[g] = v;
            for (int p = 0; p < H.G; ++p) if (p != H.me) H.at(p, lrow)[g] = v;
        }
    }
};

// final ordering: stable enumeration sort of all n eigenvalues (src/filehandling.c:315-321)
struct FinalRank {
    int n;
    const double* lam;
    int* perm;
    double* lam_sorted;
    template <class L>
    CUPPEN_HD void operator()(long g, const L& lanes) const {
        const double v = lam[g];
        int cnt = 0;
        // one comparison per pair: indices below g come first on ties (`<=`), indices above do not (`<`)
        const int nl = lanes.lanes(), me = (int)g;
        for (int j = lanes.lane(); j < me; j += nl) cnt += (lam[j] <= v) ? 1 : 0;
        for (int j = me + 1 + (lanes.lane() + nl - (me + 1) % nl) % nl; j < n; j += nl) cnt += (lam[j] < v) ? 1 : 0;
        cnt = lanes.isum(cnt);
        if (lanes.lane() == 0) { perm[cnt] = (int)g; lam_sorted[cnt] = v; }
    }
};

// one local row of V in ascending-lambda column order
struct ExtractRowVec {
    const double* Q;
    long ldq;
    long rowlocal;
    const int* perm;
    double* out;
    CUPPEN_HD void operator()(long c) const { out[c] = Q[rowlocal + (long)perm[c] * ldq]; }
};

}  // namespace cuppen
#endif
