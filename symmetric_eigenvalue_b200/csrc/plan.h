// Host-side merge schedule: the divide tree and the divide phase.
//
// Replaces EVRepTree/initEVRepTree (/root/reference/src/backtransformation.c:28-114) and the
// divide loop of main (/root/reference/src/main.c:274-284,339-421).  The upper part of the tree
// is the reference's own P-leaf tree (same node sizes, same theta rule, same thresholds at the
// merges) so that results are reproducible against `mpirun -n P cuppens`; every reference leaf,
// which the reference hands to LAPACKE_dsteqr (src/main.c:460), is solved by a deeper D&C tree
// with LAPACK-grade tolerances down to leaves of <= leaf_max rows.
#ifndef CUPPEN_PLAN_H
#define CUPPEN_PLAN_H

#include <math.h>
#include <vector>

namespace cuppen {

struct PlanNode {
    int off = 0, n = 0;
    int left = -1, right = -1;     // children (node ids); -1/-1 for a leaf
    int n1 = 0;                    // size of the left child
    int height = 0;                // leaves 0
    int depth = 0;                 // root 0 (real splits only)
    int mode = 0;                  // MODE_ACCURATE / MODE_REFERENCE at the merge of this node
    double beta = 0, theta = 1, rho = 0, zscale = 1;
};

struct Plan {
    int n = 0, P = 1, leaf_max = 32;
    int root = -1;
    std::vector<PlanNode> nodes;
    std::vector<std::vector<int>> by_height;   // merge nodes per height (>=1)
    std::vector<int> leaves;
    std::vector<double> D;                     // diagonal after all splits
};

namespace detail {
inline int build_accurate(Plan& p, int off, int n, int depth) {
    PlanNode nd;
    nd.off = off; nd.n = n; nd.depth = depth; nd.mode = 0;
    int id = (int)p.nodes.size();
    p.nodes.push_back(nd);
    if (n > p.leaf_max) {
        int n1 = n / 2;
        int l = build_accurate(p, off, n1, depth + 1);
        int r = build_accurate(p, off + n1, n - n1, depth + 1);
        p.nodes[id].left = l; p.nodes[id].right = r; p.nodes[id].n1 = n1;
    }
    return id;
}
// reference tree over leaves [l0, l1) of the P-leaf partition; span = power-of-two width of this
// subtree in leaves (the reference's left-packed tree: left child takes the first span/2 leaves)
inline int build_reference(Plan& p, const std::vector<int>& loff, const std::vector<int>& lsz, int l0, int l1,
                           int span, int depth) {
    if (l1 - l0 == 1 && span == 1) return build_accurate(p, loff[l0], lsz[l0], depth);
    int half = span / 2;
    if (l1 - l0 <= half)          // single-child pass-through node (src/main.c:353,400-402)
        return build_reference(p, loff, lsz, l0, l1, half, depth);
    PlanNode nd;
    nd.off = loff[l0]; nd.depth = depth; nd.mode = 1;
    int id = (int)p.nodes.size();
    p.nodes.push_back(nd);
    int l = build_reference(p, loff, lsz, l0, l0 + half, half, depth + 1);
    int r = build_reference(p, loff, lsz, l0 + half, l1, half, depth + 1);
    p.nodes[id].left = l; p.nodes[id].right = r;
    p.nodes[id].n1 = p.nodes[l].n;
    p.nodes[id].n = p.nodes[l].n + p.nodes[r].n;
    return id;
}
inline int set_heights(Plan& p, int id) {
    PlanNode& nd = p.nodes[id];
    if (nd.left < 0) { nd.height = 0; p.leaves.push_back(id); return 0; }
    int hl = set_heights(p, nd.left), hr = set_heights(p, nd.right);
    int h = 1 + (hl > hr ? hl : hr);
    p.nodes[id].height = h;
    if ((int)p.by_height.size() <= h) p.by_height.resize(h + 1);
    p.by_height[h].push_back(id);
    return h;
}
}  // namespace detail

// The divide phase on an existing tree: top-down on the already modified diagonal (main.c:339-421).  The tree
// shape depends on (n, P, leaf_max) only, so a new matrix on the same handle needs nothing but this pass.
// D, E may have been scaled by a power of two s (matrices of extreme norm); inv_scale = 1/s then restores the
// reference's theta magnitudes 1000*beta / beta/1000, which are not scale equivariant (theta is dimensionless
// everywhere else).
inline void plan_divide(Plan& p, const double* D, const double* E, double inv_scale = 1.0) {
    const int n = p.n;
    p.D.assign(D, D + n);
    std::vector<int> order;
    order.push_back(p.root);
    for (size_t q = 0; q < order.size(); ++q) {                 // breadth first = stage by stage
        PlanNode& nd = p.nodes[order[q]];
        if (nd.left < 0) continue;
        int g = nd.off + nd.n1;                                 // first row of the right block
        nd.beta = E[g - 1];
        if (nd.mode == 1) {
            double dl = p.D[g - 1], df = p.D[g];
            if ((dl > 0 && df > 0) || (dl < 0 && df < 0)) {     // main.c:370-375
                nd.theta = ((dl * (-nd.beta)) < 0) ? -1 : 1;
            } else {                                            // main.c:376-389
                const double borig = nd.beta * inv_scale;
                if (fabs(nd.beta) < fabs(df)) nd.theta = 1000 * borig;
                else nd.theta = borig / 1000;
            }
            p.D[g - 1] -= nd.theta * nd.beta;                   // main.c:392-394
            p.D[g] -= 1.0 / nd.theta * nd.beta;
            nd.rho = nd.beta * nd.theta;                        // eigenvalues.c:54
            nd.zscale = 1.0;
        } else {
            // T = diag(T1,T2) + beta v v^T, v = e_n1 + e_(n1+1); z = Q^T v / sqrt(2), rho = 2 beta
            nd.theta = 1.0;
            p.D[g - 1] -= nd.beta;
            p.D[g] -= nd.beta;
            nd.rho = 2.0 * nd.beta;
            nd.zscale = 0.70710678118654752440;
        }
        order.push_back(nd.left);
        order.push_back(nd.right);
    }
}

// Returns 0, or 4 when n < P ("Leaf Size is too small", src/main.c:324-327).
inline int build_plan(Plan& p, int n, const double* D, const double* E, int P, int leaf_max) {
    p = Plan();
    p.n = n; p.P = P; p.leaf_max = leaf_max;
    if (P < 1 || n < 1 || n / P == 0) return 4;
    std::vector<int> loff(P), lsz(P);
    int leafSize = n / P, rem = n % P, off = 0;                 // backtransformation.c:85-96
    for (int i = 0; i < P; ++i) { lsz[i] = leafSize + (i < rem ? 1 : 0); loff[i] = off; off += lsz[i]; }
    int span = 1;
    while (span < P) span *= 2;                                 // main.c:274-281
    p.root = detail::build_reference(p, loff, lsz, 0, P, span, 0);
    detail::set_heights(p, p.root);
    plan_divide(p, D, E);
    return 0;
}

}  // namespace cuppen
#endif
