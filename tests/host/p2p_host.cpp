// TEST-ONLY: the peer-memory functors of csrc/p2p.h (CUPPEN_HD) run on the host with the ranks' symmetric heaps emulated
// as separate buffers of one process: checks the index logic of the halo-row exchange, the subtree-vector replication and
// the residual partial sums for any (ranks G, subtrees S) layout.  Built by tests/host/Makefile, driven by tests/test_p2p_host.py.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "../../symmetric_eigenvalue_b200/csrc/p2p.h"

namespace cuppen { LaunchCounter g_launches; }
using namespace cuppen;

namespace {
struct Layout {
    int n, G, S;
    std::vector<int> sub_off, sub_n;
    int slice_lo(int s, int j) const { return j >= G ? sub_n[s] : (int)(((long)j * sub_n[s] / G) & ~1L); }
    Layout(int n_, int G_, int S_) : n(n_), G(G_), S(S_) {
        for (int s = 0; s < S; ++s) { sub_off.push_back((int)((long)n * s / S)); }
        for (int s = 0; s < S; ++s) sub_n.push_back((s + 1 < S ? sub_off[s + 1] : n) - sub_off[s]);
    }
};
double val(int row, int col) { return row * 1000.0 + col + 0.25; }
}  // namespace

extern "C" int p2p_host_check(int n, int G, int S) {
    if (G > P2P_MAX || S > P2P_MAX || S < G || n < 4 * S * G) return -1;
    Layout L(n, G, S);
    // heap layout per rank: halo_lo [S][n] | halo_hi [S][n] | part [G][n] | lam [n] | frow [n] | lrow [n]
    const size_t hl = 0, hh = (size_t)S * n, pt = 2 * (size_t)S * n, lm = pt + (size_t)G * n, fr = lm + n, lr = fr + n, total = lr + n;
    std::vector<std::vector<double>> heap(G, std::vector<double>(total, -7.0));
    std::vector<int> perm(n);
    for (int c = 0; c < n; ++c) perm[c] = (int)(((long)c * 7919 + 13) % n);          // a permutation when gcd(7919, n) == 1
    { std::vector<int> seen(n, 0); for (int c : perm) seen[c]++; for (int x : seen) if (x != 1) return -2; }
    // every rank's Q in slice layout C: local row crow0[s] + i  <->  global row sub_off[s] + slice_lo(s, r) + i
    std::vector<std::vector<double>> Q(G);
    std::vector<std::vector<int>> crow0(G, std::vector<int>(S + 1, 0));
    long ldq = 0;
    for (int r = 0; r < G; ++r) {
        for (int s = 0; s < S; ++s) crow0[r][s + 1] = crow0[r][s] + L.slice_lo(s, r + 1) - L.slice_lo(s, r);
        ldq = std::max<long>(ldq, crow0[r][S]);
    }
    for (int r = 0; r < G; ++r) {
        Q[r].assign((size_t)ldq * n, 0.0);
        for (int s = 0; s < S; ++s)
            for (int i = 0; i < crow0[r][s + 1] - crow0[r][s]; ++i)
                for (int col = 0; col < n; ++col) Q[r][(size_t)col * ldq + crow0[r][s] + i] = val(L.sub_off[s] + L.slice_lo(s, r) + i, col);
    }
    int bad = 0;
    for (int r = 0; r < G; ++r) {                                // "launch" PushHaloRows on every rank
        HaloCtx h;
        h.H.me = r; h.H.G = G;
        for (int q = 0; q < G; ++q) h.H.base[q] = (char*)heap[q].data();
        h.Q = Q[r].data(); h.ldq = ldq; h.perm = perm.data(); h.n = n; h.S = S;
        h.halo_lo = heap[r].data() + hl; h.halo_hi = heap[r].data() + hh;
        for (int s = 0; s <= S; ++s) h.crow0[s] = crow0[r][s];
        PushHaloRows f{h};
        for (long t = 0; t < (long)S * n; ++t) f(t);
    }
    for (int r = 0; r < G; ++r)
        for (int s = 0; s < S; ++s) {
            const int g0 = L.sub_off[s] + L.slice_lo(s, r), g1 = L.sub_off[s] + L.slice_lo(s, r + 1);
            for (int c = 0; c < n; ++c) {
                if (g0 > 0 && heap[r][hl + (size_t)s * n + c] != val(g0 - 1, perm[c])) ++bad;       // row above my first row
                if (g1 < n && heap[r][hh + (size_t)s * n + c] != val(g1, perm[c])) ++bad;           // row below my last row
            }
        }
    // subtree vectors: rank r owns the index range of its subtrees and replicates it everywhere
    for (int r = 0; r < G; ++r) {
        const int s0 = (int)((long)S * r / G), s1 = (int)((long)S * (r + 1) / G);
        const int lo = L.sub_off[s0], hi = s1 < S ? L.sub_off[s1] : n;
        for (int g = lo; g < hi; ++g) { heap[r][lm + g] = 3.0 * g; heap[r][fr + g] = 5.0 * g; heap[r][lr + g] = 7.0 * g; }
        SymHeap H; H.me = r; H.G = G;
        for (int q = 0; q < G; ++q) H.base[q] = (char*)heap[q].data();
        PushSubtreeVectors f{H, heap[r].data() + lm, heap[r].data() + fr, heap[r].data() + lr, lo};
        for (long i = 0; i < hi - lo; ++i) f(i);
    }
    for (int r = 0; r < G; ++r)
        for (int g = 0; g < n; ++g)
            if (heap[r][lm + g] != 3.0 * g || heap[r][fr + g] != 5.0 * g || heap[r][lr + g] != 7.0 * g) ++bad;
    // residual partial sums: every rank ends with the same sum, taken in rank order
    std::vector<std::vector<double>> res(G, std::vector<double>(n));
    for (int r = 0; r < G; ++r) {
        for (int c = 0; c < n; ++c) res[r][c] = 1.0 / (1 + r) + c;
        SymHeap H; H.me = r; H.G = G;
        for (int q = 0; q < G; ++q) H.base[q] = (char*)heap[q].data();
        PushResidualPartials f{H, res[r].data(), heap[r].data() + pt, n};
        for (long c = 0; c < n; ++c) f(c);
    }
    for (int r = 0; r < G; ++r) {
        std::vector<double> out(n);
        SumResidualPartials f{heap[r].data() + pt, out.data(), n, G};
        for (long c = 0; c < n; ++c) f(c);
        for (int c = 0; c < n; ++c) {
            double want = 0;
            for (int q = 0; q < G; ++q) want += 1.0 / (1 + q) + c;
            if (out[c] != want) ++bad;
        }
    }
    return bad;
}
