// TEST-ONLY: the GEMM work list of a tree level (work_fill_problem / work_tile_at / work_tail_tiles / work_entry_* of
// csrc/matrix_stages.h, the functions the device builder build_gemm_work_body is made of) built on the host for arbitrary
// merge descriptors, so that tests/test_worklist_host.py can check the enumeration: every element of every problem is
// covered exactly once -- whole tiles in L2-sized super-columns, the tiles of a short last wave as two 64-column halves.
// Built by tests/host/Makefile.
#include <vector>

#include "../../include/cuppen_b200.h"
#include "../../symmetric_eigenvalue_b200/csrc/host_twins.h"

namespace cuppen { LaunchCounter g_launches; }
using namespace cuppen;

// merges [nd] of sizes m[i] = n1[i] + n2[i] (stored one after the other), k[i] live roots, ktop / kbot live columns per half;
// one rank holding all rows.  out_tiles: 4 ints per entry (problem, m0, n0, width); out_probs: 3 ints per problem (M, N, K).
// Returns the number of entries, or -1 when the list does not fit `cap`.
extern "C" int worklist_build(int nd, const int* n1, const int* n2, const int* k, const int* ktop, const int* kbot, int p0, int width,
                              int BMN, int split_grid, int supercol_mb, int cap, int* out_tiles, int* out_probs) {
    std::vector<MergeDesc> desc(nd);
    int off = 0;
    for (int i = 0; i < nd; ++i) {
        MergeDesc& D = desc[i];
        memset(&D, 0, sizeof D);
        D.off = off; D.n1 = n1[i]; D.n2 = n2[i]; D.m = n1[i] + n2[i];
        D.lr0 = off; D.lsplit = off + n1[i]; D.lr1 = off + D.m;
        D.k = k[i]; D.ktop = ktop[i]; D.kbot = kbot[i];
        off += D.m;
    }
    std::vector<GemmProblem> probs(2 * nd);
    std::vector<GemmTile> tiles(cap);
    std::vector<int> lidx(off + 64, 0);
    int ntiles[2] = {0, 0}, fail = 0;
    WorkCtx w;
    memset(&w, 0, sizeof w);
    w.desc = desc.data(); w.nd = nd; w.p0 = p0; w.width = width; w.BM = BMN; w.BN = BMN;
    w.ldq = off + 16; w.ldb = width + 16; w.Apack = nullptr; w.B = nullptr; w.Qnext = nullptr; w.lidx = lidx.data();
    w.probs = probs.data(); w.tiles = tiles.data(); w.ntiles = ntiles; w.tile_cap = cap; w.fail = &fail;
    w.supercol_mb = supercol_mb; w.split_grid = split_grid;
    build_gemm_work_host(w);
    if (fail) return -1;
    for (int t = 0; t < ntiles[0]; ++t) {
        out_tiles[4 * t] = tiles[t].prob & GEMM_TILE_PROB_MASK;
        out_tiles[4 * t + 1] = tiles[t].m0;
        out_tiles[4 * t + 2] = tiles[t].n0;
        out_tiles[4 * t + 3] = (tiles[t].prob & GEMM_TILE_HALF) ? 64 : BMN;
    }
    for (int p = 0; p < 2 * nd; ++p) { out_probs[3 * p] = probs[p].M; out_probs[3 * p + 1] = probs[p].N; out_probs[3 * p + 2] = probs[p].K; }
    return ntiles[0];
}
