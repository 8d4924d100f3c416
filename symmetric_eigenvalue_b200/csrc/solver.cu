// libcuppen_b200: host orchestration of the divide-and-conquer tree and the C ABI
// (include/cuppen_b200.h).  Replaces the conquer loop of the reference's main()
// (/root/reference/src/main.c:495-664) and the back-transformation of writeResults
// (/root/reference/src/filehandling.c:332-548): one process drives one B200 (several processes / GPUs
// cooperate through peer memory, p2p.h); all merges of a tree level are batched into the same launches.
#include <math.h>
#include <algorithm>
#include <chrono>
#include <map>

#include "../../include/cuppen_b200.h"
#include "platform.h"
#include "plan.h"
#include "merge_stages.h"
#include "matrix_stages.h"
#include "select_stages.h"
#include "orth_check.h"
#include "gemm_dmma.h"
#include "gemm_tma.h"
#include "host_twins.h"
#include "comm.h"
#include "p2p.h"
#include "dense_stages.h"

namespace cuppen {

LaunchCounter g_launches;
static thread_local std::string g_last_error;

static double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- per-category device timers (CUDA events on the solver's stream) ---------------------------
// fail buffer of a solve (device, read back with the results): [0] leaf QL did not converge (1 + row),
// [1] GEMM tile list overflow, [2] peer barrier timed out (1 + peer rank), [4..11] watchdog record of the TMA GEMM pipeline
enum { FAIL_LEAF = 0, FAIL_TILES = 1, FAIL_COMM = 2, FAIL_TMA = 4, FAIL_INTS = 16 };
enum { GEMM_HINT_MIN_ROWS = 8192 };      // merges from this size on run the tensor-map GEMM with its L2 eviction hints (gemm_tma.h)
enum { T_LEAF, T_DEFL, T_ROOT, T_EVX, T_PACK, T_UGEN, T_GEMM, T_RESID, T_APPLY, T_COMM, T_NCAT };
struct PhaseTimers {
    double acc[T_NCAT] = {0};
    bool capturing = false;     // stream capture in progress: record events as external graph nodes
    bool keep = false;          // spans belong to an instantiated graph: re-read them after every replay
#if CUPPEN_CUDA
    struct Span { int cat; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); return e;
    }
    void record(cudaEvent_t e, Stream s) {
        if (capturing) CUDA_CHECK(cudaEventRecordWithFlags(e, s, cudaEventRecordExternal));
        else CUDA_CHECK(cudaEventRecord(e, s));
    }
    void begin(int cat, Stream s) { Span sp{cat, get(), get()}; record(sp.a, s); spans.push_back(sp); }
    void end(Stream s) { record(spans.back().b, s); }
    void collect() {
        for (auto& sp : spans) {
            float ms = 0; cudaEventElapsedTime(&ms, sp.a, sp.b);
            acc[sp.cat] += ms * 1e-3;
            if (!keep) { pool.push_back(sp.a); pool.push_back(sp.b); }
        }
        if (!keep) spans.clear();
    }
    void drop_spans() { for (auto& sp : spans) { pool.push_back(sp.a); pool.push_back(sp.b); } spans.clear(); keep = false; }
    ~PhaseTimers() { for (auto e : pool) cudaEventDestroy(e); for (auto& sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);} }
#else
    int cur = -1; double t0 = 0;
    void begin(int cat, Stream) { cur = cat; t0 = wall_now(); }
    void end(Stream) { acc[cur] += wall_now() - t0; }
    void collect() {}
    void drop_spans() {}
#endif
    void reset() { for (double& a : acc) a = 0; }
};

struct Solver {
    int n = 0, P = 1, flags = 0, device = 0;
    bool want_vectors = false;
    bool select_mode = false;     // CUPPEN_FLAG_SELECT: eigenvalue-only solve + back-application of the selected columns
    Comm comm;
    Plan plan;
    bool have_matrix = false, solved = false;
    long ldq = 0, ldb = 0;
    int W = 0;                    // panel width of the U arena
    int leaf_max = 16;            // leaf size of the accurate tree (measured best of 8/16/32 at n=4096); env CUPPEN_LEAF
    Stream stream = 0;

    // ---- row ownership ---------------------------------------------------------------------------
    // G ranks (any count), S subtrees = the nodes at depth ceil(log2 G) of the divide tree, global rows
    // [sub_off[s], sub_off[s]+sub_n[s]); rank g owns the contiguous subtrees [first_sub(g), first_sub(g+1)).
    //  layout L (local phase): rank g holds all rows of its subtrees, local row = r - R0.
    //  layout C (cooperative phase, the levels above the subtrees): rank g holds slice g of EVERY subtree,
    //    rows [sub_off[s]+slice_lo(s,g), sub_off[s]+slice_lo(s,g+1)) at local rows crow0[s]... ;
    //    every cooperative merge is then split evenly over all ranks whatever its halves deflate to.
    int G = 1, glog = 0, S = 1;
    int first_sub(int r) const { return (int)((long)S * r / G); }
    int owner_of_sub(int sub) const { int r = 0; while (r + 1 < G && first_sub(r + 1) <= sub) ++r; return r; }
    int rank_row0(int r) const { return r >= G ? n : sub_off[first_sub(r)]; }
    std::vector<int> sub_off, sub_n, sub_root;
    int R0 = 0, R1 = 0, nlocL = 0, nlocC = 0;
    std::vector<int> crow0;
    int nloc_final = 0;           // local rows of the final V
    int slice_lo(int s, int j) const { return j >= G ? sub_n[s] : (int)(((long)j * sub_n[s] / G) & ~1L); }
    long rows_of(int r) const {   // rows rank r needs in either layout (every rank can work out every rank's layout)
        long c = 0;
        for (int s = 0; s < S; ++s) c += slice_lo(s, r + 1) - slice_lo(s, r);
        return std::max<long>(rank_row0(r + 1) - rank_row0(r), c);
    }

    std::vector<double> hD, hE;   // matrix as uploaded (host): the caller's T times `scale`
    // Matrices of extreme norm (max|T_ij| outside [1e-100, 1e100]) are scaled by a power of two before the divide
    // pass, as dstedc does with dlascl (products of z^2, rho and pole gaps would leave the fp64 range otherwise);
    // eigenvalues and residuals are scaled back on the way out.  Everything in between is exact under a power-of-
    // two scaling, the reference rule's absolute 1e-5 gap threshold follows the scaling (MergeDesc::dthr).
    double scale = 1.0;
    DevBuf<double> dDm, dE, dOD, dOE;
    DevBuf<double> lam, lam_sorted, frow, lrow, frow2, lrow2, fpack, lpack;
    DevBuf<double> d, z, dn, zn, gc, gs, dl, wl, zl, tau, dorgv, zhat, nrm, res2, halo, halo_all;
    DevBuf<int> G_, lsort, head, sup, prev, tpos, bpos, lidx, org, toplist, botlist, perm, fail;
    // selected-eigenvector mode: the per-level scratch vectors above get one slice per level (stride
    // lvl_stride) so that every level's U stays defined after the solve
    size_t lvl_stride = 0;
    int lvl_cap = 1;
    std::vector<int> h_sel;                  // requested ranks (ascending-lambda order), caller's order
    // several ranks: the eigenvalue-only decomposition is replicated (61 ms at n = 65536 -- cheaper than exchanging its
    // per-level vectors), the selected vectors are dealt round-robin to the ranks (vector t goes to rank t % world) and
    // gathered at the end, so every rank returns all of them (src/filehandling.c:339-348 serves -eFILE at any rank count)
    Comm sel_comm;
    int sel_rank = 0, sel_world = 1;
    std::vector<int> h_sel_local;
    DevBuf<double> Vgath, res_gath;
    int sel_per() const { return ((int)h_sel.size() + sel_world - 1) / sel_world; }
    DevBuf<int> sel_dev, leaf_off_dev, leaf_n_dev;
    DevBuf<RowSpan> colspan, colspan_sub;    // row support per storage column (matrix_stages.h); the subtree blocks (several ranks: the spans at the transition)
    DevBuf<double> Qleaf, selX, selY, selXS, selGam, sel_dorg, Vsel, res_sel;
    std::vector<double> h_res_sel;
    void enqueue_apply();
#if CUPPEN_CUDA
    cudaEvent_t ev_ap0 = nullptr, ev_ap1 = nullptr;
#endif
    DevBuf<double> Qa, Apack, B;      // the merges work in place on Qa: Apack holds the packed live columns of a level
    DevBuf<LeafDesc> leaves;
    DevBuf<GemmProblem> probs;
    DevBuf<GemmTile> tiles;
    double* Qcur = nullptr;       // block-diagonal eigenvector matrix: children before a level, parents after it, V at the end
    double* Awork = nullptr;      // the other of the two n x n buffers: packed live columns of the level in flight.  Qcur / Awork
                                  // start every solve as Qa / Apack and trade places at the peer-memory row redistribution
    double* Qfinal = nullptr;     // where a (replayed) solve leaves V and its pack buffer
    double* Afinal = nullptr;
    // ---- peer-memory back end (p2p.h): symmetric heap + IPC mappings of the peers' heaps and Q buffers -------------
    struct P2PState {
        bool on = false;
        SymHeap H;
        char* heap = nullptr;
        size_t heap_bytes = 0, used = 0;
        const double* peerQ[P2P_MAX] = {nullptr};
        void* opened[2 * P2P_MAX] = {nullptr};
        int nopened = 0;
        unsigned* flags = nullptr;     // heap
        unsigned* epoch = nullptr;     // heap (local use)
        double *halo_lo = nullptr, *halo_hi = nullptr, *res_part = nullptr;
        template <class T>
        T* carve(size_t count) {
            used = (used + 255) & ~(size_t)255;
            T* p = (T*)(heap + used);
            used += count * sizeof(T);
            if (used > heap_bytes) CUPPEN_THROW(CUPPEN_ERR_STATE, "symmetric heap overflow (%zu of %zu bytes)", used, heap_bytes);
            return p;
        }
    } p2p;
    void setup_p2p();
    void p2p_barrier();
    bool sorted_materialised = false;    // Qcur (= Apack's storage) holds the columns in ascending-lambda order
    void materialise_sorted();
    // one-GPU solves are captured into a CUDA graph on the second call and replayed afterwards
    // (the whole decomposition is enqueued without any host read-back); env CUPPEN_GRAPH=0 disables
    bool use_graph = true, graph_failed = false;
    int solves_done = 0;
    long graph_launches = 0;
    std::vector<LeafDesc> h_leaves;
    double* pin_lam = nullptr;           // pinned staging of the results read back at the end of a solve
    double* pin_res = nullptr;
    MergeDesc* pin_desc = nullptr;
    int* pin_fail = nullptr;
    size_t pin_desc_cap = 0;
#if CUPPEN_CUDA
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
#endif
    void enqueue_solve();
    void drop_graph();
    ~Solver();

    std::vector<double> h_lam_sorted, h_resid;
    std::vector<cuppen_merge_stat> stats;
    PhaseTimers pt;
    cuppen_timers timers;
    // graph replays: the ~60 per-phase event pairs are only read when somebody asks for the timers
    bool phase_timers_pending = false;
    double bt_wall_extra = 0;
    void fill_phase_timers() {
        timers.root_finding_s = pt.acc[T_ROOT];
        timers.ev_extract_s = pt.acc[T_EVX] + pt.acc[T_UGEN];
        timers.backtransform_s = pt.acc[T_PACK] + pt.acc[T_GEMM] + pt.acc[T_UGEN] + bt_wall_extra;
        timers.backtransform_ev_s = pt.acc[T_UGEN];
        timers.gemm_s = pt.acc[T_GEMM];
        timers.leaf_s = pt.acc[T_LEAF];
        timers.deflation_s = pt.acc[T_DEFL];
        timers.pack_s = pt.acc[T_PACK];
        timers.residual_s = pt.acc[T_RESID];
        timers.comm_s = pt.acc[T_COMM];
        timers.comm_mode = G <= 1 ? 0 : (p2p.on ? 2 : 1);
        if (select_mode && !h_sel.empty()) timers.backtransform_s = timers.apply_s;
    }
    double acc_pack_bytes = 0, acc_ugen_bytes = 0, acc_gemm_flop = 0;
    int resid_variant = 0;        // residual_kernel variant (0: default); env CUPPEN_RESID
    int gemm_hints = 0;           // L2 eviction hints of the tensor-map GEMM on merges of >= GEMM_HINT_MIN_ROWS rows (1: B evict_last + streaming C stores, 3: + A evict_first); env CUPPEN_GEMM_HINT
    bool split_tail = true;       // half tiles for an under-filled last GEMM wave (build_gemm_work_body); env CUPPEN_SPLIT_TAIL=0 switches it off
    bool track_spans = true;      // row support per column (RowSpan, matrix_stages.h); env CUPPEN_SPAN=0 switches it off
    int supercol_mb = 48;         // L2 budget of a super-column's B panel (work_supercol); env CUPPEN_SUPERCOL_MB
    int gemm_variant = 2;         // 0: cp.async kernel (gemm_dmma.h), 1: TMA bulk-copy lines, 2: TMA tensor maps (gemm_tma.h, default); env CUPPEN_GEMM
#if CUPPEN_CUDA
    CUtensorMap map_qa, map_apack, map_b;      // tensor maps of the two n x n buffers (either can be the pack buffer) and of the U arena
#endif
    int num_sms = 148;
#if CUPPEN_CUDA
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
#endif
    std::vector<int> parent_of;              // plan node -> parent node id

    // ---- schedule: local levels (by height), then cooperative levels (by height) ------------------
    struct LevelInfo {
        std::vector<int> ids;                // plan nodes merged at this step
        size_t desc_off = 0;                 // offset into desc_all
        int height = 0;
        bool coop = false;
        int maxm = 0, maxm_rows = 0;
        bool any_accurate = false;
        long worst_tiles_big = 0, worst_tiles_small = 0;
    };
    std::vector<LevelInfo> levels;
    int first_coop = -1;                     // index of the first cooperative level (-1: none)
    std::vector<MergeDesc> h_desc_all;
    DevBuf<MergeDesc> desc_all;
    DevBuf<int> node_of_all;                 // [levels][n]
    DevBuf<int> ntiles_dev;

    void init_layout();
    void allocate();
    void prepare_levels();          // tree-shape dependent data: once per handle (the shape depends on (n, P) only)
    void update_descriptors();      // matrix dependent fields (rho, theta, sigma): every cuppen_set_tridiagonal
    void set_matrix(const double* D, const double* E);
    void solve();
    void run_leaves();
    void run_level(int li);
    void enter_cooperative();
    void enter_cooperative_p2p();
    void finish();
    LevelCtx level_ctx(int li);
    MatCtx mat_ctx();
    int subtree_at(int off) const {
        for (int s = 0; s < S; ++s) if (sub_off[s] == off) return s;
        if (off == n) return S;
        CUPPEN_THROW(CUPPEN_ERR_STATE, "offset %d is not a subtree boundary", off);
    }
};

// ---- layout --------------------------------------------------------------------------------------
void Solver::init_layout() {
    G = comm.world;
    glog = 0;
    while ((1 << glog) < G) ++glog;
    std::vector<std::pair<int, int>> roots;      // (offset, node id) of the nodes at depth log2 G
    for (size_t id = 0; id < plan.nodes.size(); ++id)
        if (plan.nodes[id].depth == glog) roots.push_back({plan.nodes[id].off, (int)id});
    std::sort(roots.begin(), roots.end());
    long covered = 0;
    for (auto& r : roots) covered += plan.nodes[r.second].n;
    S = (int)roots.size();
    if (S < G || covered != n || S > P2P_MAX * 2)
        CUPPEN_THROW(CUPPEN_ERR_ARG, "n=%d is too small to distribute over %d GPUs: the divide tree (reference leaves %d) has %d "
                     "subtrees at depth %d", n, G, P, S, glog);
    sub_off.resize(S); sub_n.resize(S); sub_root.resize(S);
    for (int s = 0; s < S; ++s) { sub_off[s] = roots[s].first; sub_root[s] = roots[s].second; sub_n[s] = plan.nodes[roots[s].second].n; }
    parent_of.assign(plan.nodes.size(), -1);
    for (size_t id = 0; id < plan.nodes.size(); ++id)
        if (plan.nodes[id].left >= 0) { parent_of[plan.nodes[id].left] = (int)id; parent_of[plan.nodes[id].right] = (int)id; }
    R0 = rank_row0(comm.rank);
    R1 = rank_row0(comm.rank + 1);
    nlocL = R1 - R0;
    crow0.assign(S + 1, 0);
    for (int s = 0; s < S; ++s) {
        const int len = slice_lo(s, comm.rank + 1) - slice_lo(s, comm.rank);
        if (G > 1 && want_vectors && len <= 0)
            CUPPEN_THROW(CUPPEN_ERR_ARG, "n=%d is too small to distribute over %d GPUs", n, G);
        crow0[s + 1] = crow0[s] + len;
    }
    nlocC = crow0[S];
    nloc_final = (G > 1) ? nlocC : nlocL;
    // one leading dimension for all ranks (the largest row count): peers address each other's buffers with it
    long rows = 0;
    for (int r = 0; r < G; ++r) rows = std::max(rows, rows_of(r));
    if (rows < std::max(nlocL, nlocC)) CUPPEN_THROW(CUPPEN_ERR_STATE, "inconsistent row layout");
    ldq = round_up(rows, 16);
}

// Function attributes are per device: every handle opts its device in to the >48 KB dynamic shared memory of the
// GEMM, secular and Gram kernels (a process-wide "done" flag would leave a second device without it).
static void set_kernel_attributes() {
#if CUPPEN_CUDA
    CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<64, 64, 16, 2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)DmmaCfg<64, 64, 16, 2, 2, 3>::SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<128, 128, 16, 2, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)DmmaCfg<128, 128, 16, 2, 4, 3>::SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes()));
    CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes()));
    CUDA_CHECK(cudaFuncSetAttribute(secular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * SEC_SMEM_K * sizeof(double))));
    CUDA_CHECK(cudaFuncSetAttribute(fused_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_front_smem_bytes(FUSE_MAXM)));
    CUDA_CHECK(cudaFuncSetAttribute(gram_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_smem_bytes()));
#endif
}

void Solver::allocate() {
    const size_t N = (size_t)n;
    set_kernel_attributes();
    // several GPUs over NCCL: the exchanged vectors live in a symmetric heap that the peers map (p2p.h); env
    // CUPPEN_P2P=0 keeps them private and exchanges them with NCCL collectives instead
    bool want_p2p = false;
#if CUPPEN_CUDA
    {
        const char* pe = getenv("CUPPEN_P2P");
        want_p2p = G > 1 && G <= P2P_MAX && S <= P2P_MAX && comm.nccl != nullptr && !(pe && !strcmp(pe, "0"));
    }
    if (want_p2p) {
        p2p.heap_bytes = (size_t)(8 + 2 * S + G) * (N + 64) * sizeof(double) + 64 * 256;
        CUDA_CHECK(cudaMalloc((void**)&p2p.heap, p2p.heap_bytes));
        CUDA_CHECK(cudaMemsetAsync(p2p.heap, 0, p2p.heap_bytes, stream));
        lam.attach(p2p.carve<double>(N + 64), N + 64);
        frow.attach(p2p.carve<double>(N + 64), N + 64);
        lrow.attach(p2p.carve<double>(N + 64), N + 64);
        tau.attach(p2p.carve<double>(N + 64), N + 64);
        dorgv.attach(p2p.carve<double>(N + 64), N + 64);
        org.attach(p2p.carve<int>(N + 64), N + 64);
        zhat.attach(p2p.carve<double>(N + 64), N + 64);
        nrm.attach(p2p.carve<double>(N + 64), N + 64);
        p2p.halo_lo = p2p.carve<double>((size_t)S * N);
        p2p.halo_hi = p2p.carve<double>((size_t)S * N);
        p2p.res_part = p2p.carve<double>((size_t)G * N);
        p2p.flags = p2p.carve<unsigned>(64);
        p2p.epoch = p2p.carve<unsigned>(64);
    }
#endif
    for (DevBuf<double>* b : {&dDm, &dE, &dOD, &dOE, &lam, &lam_sorted, &frow, &lrow, &frow2, &lrow2, &fpack, &lpack, &res2})
        if (b->p == nullptr) b->alloc(N + 64);
    perm.alloc(N + 64);
    colspan.alloc(N + 64);
    {
        // (defined before the first leaf kernel has run; several ranks: a rank tracks the columns of its own subtrees only,
        // so the cooperative phase starts from the subtree blocks on every rank)
        std::vector<RowSpan> full(N + 64, RowSpan{0, n});
        dev_h2d(colspan.p, full.data(), sizeof(RowSpan) * full.size(), stream);
        if (G > 1) {
            for (int t = 0; t < S; ++t)
                for (int g = sub_off[t]; g < sub_off[t] + sub_n[t]; ++g) full[g] = RowSpan{sub_off[t], sub_off[t] + sub_n[t]};
            colspan_sub.alloc(N + 64);
            dev_h2d(colspan_sub.p, full.data(), sizeof(RowSpan) * full.size(), stream);
        }
        dev_sync(stream);
    }
    // scratch vectors of a level: shared by all levels, or one slice per level in selected-eigenvector mode
    lvl_stride = N + 64;
    lvl_cap = select_mode ? (int)plan.by_height.size() + 1 : 1;
    for (DevBuf<double>* b : {&d, &z, &dn, &zn, &gc, &gs, &dl, &wl, &zl, &tau, &dorgv, &zhat, &nrm})
        if (b->p == nullptr) b->alloc(lvl_stride * lvl_cap);
    for (DevBuf<int>* b : {&G_, &lsort, &head, &sup, &prev, &tpos, &bpos, &lidx, &org, &toplist, &botlist})
        if (b->p == nullptr) b->alloc(lvl_stride * lvl_cap);
    // (ugen_kernel reads both K-list entries of a row before it knows which one is defined: never uninitialised memory)
    dev_zero(toplist.p, toplist.bytes(), stream);
    dev_zero(botlist.p, botlist.bytes(), stream);
    if (select_mode) {
        Qleaf.alloc(N * LEAF_MAX + 64);
        leaf_off_dev.alloc(N + 64);
        leaf_n_dev.alloc(N + 64);
        for (DevBuf<double>* b : {&selX, &selY, &selXS, &selGam}) b->alloc(N * SEL_NV + 64);
        sel_dorg.alloc(N + 64);
    }
    fail.alloc(FAIL_INTS);
    if (want_p2p) { halo.alloc(64); halo_all.alloc(64); }          // the halo rows go straight into the peers' heaps
    else {
        halo.alloc(2 * N * (size_t)std::max(1, S) + 64);
        halo_all.alloc(2 * N * (size_t)S * (size_t)G + 64);
    }
    leaves.alloc(std::max<size_t>(1, plan.leaves.size()));
    if (want_vectors) {
        const size_t qelems = (size_t)ldq * (N + K_PAD + 2) + 4096;
        Qa.alloc(qelems);
        Apack.alloc(qelems);
        // U arena: rows indexed by global pole index, W columns per panel (<= 2 GiB)
        const size_t cap = (size_t)1 << 28;
        W = (int)std::min<size_t>(N, std::max<size_t>(256, cap / N));
        ldb = round_up(W, 16) + 16;
        B.alloc((N + 2 * K_PAD + 8) * (size_t)ldb + 4096);
        dev_zero(B.p, B.bytes(), stream);
        dev_zero(Apack.p, Apack.bytes(), stream);
        dev_zero(Qa.p, Qa.bytes(), stream);
#if CUPPEN_CUDA
        const char* rv = getenv("CUPPEN_RESID");
        if (rv && atoi(rv) > 0) resid_variant = atoi(rv);
        const char* gv = getenv("CUPPEN_GEMM");
        if (gv && (!strcmp(gv, "cpasync") || !strcmp(gv, "v1"))) gemm_variant = 0;
        if (gv && !strcmp(gv, "bulk")) gemm_variant = 1;
        const char* stv = getenv("CUPPEN_SPLIT_TAIL");
        if (stv && atoi(stv) == 0) split_tail = false;
        const char* tv = getenv("CUPPEN_SPAN");
        if (tv && atoi(tv) == 0) track_spans = false;
        const char* hv = getenv("CUPPEN_GEMM_HINT");
        if (hv) gemm_hints = atoi(hv) & 3;
        const char* sv = getenv("CUPPEN_SUPERCOL_MB");
        if (sv && atoi(sv) >= 1 && atoi(sv) <= 120) supercol_mb = atoi(sv);
        if (gemm_variant == 2) {
            // tensor maps of the operand buffers (35.9 vs 35.5 TFLOP/s with bulk-copy lines at 8192^3, profiles/r02_gemm_bench.txt)
            const long cols = (long)(qelems / (size_t)ldq), brows = (long)(B.n / (size_t)ldb);
            if (!tma_encode_map(&map_qa, Qa.p, ldq, cols, ldq) || !tma_encode_map(&map_apack, Apack.p, ldq, cols, ldq) ||
                !tma_encode_map(&map_b, B.p, ldb, brows, ldb)) {
                if (gv && !strcmp(gv, "tensor")) CUPPEN_THROW(CUPPEN_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUPPEN_GEMM=tensor)");
                gemm_variant = 1;                   // no tensor maps from this driver: bulk-copy lines
            }
        }
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        num_sms = prop.multiProcessorCount;
#endif
    }
    dev_zero(halo.p, halo.bytes(), stream);
#if CUPPEN_CUDA
    CUDA_CHECK(cudaMallocHost((void**)&pin_lam, sizeof(double) * N));
    CUDA_CHECK(cudaMallocHost((void**)&pin_res, sizeof(double) * N));
    CUDA_CHECK(cudaMallocHost((void**)&pin_fail, sizeof(int) * FAIL_INTS));
    const char* ge = getenv("CUPPEN_GRAPH");
    if (ge && !strcmp(ge, "0")) use_graph = false;
#else
    pin_lam = (double*)malloc(sizeof(double) * N);
    pin_res = (double*)malloc(sizeof(double) * N);
    pin_fail = (int*)malloc(sizeof(int) * FAIL_INTS);
    use_graph = false;
#endif
    dev_sync(stream);
    if (want_p2p) setup_p2p();
    if (G > 1 && !p2p.on) use_graph = false;         // NCCL collectives / callbacks between the kernels: launched eagerly
}

// Exchange the CUDA IPC handles of the symmetric heap and of the Q buffer (NCCL all-gather: the bootstrap) and map the
// peers' copies.  All ranks agree on the outcome (a rank that cannot map a peer makes everybody fall back to the
// collective back end).
void Solver::setup_p2p() {
#if CUPPEN_CUDA
    struct Handles { cudaIpcMemHandle_t heap, q; int has_q; int pad[15]; };
    static_assert(sizeof(Handles) == 2 * sizeof(cudaIpcMemHandle_t) + 64, "handle record");
    Handles mine;
    memset(&mine, 0, sizeof mine);
    int ok = 1;
    if (cudaIpcGetMemHandle(&mine.heap, p2p.heap) != cudaSuccess) ok = 0;
    mine.has_q = want_vectors ? 1 : 0;
    if (ok && want_vectors && cudaIpcGetMemHandle(&mine.q, Qa.p) != cudaSuccess) ok = 0;
    cudaGetLastError();
    DevBuf<unsigned char> sbuf, rbuf;
    sbuf.alloc(sizeof mine); rbuf.alloc(sizeof mine * G);
    dev_h2d(sbuf.p, &mine, sizeof mine, stream);
    comm.allgather(sbuf.p, rbuf.p, sizeof mine, stream);
    std::vector<Handles> all(G);
    dev_d2h(all.data(), rbuf.p, sizeof mine * G, stream);
    dev_sync(stream);
    p2p.H.me = comm.rank; p2p.H.G = G;
    for (int r = 0; r < P2P_MAX; ++r) p2p.H.base[r] = nullptr;
    p2p.H.base[comm.rank] = p2p.heap;
    p2p.peerQ[comm.rank] = Qa.p;
    for (int r = 0; r < G && ok; ++r) {
        if (r == comm.rank) continue;
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r].heap, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        p2p.opened[p2p.nopened++] = ptr;
        p2p.H.base[r] = (char*)ptr;
        if (want_vectors) {
            if (cudaIpcOpenMemHandle(&ptr, all[r].q, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            p2p.opened[p2p.nopened++] = ptr;
            p2p.peerQ[r] = (const double*)ptr;
        }
    }
    cudaGetLastError();
    // agreement: sum of the failure counts over the ranks
    DevBuf<double> agree;
    agree.alloc(8);
    double bad = ok ? 0.0 : 1.0;
    dev_h2d(agree.p, &bad, sizeof bad, stream);
    comm.allreduce_sum(agree.p, 1, stream);
    dev_d2h(&bad, agree.p, sizeof bad, stream);
    dev_sync(stream);
    p2p.on = (bad == 0.0);
    if (!p2p.on) {
        const char* strict = getenv("CUPPEN_P2P");
        if (strict && !strcmp(strict, "require"))
            CUPPEN_THROW(CUPPEN_ERR_COMM, "peer memory (CUDA IPC) is not available between the %d GPUs", G);
        p2p.H.G = 0;
        // the collective back end needs its staging buffers after all
        halo.alloc(2 * (size_t)n * (size_t)S + 64);
        halo_all.alloc(2 * (size_t)n * (size_t)S * (size_t)G + 64);
        dev_zero(halo.p, halo.bytes(), stream);
        dev_sync(stream);
    }
#endif
}

void Solver::p2p_barrier() {
#if CUPPEN_CUDA
    p2p_barrier_kernel<<<1, 32, 0, stream>>>(p2p.H, p2p.flags, p2p.epoch, fail.p + FAIL_COMM);
    CUDA_CHECK(cudaGetLastError());
    g_launches.launches++;
#endif
}

Solver::~Solver() {
    drop_graph();
#if CUPPEN_CUDA
    for (int i = 0; i < p2p.nopened; ++i) cudaIpcCloseMemHandle(p2p.opened[i]);
    if (p2p.heap) cudaFree(p2p.heap);
    if (pin_lam) cudaFreeHost(pin_lam);
    if (pin_res) cudaFreeHost(pin_res);
    if (pin_fail) cudaFreeHost(pin_fail);
    if (pin_desc) cudaFreeHost(pin_desc);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (ev_ap0) cudaEventDestroy(ev_ap0);
    if (ev_ap1) cudaEventDestroy(ev_ap1);
#else
    free(pin_lam); free(pin_res); free(pin_fail); free(pin_desc);
#endif
}

void Solver::drop_graph() {
#if CUPPEN_CUDA
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
    if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
#endif
    pt.drop_spans();
    solves_done = 0;
}

void Solver::set_matrix(const double* D, const double* E) {
    // a failed call must not leave the previous matrix's decomposition looking valid
    have_matrix = false;
    solved = false;
    double amax = 0;
    for (int i = 0; i < n; ++i) {
        if (!std::isfinite(D[i])) CUPPEN_THROW(CUPPEN_ERR_ARG, "D[%d] is not finite", i);
        amax = std::max(amax, fabs(D[i]));
    }
    for (int i = 0; i + 1 < n; ++i) {
        if (!std::isfinite(E[i])) CUPPEN_THROW(CUPPEN_ERR_ARG, "E[%d] is not finite", i);
        amax = std::max(amax, fabs(E[i]));
    }
    scale = 1.0;
    if (amax > 0 && (amax < 1e-100 || amax > 1e100)) scale = ldexp(1.0, -ilogb(amax));
    hD.assign(D, D + n);
    hE.assign(E, E + std::max(0, n - 1));
    if (scale != 1.0) {
        for (double& v : hD) v *= scale;
        for (double& v : hE) v *= scale;
    }
    // the tree was built in cuppen_create (its shape depends on (n, P, leaf size) only): a new matrix needs the
    // divide pass alone
    if (plan.n != n || plan.P != P || plan.nodes.empty()) CUPPEN_THROW(CUPPEN_ERR_STATE, "handle without a divide tree");
    plan_divide(plan, hD.data(), hE.data(), 1.0 / scale);
    // only the reference-rule splits need beta != 0 (assert at src/main.c:196-200, src/eigenvalues.c:68)
    for (const PlanNode& nd : plan.nodes)
        if (nd.left >= 0 && nd.mode == MODE_REFERENCE && nd.rho == 0.0)
            CUPPEN_THROW(CUPPEN_ERR_ZERO, "zero off-diagonal entry at a reference split (row %d)", nd.off + nd.n1);
    update_descriptors();
    dev_h2d(dDm.p, plan.D.data(), sizeof(double) * n, stream);
    dev_h2d(dOD.p, hD.data(), sizeof(double) * n, stream);
    if (n > 1) {
        dev_h2d(dE.p, hE.data(), sizeof(double) * (n - 1), stream);
        dev_h2d(dOE.p, hE.data(), sizeof(double) * (n - 1), stream);
    }
    dev_sync(stream);
    have_matrix = true;
}

LevelCtx Solver::level_ctx(int li) {
    LevelCtx c;
    if (select_mode && li >= lvl_cap) CUPPEN_THROW(CUPPEN_ERR_STATE, "level %d beyond the %d per-level slices", li, lvl_cap);
    const size_t o = select_mode ? (size_t)li * lvl_stride : 0;
    c.n = n; c.desc = desc_all.p + levels[li].desc_off; c.node_of = node_of_all.p + (size_t)li * n; c.lam = lam.p; c.frow = frow.p; c.lrow = lrow.p;
    // one GPU with eigenvectors: z is read from the rows of Q directly (no ExtractRows pass between the levels)
    c.Qz = (want_vectors && G == 1) ? Qcur : nullptr;
    c.ldqz = ldq;
    c.d = d.p + o; c.z = z.p + o; c.dn = dn.p + o; c.zn = zn.p + o; c.G = G_.p + o; c.gc = gc.p + o; c.gs = gs.p + o; c.lsort = lsort.p + o;
    c.head = head.p + o; c.sup = sup.p + o; c.prev = prev.p + o; c.tpos = tpos.p + o; c.bpos = bpos.p + o; c.dl = dl.p + o; c.wl = wl.p + o; c.zl = zl.p + o;
    c.lidx = lidx.p + o; c.org = org.p + o; c.tau = tau.p + o; c.dorgv = dorgv.p + o; c.zhat = zhat.p + o; c.nrm = nrm.p + o; c.toplist = toplist.p + o;
    c.botlist = botlist.p + o;
    return c;
}

MatCtx Solver::mat_ctx() {
    MatCtx M;
    M.n = n; M.ldq = ldq; M.Qold = Qcur; M.Qnew = Qcur; M.Apack = Awork; M.B = B.p; M.ldb = ldb;       // in place
    M.span = (want_vectors && track_spans) ? colspan.p : nullptr;
    return M;
}

// ---- leaves ---------------------------------------------------------------------------------------
void Solver::run_leaves() {
    const std::vector<LeafDesc>& hl = h_leaves;
    dev_zero(fail.p, sizeof(int) * FAIL_INTS, stream);
    if (hl.empty()) return;
    pt.begin(T_LEAF, stream);
    // selected-eigenvector mode keeps the leaf eigenvectors compactly: (row, column c of its leaf) at row + c*n
    double* Q = want_vectors ? Qcur : (select_mode ? Qleaf.p : nullptr);
    const long ldleaf = select_mode ? (long)n : ldq;
    const int compact = select_mode ? 1 : 0;
#if CUPPEN_CUDA
    leaf_ql_kernel<<<(unsigned)((hl.size() + 3) / 4), 128, 0, stream>>>(leaves.p, (int)hl.size(), dDm.p, dE.p, lam.p, frow.p,
                                                                       lrow.p, Q, ldleaf, R0, fail.p, compact, (want_vectors && track_spans) ? colspan.p : nullptr);
    CUDA_CHECK(cudaGetLastError());
#else
    leaf_ql_host(leaves.p, (int)hl.size(), dDm.p, dE.p, lam.p, frow.p, lrow.p, Q, ldleaf, R0, fail.p, compact, (want_vectors && track_spans) ? colspan.p : nullptr);
#endif
    g_launches.launches++;
    pt.end(stream);
}

// ---- one tree level ---------------------------------------------------------------------------------
template <int BM, int BN, int BK, int WMs, int WNs, int STAGES>
static void launch_gemm(Stream s, const GemmProblem* probs, const GemmTile* tiles, const int* ntiles_ptr, long grid_want) {
#if CUPPEN_CUDA
    using Cfg = DmmaCfg<BM, BN, BK, WMs, WNs, STAGES>;
    auto kern = dgemm_dmma_kernel<BM, BN, BK, WMs, WNs, STAGES>;
    int grid = (int)std::max<long>(1, grid_want);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(probs, tiles, ntiles_ptr);
    CUDA_CHECK(cudaGetLastError());
#else
    (void)s; (void)probs; (void)tiles; (void)ntiles_ptr; (void)grid_want;
#endif
}

void Solver::update_descriptors() {
    double sigma = 0;
    for (double v : hD) sigma = std::max(sigma, fabs(v));
    for (double v : hE) sigma = std::max(sigma, fabs(v));
    if (!(sigma > 0)) sigma = 1.0;
    for (const LevelInfo& L : levels)
        for (size_t t = 0; t < L.ids.size(); ++t) {
            const PlanNode& nd = plan.nodes[L.ids[t]];
            MergeDesc& D = h_desc_all[L.desc_off + t];
            if (D.off != nd.off || D.m != nd.n || D.n1 != nd.n1) CUPPEN_THROW(CUPPEN_ERR_STATE, "divide tree changed shape");
            D.rho = nd.rho; D.theta = nd.theta; D.zscale = nd.zscale; D.sigma = sigma; D.mode = nd.mode; D.dthr = 1e-5 * scale;
        }
    if (!h_desc_all.empty()) dev_h2d(desc_all.p, h_desc_all.data(), sizeof(MergeDesc) * h_desc_all.size(), stream);
}

void Solver::prepare_levels() {
    drop_graph();
    h_leaves.clear();
    for (int id : plan.leaves) {
        const PlanNode& nd = plan.nodes[id];
        if (nd.off >= R0 && nd.off + nd.n <= R1) h_leaves.push_back(LeafDesc{nd.off, nd.n});
        else if (nd.off < R1 && nd.off + nd.n > R0) CUPPEN_THROW(CUPPEN_ERR_STATE, "leaf straddles a rank boundary");
    }
    if (!h_leaves.empty()) dev_h2d(leaves.p, h_leaves.data(), sizeof(LeafDesc) * h_leaves.size(), stream);
    if (select_mode) {
        std::vector<int> lo(n, 0), ln(n, 0);
        for (const LeafDesc& lf : h_leaves)
            for (int r = lf.off; r < lf.off + lf.n; ++r) { lo[r] = lf.off; ln[r] = lf.n; }
        dev_h2d(leaf_off_dev.p, lo.data(), sizeof(int) * n, stream);
        dev_h2d(leaf_n_dev.p, ln.data(), sizeof(int) * n, stream);
        dev_sync(stream);
    }
    // (phase, height) -> level; phase 0: nodes inside my subtree, phase 1: nodes above the subtrees
    std::map<std::pair<int, int>, std::vector<int>> groups;
    for (size_t id = 0; id < plan.nodes.size(); ++id) {
        const PlanNode& nd = plan.nodes[id];
        if (nd.left < 0) continue;
        const bool coop = nd.depth < glog;
        if (!coop && !(nd.off >= R0 && nd.off + nd.n <= R1)) continue;      // another rank's subtree
        groups[{coop ? 1 : 0, nd.height}].push_back((int)id);
    }
    levels.clear();
    first_coop = -1;
    h_desc_all.clear();
    std::vector<int> level_of(plan.nodes.size(), -1);        // schedule index at which a node's Q is produced
    for (auto& kv : groups) {
        LevelInfo L;
        L.coop = kv.first.first == 1;
        L.height = kv.first.second;
        L.ids = kv.second;
        std::sort(L.ids.begin(), L.ids.end(), [&](int a, int b) { return plan.nodes[a].off < plan.nodes[b].off; });
        if (L.coop && first_coop < 0) first_coop = (int)levels.size();
        for (int id : L.ids) level_of[id] = (int)levels.size();
        levels.push_back(L);
    }
    const int NL = (int)levels.size();
    std::vector<int> hnode((size_t)std::max(NL, 1) * n, -1);
    // local row range of a node's block in the layout of the phase in which it is consumed
    auto local_rows = [&](const PlanNode& nd, bool coop_layout, int& lr0, int& lsplit, int& lr1) {
        if (!coop_layout) { lr0 = nd.off - R0; lsplit = nd.off + nd.n1 - R0; lr1 = nd.off + nd.n - R0; return; }
        const int s0 = subtree_at(nd.off), s1 = subtree_at(nd.off + nd.n);
        const int sm = nd.left >= 0 ? subtree_at(nd.off + nd.n1) : s1;
        lr0 = crow0[s0]; lsplit = crow0[sm]; lr1 = crow0[s1];
    };
    for (int li = 0; li < NL; ++li) {
        LevelInfo& L = levels[li];
        L.desc_off = h_desc_all.size();
        for (size_t t = 0; t < L.ids.size(); ++t) {
            const PlanNode& nd = plan.nodes[L.ids[t]];
            MergeDesc D;
            memset(&D, 0, sizeof D);
            D.off = nd.off; D.n1 = nd.n1; D.n2 = nd.n - nd.n1; D.m = nd.n; D.mode = nd.mode;
            D.rho = nd.rho; D.theta = nd.theta; D.zscale = nd.zscale; D.sigma = 1.0; D.dthr = 1e-5;
            local_rows(nd, L.coop, D.lr0, D.lsplit, D.lr1);
            D.own_first = L.coop ? (comm.rank == 0) : 1;
            D.own_last = L.coop ? (comm.rank == G - 1) : 1;
            if (!want_vectors) { D.lr0 = D.lsplit = D.lr1 = 0; }
            h_desc_all.push_back(D);
            L.any_accurate = L.any_accurate || nd.mode == MODE_ACCURATE;
            for (int g = nd.off; g < nd.off + nd.n; ++g) hnode[(size_t)li * n + g] = (int)t;
            L.maxm = std::max(L.maxm, nd.n);
            L.maxm_rows = std::max(L.maxm_rows, D.lr1 - D.lr0);
            const long N = std::min(W > 0 ? W : nd.n, nd.n);
            for (int half = 0; half < 2; ++half) {
                const int rs = half ? D.lsplit : D.lr0, re = half ? D.lr1 : D.lsplit;
                if (re <= rs) continue;
                L.worst_tiles_big += (long)((re - rs + 127) / 128) * ((N + 127) / 128);
                L.worst_tiles_small += (long)((re - rs + 63) / 64) * ((N + 63) / 64);
            }
        }
    }
    if (desc_all.n < h_desc_all.size() + 1) desc_all.alloc(h_desc_all.size() + 1);
    if (node_of_all.n < hnode.size()) node_of_all.alloc(hnode.size());
    if (!h_desc_all.empty()) dev_h2d(desc_all.p, h_desc_all.data(), sizeof(MergeDesc) * h_desc_all.size(), stream);
    dev_h2d(node_of_all.p, hnode.data(), sizeof(int) * hnode.size(), stream);
    if (want_vectors) {
        long worst_t = 1;
        size_t worst_p = 2;
        for (const LevelInfo& L : levels) {
            worst_t = std::max(worst_t, std::max(L.worst_tiles_big, L.worst_tiles_small));
            worst_p = std::max(worst_p, 2 * L.ids.size());
        }
        if (tiles.n < (size_t)worst_t + 256) tiles.alloc((size_t)worst_t + 256);      // (+ the half tiles of a split last wave: at most one per two SMs)
        if (probs.n < worst_p) probs.alloc(worst_p);
    }
    if (ntiles_dev.n < 4) ntiles_dev.alloc(4);
    if (pin_desc_cap < h_desc_all.size() + 1) {
#if CUPPEN_CUDA
        if (pin_desc) cudaFreeHost(pin_desc);
        CUDA_CHECK(cudaMallocHost((void**)&pin_desc, sizeof(MergeDesc) * (h_desc_all.size() + 1)));
#else
        free(pin_desc);
        pin_desc = (MergeDesc*)malloc(sizeof(MergeDesc) * (h_desc_all.size() + 1));
#endif
        pin_desc_cap = h_desc_all.size() + 1;
    }
    dev_sync(stream);
}

// ---- transition to the cooperative phase: replicate the subtree vectors, redistribute the rows ------
// Peer-memory version: push the subtree vectors, then PULL the rows of the slice layout straight out of the peers' Q
// blocks into the other n x n buffer (no staging copies, no collective); the two buffers trade places.
void Solver::enter_cooperative_p2p() {
#if CUPPEN_CUDA
    pt.begin(T_COMM, stream);
    p2p_barrier();                                   // nobody is still reading the heap vectors of the previous solve
    launch_items(stream, R1 - R0, PushSubtreeVectors{p2p.H, lam.p, frow.p, lrow.p, R0});
    p2p_barrier();                                   // every subtree is finished: vectors replicated, Q blocks final
    if (want_vectors) {
        PullCtx pc;
        memset(&pc, 0, sizeof pc);
        pc.dst = Awork; pc.ldq = ldq; pc.S = S;
        int maxlen = 1;
        for (int s = 0; s < S; ++s) {
            const int o = owner_of_sub(s);           // layout L of the owner: local row = global row - its first row
            pc.src[s] = p2p.peerQ[o];
            pc.src_ld[s] = ldq;
            pc.sub_off[s] = sub_off[s];
            pc.slo[s] = sub_off[s] - rank_row0(o) + slice_lo(s, comm.rank);
            pc.crow0[s] = crow0[s];
            maxlen = std::max(maxlen, crow0[s + 1] - crow0[s]);
        }
        pc.sub_off[S] = n; pc.crow0[S] = crow0[S];
        dim3 grid((unsigned)n, (unsigned)std::max(1, std::min(16, (maxlen + 255) / 256)));
        p2p_pull_rows_kernel<<<grid, 256, 0, stream>>>(pc);
        CUDA_CHECK(cudaGetLastError());
        g_launches.launches++;
        p2p_barrier();                               // every rank has taken its slices: the source buffers may be reused
        std::swap(Qcur, Awork);
    }
    pt.end(stream);
#endif
}

void Solver::enter_cooperative() {
    if (G <= 1) return;
    if (want_vectors && track_spans) dev_d2d(colspan.p, colspan_sub.p, sizeof(RowSpan) * n, stream);     // row supports: the subtree blocks
    if (p2p.on) { enter_cooperative_p2p(); return; }
    // every rank has lam / first row / last row of its own subtree: zero the rest and sum
    for (DevBuf<double>* b : {&lam, &frow, &lrow}) {
        if (R0 > 0) dev_zero(b->p, sizeof(double) * R0, stream);
        if (R1 < n) dev_zero(b->p + R1, sizeof(double) * (n - R1), stream);
        comm.allreduce_sum(b->p, n, stream);
    }
    if (!want_vectors) return;
    // rows: each of my subtree blocks is cut into G slices; slice j goes to rank j.  Staging inside Apack: first the
    // send blocks (per destination: one contiguous len x sub_n[s] block per subtree of mine), then the receive blocks
    // (per source rank: one block per subtree it owns)
    const int me = comm.rank;
    auto copy_block = [&](double* dst, long dpitch, const double* src, long spitch, int rows, int cols) {
        if (rows <= 0 || cols <= 0) return;
#if CUPPEN_CUDA
        CUDA_CHECK(cudaMemcpy2DAsync(dst, sizeof(double) * dpitch, src, sizeof(double) * spitch, sizeof(double) * rows, cols,
                                     cudaMemcpyDeviceToDevice, stream));
#else
        for (int col = 0; col < cols; ++col) memcpy(dst + (size_t)col * dpitch, src + (size_t)col * spitch, sizeof(double) * rows);
#endif
    };
    std::vector<const void*> sp(G, nullptr);
    std::vector<void*> rp(G, nullptr);
    std::vector<size_t> sb(G, 0), rb(G, 0);
    size_t sofs = 0, rofs = 0;
    std::vector<size_t> rofs_of(S, 0);
    for (int j = 0; j < G; ++j) {
        sp[j] = Apack.p + sofs;
        if (j == me) continue;
        for (int t = first_sub(me); t < first_sub(me + 1); ++t) {
            const int lo = slice_lo(t, j), len = slice_lo(t, j + 1) - lo;
            copy_block(Apack.p + sofs, len, Qcur + (long)sub_off[t] * ldq + (sub_off[t] - R0) + lo, ldq, len, sub_n[t]);
            sofs += (size_t)len * sub_n[t];
        }
        sb[j] = sizeof(double) * (size_t)((Apack.p + sofs) - (const double*)sp[j]);
    }
    double* const rstage = Apack.p + round_up((long)sofs, 32);
    for (int o = 0; o < G; ++o) {
        rp[o] = rstage + rofs;
        for (int t = first_sub(o); t < first_sub(o + 1); ++t) {
            const int len = crow0[t + 1] - crow0[t];
            rofs_of[t] = rofs;
            rofs += (size_t)len * sub_n[t];
        }
        rb[o] = sizeof(double) * (size_t)((rstage + rofs) - (double*)rp[o]);
    }
    if ((size_t)round_up((long)sofs, 32) + rofs > Apack.n)
        CUPPEN_THROW(CUPPEN_ERR_STATE, "row redistribution needs %zu doubles of staging, Apack has %zu", (size_t)round_up((long)sofs, 32) + rofs, Apack.n);
    // my own slices move inside the buffer: stage them like received blocks
    for (int t = first_sub(me); t < first_sub(me + 1); ++t) {
        const int lo = slice_lo(t, me), len = slice_lo(t, me + 1) - lo;
        copy_block(rstage + rofs_of[t], len, Qcur + (long)sub_off[t] * ldq + (sub_off[t] - R0) + lo, ldq, len, sub_n[t]);
    }
    comm.alltoallv(sp, sb, rp, rb, stream);
    // unpack into Qcur in layout C: slice of subtree t at local rows crow0[t], columns of subtree t
    for (int t = 0; t < S; ++t)
        copy_block(Qcur + (long)sub_off[t] * ldq + crow0[t], ldq, rstage + rofs_of[t], crow0[t + 1] - crow0[t], crow0[t + 1] - crow0[t], sub_n[t]);
}

void Solver::run_level(int li) {
    const LevelInfo& L = levels[li];
    if (L.ids.empty()) return;
    if (li == first_coop) enter_cooperative();
    const int nd_cnt = (int)L.ids.size();
    const bool coop = L.coop && G > 1;
    LevelCtx c = level_ctx(li);
    RowCtx rowc{frow.p, lrow.p, frow2.p, lrow2.p, fpack.p, lpack.p};
    int lo_idx = n, hi_idx = 0;              // index range covered by this level's nodes
    for (int id : L.ids) { lo_idx = std::min(lo_idx, plan.nodes[id].off); hi_idx = std::max(hi_idx, plan.nodes[id].off + plan.nodes[id].n); }

#if CUPPEN_CUDA
    const bool fused = (!coop && L.maxm <= FUSE_MAXM);
#else
    const bool fused = false;
#endif
    if (fused) {
#if CUPPEN_CUDA
        pt.begin(T_DEFL, stream);
        launch_fused_front(stream, num_sms, nd_cnt, L.maxm, c, rowc, want_vectors ? 0 : 1);
        g_launches.launches++;
        pt.end(stream);
#endif
    } else {
        pt.begin(T_DEFL, stream);
        if (L.any_accurate) {
            launch_items(stream, n, ZAssemble{c});
#if CUPPEN_CUDA
            merge_tol_kernel<<<(unsigned)nd_cnt, 256, 0, stream>>>(c);
            CUDA_CHECK(cudaGetLastError());
            g_launches.launches++;
#else
            launch_warps(stream, nd_cnt, MergeTol{c});
#endif
            launch_items(stream, n, FlagDeflate{c});
        } else launch_items(stream, n, ZAssembleFlag{c});
#if CUPPEN_CUDA
        {
            const dim3 rk_grid((unsigned)((L.maxm + TL_TJ - 1) / TL_TJ), (unsigned)nd_cnt);
            rank_tiled_kernel<<<rk_grid, TL_THREADS, 0, stream>>>(c);
            CUDA_CHECK(cudaGetLastError());
            g_launches.launches++;
        }
#else
        launch_warps(stream, n, RankLive{c});
#endif
#if !CUPPEN_CUDA
        launch_items(stream, n, GivensSweep{c});
#endif
#if CUPPEN_CUDA
        launch_compact_scan(stream, nd_cnt, L.maxm, c);                                // (Givens sweep + compaction)
        g_launches.launches++;
#else
        launch_warps(stream, n, Compact{c});
#endif
        pt.end(stream);

        // secular roots: worst-case grid, the kernel reads the live count k of every merge on the device.
        // Cooperative levels: every rank solves its contiguous share of the roots of every merge; the
        // (tau, origin) arrays are zeroed first so that a sum over ranks assembles them.
        pt.begin(T_ROOT, stream);
        const int part = coop ? comm.rank : 0, nparts = coop ? G : 1;
        if (coop && !p2p.on) {
            dev_zero(tau.p + lo_idx, sizeof(double) * (hi_idx - lo_idx), stream);
            dev_zero(org.p + lo_idx, sizeof(int) * (hi_idx - lo_idx), stream);
        }
        const int per = (L.maxm + nparts - 1) / nparts;
#if CUPPEN_CUDA
        {
            int kcap = std::min((int)round_up(L.maxm, 32), (int)SEC_SMEM_K);
            size_t smem = (size_t)2 * kcap * sizeof(double);
            dim3 grid((unsigned)((per + SEC_WARPS - 1) / SEC_WARPS), (unsigned)nd_cnt);
            secular_kernel<<<grid, SEC_WARPS * 32, smem, stream>>>(c, kcap, part, nparts, (coop && p2p.on) ? p2p.H : SymHeap());
            CUDA_CHECK(cudaGetLastError());
        }
#else
        secular_host(c, nd_cnt, part, nparts);
#endif
        g_launches.launches++;
        pt.end(stream);
        if (coop && p2p.on) {
            pt.begin(T_COMM, stream);
            p2p_barrier();                           // every rank's root slice has landed in every heap (secular_kernel pushes)
            pt.end(stream);
        } else if (coop) {
            comm.allreduce_sum(tau.p + lo_idx, hi_idx - lo_idx, stream);
            comm.allreduce_sum_i32(org.p + lo_idx, hi_idx - lo_idx, stream);
            launch_items(stream, n, FillDorg{c});
        }
        pt.begin(T_EVX, stream);
#if CUPPEN_CUDA
        // tiled kernels from m = TILED_MIN_M on; below, the warp-per-output functors are a few microseconds
        // faster per launch (measured on `-s 1 -n 4096`, profiles/README.md)
        const bool tiled = L.maxm >= TILED_MIN_M;
        const dim3 tl_grid((unsigned)((L.maxm + TL_TJ - 1) / TL_TJ), (unsigned)nd_cnt);
        if (tiled && coop && p2p.on) {
            // cooperative level, peer memory: every rank computes its contiguous share of the Loewner vector and of the
            // norms (O(k^2 / G) each instead of O(k^2) replicated) and stores it into every rank's copy
            const int per_out = (int)round_up((L.maxm + G - 1) / G, TL_TJ);
            const dim3 part_grid((unsigned)(per_out / TL_TJ), (unsigned)nd_cnt);
            loewner_tiled_kernel<<<part_grid, TL_THREADS, 0, stream>>>(c, comm.rank, G, p2p.H);
            CUDA_CHECK(cudaGetLastError());
            g_launches.launches++;
            pt.end(stream);
            pt.begin(T_COMM, stream); p2p_barrier(); pt.end(stream);
            pt.begin(T_EVX, stream);
            norms_tiled_kernel<<<part_grid, TL_THREADS, 0, stream>>>(c, comm.rank, G, p2p.H);
            CUDA_CHECK(cudaGetLastError());
            g_launches.launches++;
            pt.end(stream);
            pt.begin(T_COMM, stream); p2p_barrier(); pt.end(stream);
            pt.begin(T_EVX, stream);
        } else if (tiled) {
            loewner_tiled_kernel<<<tl_grid, TL_THREADS, 0, stream>>>(c, 0, 1, SymHeap());
            CUDA_CHECK(cudaGetLastError());
            norms_tiled_kernel<<<tl_grid, TL_THREADS, 0, stream>>>(c, 0, 1, SymHeap());
            CUDA_CHECK(cudaGetLastError());
            g_launches.launches += 2;
        } else {
            launch_warps(stream, n, Loewner{c});
            launch_warps(stream, n, Norms{c});
        }
#else
        launch_warps(stream, n, Loewner{c});
        launch_warps(stream, n, Norms{c});
#endif
#if CUPPEN_CUDA
        if (!want_vectors) launch_items(stream, n, NewLambda{c});      // with eigenvectors: inside the first ugen_kernel launch
#else
        launch_items(stream, n, NewLambda{c});
#endif
        pt.end(stream);
        if (!want_vectors) {
            pt.begin(T_EVX, stream);
            launch_items(stream, n, RowPack{c, rowc});
#if CUPPEN_CUDA
            if (tiled) {
                rowgemv_tiled_kernel<<<tl_grid, TL_THREADS, 0, stream>>>(c, rowc);
                CUDA_CHECK(cudaGetLastError());
                g_launches.launches++;
            } else launch_warps(stream, n, RowGemv{c, rowc});
#else
            launch_warps(stream, n, RowGemv{c, rowc});
#endif
            launch_items(stream, n, RowCommit{c, rowc, frow.p, lrow.p});
            pt.end(stream);
        }
    }
    if (!want_vectors) return;

    MatCtx M = mat_ctx();
    pt.begin(T_PACK, stream);
#if CUPPEN_CUDA
    if (L.maxm_rows > 0) {
        dim3 grid((unsigned)n, (unsigned)((L.maxm_rows + PACK_THREADS * PACK_ROWS - 1) / (PACK_THREADS * PACK_ROWS)));
        pack_kernel<<<grid, PACK_THREADS, 0, stream>>>(c, M);          // (also zeroes the K tail of Apack)
        CUDA_CHECK(cudaGetLastError());
    }
#else
    pack_host(c, M);
    pack_tail_host(c, M, nd_cnt);
#endif
#if CUPPEN_CUDA
    g_launches.launches += 1;
#else
    g_launches.launches += 2;
#endif
    pt.end(stream);

    // panels of W root columns: U generation, device-built work list, one GEMM launch over all
    // (merge, half) problems of the level
    const bool small_tiles = (L.maxm_rows <= 256);
    const int BMN = small_tiles ? 64 : 128;
    for (int p0 = 0; p0 < L.maxm && L.maxm_rows > 0; p0 += W) {
        const int width = std::min(W, L.maxm - p0);
        WorkCtx w;
        w.desc = c.desc; w.nd = nd_cnt; w.p0 = p0; w.width = width; w.BM = BMN; w.BN = BMN;
        w.ldq = ldq; w.ldb = ldb; w.Apack = Awork; w.B = B.p; w.Qnext = Qcur; w.lidx = lidx.p;
        w.probs = probs.p; w.tiles = tiles.p; w.ntiles = ntiles_dev.p; w.tile_cap = (int)std::min<size_t>(tiles.n, 0x7fffffff);
        w.fail = fail.p + FAIL_TILES;
        w.supercol_mb = supercol_mb;
#if CUPPEN_CUDA
        w.split_grid = (!small_tiles && gemm_variant == 2 && split_tail) ? num_sms : 0;
#else
        w.split_grid = (!small_tiles && split_tail) ? 148 : 0;      // (host test build: the split tile list of a 148-SM device)
#endif
        pt.begin(T_UGEN, stream);
#if CUPPEN_CUDA
        {
            // (+1 block: the GEMM work list of the panel is built by the last block of the same launch)
            // column chunks per row tile: enough blocks to fill the SMs a few times over, but no more -- a block's prologue
            // (node -> descriptor -> K list -> pole, four dependent L2 round trips) is amortised over its columns
            const int row_tiles = (n + UG_ROWS - 1) / UG_ROWS;
            const int ychunks = std::max(1, std::min({8, (std::min(width, L.maxm) + 255) / 256, (12 * num_sms + row_tiles - 1) / row_tiles}));
            dim3 grid((unsigned)row_tiles + 1, (unsigned)ychunks);
            ugen_kernel<<<grid, 256, sizeof(int) * 2 * nd_cnt, stream>>>(c, M, p0, width, (!fused && p0 == 0) ? 1 : 0, w);
            CUDA_CHECK(cudaGetLastError());
        }
#else
        ugen_host(c, M, p0, width);
#endif
        g_launches.launches++;
        pt.end(stream);

        const long worst = small_tiles ? L.worst_tiles_small : L.worst_tiles_big;
        pt.begin(T_GEMM, stream);
#if CUPPEN_CUDA
        // (tensor-map kernel: up to twice the worst-case tile count, a short level is split into half tiles)
        const int grid = (int)std::min<long>((!small_tiles && w.split_grid > 0) ? 2 * worst : worst, small_tiles ? num_sms * 8L : (long)num_sms);
        if (small_tiles) launch_gemm<64, 64, 16, 2, 2, 3>(stream, probs.p, tiles.p, ntiles_dev.p, grid);
        else if (gemm_variant == 2)
            launch_gemm_tma(stream, probs.p, tiles.p, ntiles_dev.p, grid, fail.p + FAIL_TMA, Awork == Apack.p ? &map_apack : &map_qa, &map_b,
                            L.maxm >= GEMM_HINT_MIN_ROWS ? gemm_hints : 0);
        else if (gemm_variant == 1) launch_gemm_tma(stream, probs.p, tiles.p, ntiles_dev.p, grid, fail.p + FAIL_TMA);
        else launch_gemm<128, 128, 16, 2, 4, 3>(stream, probs.p, tiles.p, ntiles_dev.p, std::min<long>(worst, num_sms * 2L));
#else
        (void)worst;
        build_gemm_work_host(w);
        gemm_host(probs.p, tiles.p, ntiles_dev.p, BMN, BMN);
        g_launches.launches++;
#endif
        g_launches.launches++;
        pt.end(stream);
    }

    // (the boundary rows are pushed to the peers only when another level follows: a push after the last barrier of a solve
    // could land in a heap that its owner is already tearing down)
    const bool more_levels = li + 1 < (int)levels.size();
    if (c.Qz == nullptr) launch_items(stream, n, ExtractRows{c, Qcur, ldq, frow.p, lrow.p, (coop && p2p.on && more_levels) ? p2p.H : SymHeap()});
    if (coop && p2p.on) {
        // (first rows were pushed by rank 0, last rows by rank G-1, straight from the GEMM output)
        if (more_levels) { pt.begin(T_COMM, stream); p2p_barrier(); pt.end(stream); }
    } else if (coop && more_levels) {
        // first rows live on rank 0, last rows on rank G-1 (slice layout): replicate them for the next level
        comm.group_bcast(frow.p + lo_idx, sizeof(double) * (hi_idx - lo_idx), 0, 0, G, stream);
        comm.group_bcast(lrow.p + lo_idx, sizeof(double) * (hi_idx - lo_idx), G - 1, 0, G, stream);
    }
    // in place: blocks that wait for a higher parent simply stay where they are
}

// ---- final ordering, eigenvector gather, residuals -------------------------------------------------
void Solver::finish() {
    if (first_coop < 0 && G > 1) enter_cooperative();         // (cannot happen: G > 1 implies cooperative levels)
    launch_warps(stream, n, FinalRank{n, lam.p, perm.p, lam_sorted.p});     // (a tiled variant measured slower: DSETP-bound either way)
    dev_d2h(pin_lam, lam_sorted.p, sizeof(double) * n, stream);
    if (want_vectors) {
        pt.begin(T_RESID, stream);
        // V stays in storage order; (perm, lam_sorted) define the ascending order.  The sorted copy is
        // only materialised when the caller asks for the eigenvectors (cuppen_copy_eigenvectors).
        sorted_materialised = false;
        if (!(flags & CUPPEN_FLAG_NO_RESIDUALS)) {
            // slices of global rows held here: one (G == 1) or one per subtree
            struct Slice { int g0, l0, cnt; const double* lo; const double* hi; };
            std::vector<Slice> sl;
            if (G == 1) sl.push_back(Slice{0, 0, n, halo.p, halo.p});
            else if (p2p.on) {
                // halo rows: stored straight into the heap of the rank that needs them, one barrier
                HaloCtx hc;
                hc.H = p2p.H; hc.Q = Qcur; hc.ldq = ldq; hc.perm = perm.p; hc.halo_lo = p2p.halo_lo; hc.halo_hi = p2p.halo_hi; hc.n = n; hc.S = S;
                for (int s = 0; s <= S; ++s) hc.crow0[s] = crow0[s];
                launch_items(stream, (long)S * n, PushHaloRows{hc});
                p2p_barrier();
                for (int s = 0; s < S; ++s)
                    sl.push_back(Slice{sub_off[s] + slice_lo(s, comm.rank), crow0[s], crow0[s + 1] - crow0[s],
                                       p2p.halo_lo + (size_t)s * n, p2p.halo_hi + (size_t)s * n});
            } else {
                // halo rows: first and last local row of every slice of every rank
                for (int s = 0; s < S; ++s) {
                    launch_items(stream, n, ExtractRowVec{Qcur, ldq, (long)crow0[s], perm.p, halo.p + (size_t)(2 * s) * n});
                    launch_items(stream, n, ExtractRowVec{Qcur, ldq, (long)crow0[s + 1] - 1, perm.p, halo.p + (size_t)(2 * s + 1) * n});
                }
                comm.allgather(halo.p, halo_all.p, sizeof(double) * 2 * n * S, stream);
                auto row_of = [&](int rank, int s, int which) { return halo_all.p + ((size_t)rank * 2 * S + 2 * s + which) * n; };
                const int me = comm.rank;
                for (int s = 0; s < S; ++s) {
                    Slice x;
                    x.g0 = sub_off[s] + slice_lo(s, me); x.l0 = crow0[s]; x.cnt = crow0[s + 1] - crow0[s];
                    x.lo = (me > 0) ? row_of(me - 1, s, 1) : (s > 0 ? row_of(G - 1, s - 1, 1) : halo.p);
                    x.hi = (me < G - 1) ? row_of(me + 1, s, 0) : (s < S - 1 ? row_of(0, s + 1, 0) : halo.p);
                    sl.push_back(x);
                }
            }
#if CUPPEN_CUDA
            if (sl.size() > RES_MAX_SLICES) CUPPEN_THROW(CUPPEN_ERR_STATE, "%zu row slices per rank (at most %d)", sl.size(), (int)RES_MAX_SLICES);
            {
                ResSlices rs;
                memset(&rs, 0, sizeof rs);
                rs.ns = (int)sl.size();
                for (size_t i = 0; i < sl.size(); ++i) { rs.g0[i] = sl[i].g0; rs.l0[i] = sl[i].l0; rs.cnt[i] = sl[i].cnt; rs.lo[i] = sl[i].lo; rs.hi[i] = sl[i].hi; }
                launch_residual(stream, resid_variant, Qcur, ldq, n, rs, dOD.p, dOE.p, lam_sorted.p, perm.p, res2.p, track_spans ? colspan.p : nullptr);
                g_launches.launches++;
            }
#else
            for (size_t i = 0; i < sl.size(); ++i) {
                residual_host(Qcur, ldq, n, sl[i].g0, sl[i].l0, sl[i].cnt, dOD.p, dOE.p, lam_sorted.p, perm.p, sl[i].lo, sl[i].hi, res2.p, i > 0 ? 1 : 0, track_spans ? colspan.p : nullptr);
                g_launches.launches++;
            }
#endif
            if (p2p.on) {
                launch_items(stream, n, PushResidualPartials{p2p.H, res2.p, p2p.res_part, n});
                p2p_barrier();
                launch_items(stream, n, SumResidualPartials{p2p.res_part, res2.p, n, G});
            } else comm.allreduce_sum(res2.p, n, stream);
            dev_d2h(pin_res, res2.p, sizeof(double) * n, stream);
        }
        pt.end(stream);
    }
}

// column gather into ascending-lambda order (src/filehandling.c:315-321), on demand
void Solver::materialise_sorted() {
    if (sorted_materialised || !want_vectors) return;
#if CUPPEN_CUDA
    {
        dim3 grid((unsigned)n, (unsigned)std::max(1, std::min(64, (nloc_final + 255) / 256)));
        gather_cols_kernel<<<grid, 256, 0, stream>>>(Qcur, Awork, ldq, nloc_final, perm.p);
        CUDA_CHECK(cudaGetLastError());
    }
#else
    gather_cols_host(Qcur, Awork, ldq, nloc_final, perm.p, n);
#endif
    g_launches.launches++;
    std::swap(Qcur, Awork);          // until the next solve, which starts from Qa / Apack again
    dev_sync(stream);
    sorted_materialised = true;
}

// selected-eigenvector mode: push the unit vectors of the requested columns of the root down the tree
// (select_stages.h).  Runs after the (possibly graph-replayed) eigenvalue solve on the same stream.
void Solver::enqueue_apply() {
    if (!select_mode || h_sel.empty()) return;
    const int cnt = (int)h_sel_local.size();           // vectors of this rank (all of them on one rank)
    const int per = sel_per();
    double* Vloc = sel_world > 1 ? Vgath.p + (size_t)sel_rank * per * n : Vsel.p;
    double* rloc = sel_world > 1 ? res_gath.p + (size_t)sel_rank * per : res_sel.p;
#if CUPPEN_CUDA
    if (!ev_ap0) { CUDA_CHECK(cudaEventCreate(&ev_ap0)); CUDA_CHECK(cudaEventCreate(&ev_ap1)); }
    CUDA_CHECK(cudaEventRecord(ev_ap0, stream));
#endif
    for (int v0 = 0; v0 < cnt; v0 += SEL_NV) {
        SelCtx s;
        s.n = n; s.nv = std::min((int)SEL_NV, cnt - v0); s.sel = sel_dev.p + v0; s.perm = perm.p;
        s.X = selX.p; s.Y = selY.p; s.XS = selXS.p; s.Gam = selGam.p; s.dorg = sel_dorg.p;
        launch_items(stream, n, ApplyInit{s});
        for (int li = (int)levels.size() - 1; li >= 0; --li) {
            const LevelInfo& L = levels[li];
            if (L.ids.empty()) continue;
            LevelCtx c = level_ctx(li);
            launch_items(stream, n, ApplyPrep{c, s});
#if CUPPEN_CUDA
            {
                dim3 grid((unsigned)((L.maxm + CA_TJ - 1) / CA_TJ), (unsigned)L.ids.size());
                cauchy_apply_kernel<<<grid, CA_THREADS, 0, stream>>>(c, s);
                CUDA_CHECK(cudaGetLastError());
            }
#else
            cauchy_apply_host(c, s, (int)L.ids.size());
#endif
            g_launches.launches++;
            launch_items(stream, n, ApplyChains{c, s});
            std::swap(s.X, s.Y);
        }
        launch_items(stream, n, LeafApply{s, leaf_off_dev.p, leaf_n_dev.p, Qleaf.p, Vloc + (size_t)v0 * n});
    }
    if (cnt > 0) launch_warps(stream, cnt, SelResidual{n, Vloc, dOD.p, dOE.p, lam_sorted.p, sel_dev.p, rloc});
    if (sel_world > 1) {
        // every rank ends up with all vectors, in the caller's order
        sel_comm.allgather(Vloc, Vgath.p, sizeof(double) * (size_t)per * n, stream);
        sel_comm.allgather(rloc, res_gath.p, sizeof(double) * (size_t)per, stream);
        launch_items(stream, (long)h_sel.size() * n, SelReorder{Vgath.p, Vsel.p, res_gath.p, res_sel.p, n, sel_world, per});
    }
#if CUPPEN_CUDA
    CUDA_CHECK(cudaEventRecord(ev_ap1, stream));
#endif
}

// everything a solve does on the device, enqueued on `stream` without a single host synchronisation
void Solver::enqueue_solve() {
#if CUPPEN_CUDA
    pt.record(ev_begin, stream);
#endif
    Qcur = Qa.p;
    Awork = Apack.p;
    run_leaves();
    for (int li = 0; li < (int)levels.size(); ++li) run_level(li);
    if (!h_desc_all.empty()) dev_d2h(pin_desc, desc_all.p, sizeof(MergeDesc) * h_desc_all.size(), stream);
    finish();
    dev_d2h(pin_fail, fail.p, sizeof(int) * FAIL_INTS, stream);
    Qfinal = Qcur;
    Afinal = Awork;
#if CUPPEN_CUDA
    pt.record(ev_end, stream);
#endif
}

void Solver::solve() {
    if (!have_matrix) CUPPEN_THROW(CUPPEN_ERR_STATE, "cuppen_set_tridiagonal has not been called");
    const double t0 = wall_now();
    const long l0 = g_launches.launches;
#if CUPPEN_CUDA
    if (!ev_begin) { CUDA_CHECK(cudaEventCreate(&ev_begin)); CUDA_CHECK(cudaEventCreate(&ev_end)); }
#endif
    stats.clear();
    pt.reset();
    memset(&timers, 0, sizeof timers);
    acc_pack_bytes = acc_ugen_bytes = acc_gemm_flop = 0;
    sorted_materialised = false;
    bool replayed = false;
#if CUPPEN_CUDA
    if (use_graph && !graph_failed && graph_exec) {
        CUDA_CHECK(cudaGraphLaunch(graph_exec, stream));
        Qcur = Qfinal;
        Awork = Afinal;
        g_launches.launches += graph_launches;
        replayed = true;
    } else if (use_graph && !graph_failed && solves_done >= 1) {
        // second solve of this matrix: capture, instantiate, run
        pt.drop_spans();
        pt.capturing = true;
        cudaError_t ce = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
        if (ce == cudaSuccess) {
            try { enqueue_solve(); } catch (...) { cudaGraph_t g2 = nullptr; cudaStreamEndCapture(stream, &g2); if (g2) cudaGraphDestroy(g2); pt.capturing = false; graph_failed = true; throw; }
            ce = cudaStreamEndCapture(stream, &graph);
        }
        pt.capturing = false;
        if (ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&graph_exec, graph, 0);
        if (ce == cudaSuccess && graph_exec) {
            graph_launches = g_launches.launches - l0;
            pt.keep = true;
            CUDA_CHECK(cudaGraphLaunch(graph_exec, stream));
            replayed = true;
        } else {
            cudaGetLastError();
            graph_failed = true;
            pt.drop_spans();
            if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
        }
    }
#endif
    if (!replayed) enqueue_solve();
    const double t_ap = wall_now();
    enqueue_apply();
    const double t1 = wall_now();
    dev_sync(stream);
    if (pt.keep) phase_timers_pending = true;          // events of an instantiated graph: read on demand (cuppen_get_timers)
    else { pt.collect(); phase_timers_pending = false; }
    const double t2 = wall_now();
    if (select_mode && !h_sel.empty()) {
        h_res_sel.resize(h_sel.size());
        dev_d2h(h_res_sel.data(), res_sel.p, sizeof(double) * h_sel.size(), stream);
        dev_sync(stream);
        for (double& r : h_res_sel) r = sqrt(r) / scale;
    }
    solves_done++;
    const double unscale = 1.0 / scale;
    h_lam_sorted.assign(pin_lam, pin_lam + n);
    if (scale != 1.0) for (double& v : h_lam_sorted) v *= unscale;
    h_resid.clear();
    if (want_vectors && !(flags & CUPPEN_FLAG_NO_RESIDUALS)) {
        h_resid.assign(pin_res, pin_res + n);
        for (double& r : h_resid) r = sqrt(r) * unscale;
    }
    if (!h_desc_all.empty()) memcpy(h_desc_all.data(), pin_desc, sizeof(MergeDesc) * h_desc_all.size());
    int hfail[FAIL_INTS];
    memcpy(hfail, pin_fail, sizeof hfail);
    // per-merge records and executed work, from the descriptors the device filled in
    for (size_t li = 0; li < levels.size(); ++li)
        for (size_t t = 0; t < levels[li].ids.size(); ++t) {
            const MergeDesc& D = h_desc_all[levels[li].desc_off + t];
            const PlanNode& nd = plan.nodes[levels[li].ids[t]];
            cuppen_merge_stat st;
            st.offset = D.off; st.m = D.m; st.n1 = D.n1; st.mode = D.mode; st.zdefl = D.m - D.nlive1;
            st.givens = D.nlive1 - D.k; st.k = D.k; st.height = nd.height; st.rho = nd.beta * nd.theta / scale;
            stats.push_back(st);
            if (!want_vectors) continue;
            const double rows = D.lr1 - D.lr0;
            // in place: a z-deflated column only gets zeros in the other half's rows; every other column is read over
            // its own half and written either in full (deflated by a rotation) or over its own half into Apack (live)
            const double zd = D.m - D.nlive1, rot = D.nlive1 - D.k, live = D.k;
            acc_pack_bytes += 8.0 * rows * (0.5 * zd + 0.5 * (rot + live) + 1.0 * rot + 0.5 * live);
            acc_gemm_flop += 2.0 * D.k * ((double)(D.lsplit - D.lr0) * D.ktop + (double)(D.lr1 - D.lsplit) * D.kbot);
            acc_ugen_bytes += 8.0 * D.k * ((double)D.ktop + D.kbot);
        }
    timers.total_s = t1 - t0;
    bt_wall_extra = t2 - t1;
    timers.kernel_launches = g_launches.launches - l0;
    timers.pack_bytes = acc_pack_bytes;
    timers.ugen_bytes = acc_ugen_bytes;
    timers.gemm_flop = acc_gemm_flop;
#if CUPPEN_CUDA
    { float ms = 0; cudaEventElapsedTime(&ms, ev_begin, ev_end); timers.device_s = ms * 1e-3; }
    if (select_mode && !h_sel.empty()) {
        float ms = 0; cudaEventElapsedTime(&ms, ev_ap0, ev_ap1);
        timers.apply_s = ms * 1e-3;
        timers.device_s += timers.apply_s;
    }
#else
    if (select_mode && !h_sel.empty()) timers.apply_s = t2 - t_ap;
    timers.device_s = t2 - t0;
#endif
    fill_phase_timers();                               // (zeros for now when the phase events are still pending)
#if CUPPEN_CUDA
    tma_check_abort(hfail + FAIL_TMA);
#endif
    if (hfail[FAIL_TILES] != 0)
        CUPPEN_THROW(CUPPEN_ERR_STATE, "GEMM tile list overflow (%zu tiles allocated): the eigenvectors of this solve are incomplete", tiles.n);
    if (hfail[FAIL_COMM] != 0)
        CUPPEN_THROW(CUPPEN_ERR_COMM, "peer barrier timed out waiting for rank %d (a peer fell out of the solve)", hfail[FAIL_COMM] - 1);
    if (hfail[0] != 0) CUPPEN_THROW(CUPPEN_ERR_CONVERGENCE, "leaf QL iteration did not converge (row %d)", hfail[0] - 1);
    solved = true;
}

}  // namespace cuppen

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace cuppen;

struct cuppen_handle_s {
    Solver s;
};

#define CUPPEN_API_BEGIN try {
#define CUPPEN_API_END                                                     \
    }                                                                      \
    catch (const cuppen::Error& e) { g_last_error = e.msg; return e.code; } \
    catch (const std::bad_alloc&) { g_last_error = "out of host memory"; return CUPPEN_ERR_NOMEM; } \
    catch (...) { g_last_error = "unknown error"; return CUPPEN_ERR_ARG; }   \
    return CUPPEN_OK;

static int create_common(cuppen_handle* h, int n, int ref_leaves, int flags, int device, Comm comm) {
    CUPPEN_API_BEGIN
    if (!h || n < 1 || ref_leaves < 1) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument (n=%d, ref_leaves=%d)", n, ref_leaves);
    if (n / ref_leaves == 0) CUPPEN_THROW(CUPPEN_ERR_LEAF, "Leaf Size is too small! Reduce number of tasks.");
    if ((flags & CUPPEN_FLAG_SELECT) && (flags & CUPPEN_FLAG_VECTORS))
        CUPPEN_THROW(CUPPEN_ERR_ARG, "CUPPEN_FLAG_SELECT and CUPPEN_FLAG_VECTORS are exclusive");
#if CUPPEN_CUDA
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        CUPPEN_THROW(CUPPEN_ERR_CUDA, "no CUDA device available (%s); libcuppen_b200 has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) CUPPEN_THROW(CUPPEN_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CUDA_CHECK(cudaSetDevice(device));
#endif
    cuppen_handle_s* hs = new cuppen_handle_s();
    Solver& s = hs->s;
    s.n = n; s.P = ref_leaves; s.flags = flags; s.device = device; s.want_vectors = (flags & CUPPEN_FLAG_VECTORS) != 0;
    s.select_mode = (flags & CUPPEN_FLAG_SELECT) != 0;
    s.comm = comm;
    if (s.select_mode && comm.world > 1) {
        // replicated eigenvalue-only decomposition, the selected vectors dealt to the ranks: the communicator is only
        // used to gather them
        s.sel_comm = comm; s.sel_rank = comm.rank; s.sel_world = comm.world;
        s.comm = Comm();
    }
    try {
#if CUPPEN_CUDA
        CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
#endif
        // the tree shape depends only on (n, P): plan with a dummy matrix to size the buffers
        std::vector<double> D0(n, 1.0), E0(std::max(1, n - 1), 1.0);
        {
            const char* le = getenv("CUPPEN_LEAF");     // leaf size of the accurate sub-tree (4..32), default 16
            if (le && atoi(le) >= 4 && atoi(le) <= LEAF_MAX) s.leaf_max = atoi(le);
        }
        if (build_plan(s.plan, n, D0.data(), E0.data(), ref_leaves, s.leaf_max) != 0)
            CUPPEN_THROW(CUPPEN_ERR_LEAF, "Leaf Size is too small! Reduce number of tasks.");
        s.init_layout();
        s.allocate();
        s.prepare_levels();
    } catch (...) { delete hs; throw; }
    *h = hs;
    CUPPEN_API_END
}

extern "C" {

int cuppen_create(cuppen_handle* h, int n, int ref_leaves, int flags, int device) {
    return create_common(h, n, ref_leaves, flags, device, Comm());
}

int cuppen_create_callbacks(cuppen_handle* h, int n, int ref_leaves, int flags, int device, int rank, int world,
                            const cuppen_comm_callbacks* cb) {
    CUPPEN_API_BEGIN
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !cb)) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad rank/world");
    Comm c;
    c.rank = rank; c.world = world;
    if (cb) { c.cb = *cb; c.use_cb = true; }
    int rc = create_common(h, n, ref_leaves, flags, device, c);
    if (rc != 0) return rc;
    CUPPEN_API_END
}

int cuppen_nccl_unique_id(unsigned char id[CUPPEN_NCCL_ID_BYTES]) {
    CUPPEN_API_BEGIN
    nccl_unique_id(id);
    CUPPEN_API_END
}

int cuppen_create_nccl(cuppen_handle* h, int n, int ref_leaves, int flags, int device, int rank, int world,
                       const unsigned char id[CUPPEN_NCCL_ID_BYTES]) {
    CUPPEN_API_BEGIN
    if (world < 1 || rank < 0 || rank >= world) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad rank/world");
    Comm c;
    c.rank = rank; c.world = world;
    if (world > 1) {
#if CUPPEN_CUDA
        CUDA_CHECK(cudaSetDevice(device));
#endif
        c.init_nccl(id);
    }
    int rc = create_common(h, n, ref_leaves, flags, device, c);
    if (rc != 0) return rc;
    CUPPEN_API_END
}

int cuppen_destroy(cuppen_handle h) {
    CUPPEN_API_BEGIN
    if (h) {
        h->s.comm.destroy();
        h->s.sel_comm.destroy();
#if CUPPEN_CUDA
        if (h->s.stream) cudaStreamDestroy(h->s.stream);
#endif
        delete h;
    }
    CUPPEN_API_END
}

int cuppen_set_tridiagonal(cuppen_handle h, const double* D, const double* E) {
    CUPPEN_API_BEGIN
    if (!h || !D || (!E && h->s.n > 1)) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(h->s.device));
#endif
    h->s.set_matrix(D, E);
    CUPPEN_API_END
}

int cuppen_solve(cuppen_handle h) {
    CUPPEN_API_BEGIN
    if (!h) CUPPEN_THROW(CUPPEN_ERR_ARG, "null handle");
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(h->s.device));
#endif
    h->s.solve();
    CUPPEN_API_END
}

int cuppen_get_eigenvalues(cuppen_handle h, double* out) {
    CUPPEN_API_BEGIN
    if (!h || !out) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    if (!h->s.solved) CUPPEN_THROW(CUPPEN_ERR_STATE, "not solved");
    memcpy(out, h->s.h_lam_sorted.data(), sizeof(double) * h->s.n);
    CUPPEN_API_END
}

int cuppen_get_residuals(cuppen_handle h, const int* idx, int cnt, double* out) {
    CUPPEN_API_BEGIN
    if (!h || !out) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (!s.solved) CUPPEN_THROW(CUPPEN_ERR_STATE, "not solved");
    if (s.select_mode) {
        // residuals exist for the selected ranks only; idx == NULL: length-n array, NaN where not selected
        std::map<int, double> have;
        for (size_t t = 0; t < s.h_sel.size() && t < s.h_res_sel.size(); ++t) have[s.h_sel[t]] = s.h_res_sel[t];
        if (!idx) {
            for (int i = 0; i < s.n; ++i) out[i] = NAN;
            for (auto& kv : have) out[kv.first] = kv.second;
        } else
            for (int i = 0; i < cnt; ++i) {
                auto it = have.find(idx[i]);
                if (it == have.end()) CUPPEN_THROW(CUPPEN_ERR_ARG, "eigenvector %d was not selected", idx[i]);
                out[i] = it->second;
            }
        return CUPPEN_OK;
    }
    if ((int)s.h_resid.size() != s.n) CUPPEN_THROW(CUPPEN_ERR_STATE, "residuals need CUPPEN_FLAG_VECTORS");
    if (!idx) { memcpy(out, s.h_resid.data(), sizeof(double) * s.n); }
    else
        for (int i = 0; i < cnt; ++i) {
            if (idx[i] < 0 || idx[i] >= s.n) CUPPEN_THROW(CUPPEN_ERR_ARG, "eigenvector index %d out of range", idx[i]);
            out[i] = s.h_resid[idx[i]];
        }
    CUPPEN_API_END
}

int cuppen_get_merge_stats(cuppen_handle h, cuppen_merge_stat* out, int capacity, int* count) {
    CUPPEN_API_BEGIN
    if (!h || !count) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    *count = (int)h->s.stats.size();
    if (out)
        for (int i = 0; i < *count && i < capacity; ++i) out[i] = h->s.stats[i];
    CUPPEN_API_END
}

int cuppen_get_timers(cuppen_handle h, cuppen_timers* out) {
    CUPPEN_API_BEGIN
    if (!h || !out) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (s.phase_timers_pending) {
#if CUPPEN_CUDA
        CUDA_CHECK(cudaSetDevice(s.device));
#endif
        s.pt.reset();
        s.pt.collect();
        s.fill_phase_timers();
        s.phase_timers_pending = false;
    }
    *out = h->s.timers;
    CUPPEN_API_END
}

int cuppen_local_rows(cuppen_handle h, int* row0, int* rows) {
    CUPPEN_API_BEGIN
    if (!h || !row0 || !rows) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    *row0 = (h->s.G > 1) ? h->s.sub_off[0] + h->s.slice_lo(0, h->s.comm.rank) : 0;
    *rows = h->s.nloc_final;
    CUPPEN_API_END
}

int cuppen_local_row_map(cuppen_handle h, int* global_rows) {
    CUPPEN_API_BEGIN
    if (!h || !global_rows) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (s.G == 1) { for (int r = 0; r < s.n; ++r) global_rows[r] = r; }
    else
        for (int sub = 0; sub < s.S; ++sub)
            for (int l = s.crow0[sub]; l < s.crow0[sub + 1]; ++l)
                global_rows[l] = s.sub_off[sub] + s.slice_lo(sub, s.comm.rank) + (l - s.crow0[sub]);
    CUPPEN_API_END
}

int cuppen_copy_eigenvectors(cuppen_handle h, double* V, long ld) {
    CUPPEN_API_BEGIN
    if (!h || !V) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (!s.solved || !s.want_vectors) CUPPEN_THROW(CUPPEN_ERR_STATE, "no eigenvectors (solve with CUPPEN_FLAG_VECTORS)");
    if (ld < s.nloc_final) CUPPEN_THROW(CUPPEN_ERR_ARG, "ld too small");
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(s.device));
#endif
    s.materialise_sorted();
#if CUPPEN_CUDA
    CUDA_CHECK(cudaMemcpy2DAsync(V, sizeof(double) * ld, s.Qcur, sizeof(double) * s.ldq, sizeof(double) * s.nloc_final, s.n,
                                 cudaMemcpyDeviceToHost, s.stream));
    dev_sync(s.stream);
#else
    for (int c = 0; c < s.n; ++c) memcpy(V + (long)c * ld, s.Qcur + (long)c * s.ldq, sizeof(double) * s.nloc_final);
#endif
    CUPPEN_API_END
}

// selected columns of V (ascending-lambda ranks idx[0..cnt)), local rows only: rows x cnt, column-major
int cuppen_copy_eigenvector_columns(cuppen_handle h, const int* idx, int cnt, double* V, long ld) {
    CUPPEN_API_BEGIN
    if (!h || cnt < 0 || (cnt > 0 && (!idx || !V))) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument");
    Solver& s = h->s;
    if (!s.solved || !s.want_vectors) CUPPEN_THROW(CUPPEN_ERR_STATE, "no eigenvectors (solve with CUPPEN_FLAG_VECTORS)");
    if (ld < s.nloc_final) CUPPEN_THROW(CUPPEN_ERR_ARG, "ld too small");
    for (int i = 0; i < cnt; ++i)
        if (idx[i] < 0 || idx[i] >= s.n) CUPPEN_THROW(CUPPEN_ERR_ARG, "eigenvector index %d out of range", idx[i]);
    if (cnt == 0) return CUPPEN_OK;
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(s.device));
#endif
    DevBuf<int> didx;
    DevBuf<double> tmp;
    didx.alloc((size_t)cnt);
    tmp.alloc((size_t)cnt * s.ldq);
    dev_h2d(didx.p, idx, sizeof(int) * cnt, s.stream);
    const int* perm = s.sorted_materialised ? nullptr : s.perm.p;      // after the sorted gather the storage order IS the rank order
#if CUPPEN_CUDA
    {
        dim3 grid((unsigned)cnt, (unsigned)std::max(1, std::min(64, (s.nloc_final + 255) / 256)));
        gather_sel_cols_kernel<<<grid, 256, 0, s.stream>>>(s.Qcur, tmp.p, s.ldq, s.nloc_final, perm, didx.p);
        CUDA_CHECK(cudaGetLastError());
        g_launches.launches++;
        CUDA_CHECK(cudaMemcpy2DAsync(V, sizeof(double) * ld, tmp.p, sizeof(double) * s.ldq, sizeof(double) * s.nloc_final, cnt,
                                     cudaMemcpyDeviceToHost, s.stream));
    }
    dev_sync(s.stream);
#else
    for (int c = 0; c < cnt; ++c) {
        const int col = perm ? perm[idx[c]] : idx[c];
        memcpy(V + (long)c * ld, s.Qcur + (long)col * s.ldq, sizeof(double) * s.nloc_final);
    }
#endif
    CUPPEN_API_END
}

int cuppen_select_eigenvectors(cuppen_handle h, const int* idx, int cnt) {
    CUPPEN_API_BEGIN
    if (!h || cnt < 0 || (cnt > 0 && !idx)) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument");
    Solver& s = h->s;
    if (!s.select_mode) CUPPEN_THROW(CUPPEN_ERR_STATE, "the handle was not created with CUPPEN_FLAG_SELECT");
    for (int i = 0; i < cnt; ++i)
        if (idx[i] < 0 || idx[i] >= s.n) CUPPEN_THROW(CUPPEN_ERR_ARG, "eigenvector index %d out of range", idx[i]);
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(s.device));
#endif
    s.h_sel.assign(idx, idx + cnt);
    s.h_sel_local.clear();
    for (int t = s.sel_rank; t < cnt; t += s.sel_world) s.h_sel_local.push_back(idx[t]);
    s.h_res_sel.clear();
    if (cnt > 0) {
        if (s.sel_dev.n < (size_t)cnt + SEL_NV) { s.sel_dev.alloc((size_t)cnt + SEL_NV); s.res_sel.alloc((size_t)cnt + SEL_NV); }
        if (s.Vsel.n < (size_t)cnt * s.n) s.Vsel.alloc((size_t)cnt * s.n);
        if (s.sel_world > 1) {
            const size_t slots = (size_t)s.sel_per() * s.sel_world;
            if (s.Vgath.n < slots * s.n) s.Vgath.alloc(slots * s.n);
            if (s.res_gath.n < slots + SEL_NV) s.res_gath.alloc(slots + SEL_NV);
        }
        if (!s.h_sel_local.empty()) dev_h2d(s.sel_dev.p, s.h_sel_local.data(), sizeof(int) * s.h_sel_local.size(), s.stream);
        dev_sync(s.stream);
    }
    s.solved = false;
    CUPPEN_API_END
}

int cuppen_copy_selected_eigenvectors(cuppen_handle h, double* V, long ld) {
    CUPPEN_API_BEGIN
    if (!h || !V) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (!s.select_mode || !s.solved) CUPPEN_THROW(CUPPEN_ERR_STATE, "no selected eigenvectors (CUPPEN_FLAG_SELECT, select, solve)");
    if (ld < s.n) CUPPEN_THROW(CUPPEN_ERR_ARG, "ld too small");
    const size_t cnt = s.h_sel.size();
    if (cnt == 0) return CUPPEN_OK;
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(s.device));
    CUDA_CHECK(cudaMemcpy2DAsync(V, sizeof(double) * ld, s.Vsel.p, sizeof(double) * s.n, sizeof(double) * s.n, cnt,
                                 cudaMemcpyDeviceToHost, s.stream));
    dev_sync(s.stream);
#else
    for (size_t t = 0; t < cnt; ++t) memcpy(V + (long)t * ld, s.Vsel.p + t * s.n, sizeof(double) * s.n);
#endif
    CUPPEN_API_END
}

int cuppen_orthogonality(cuppen_handle h, double* max_abs_dev, double* seconds) {
    CUPPEN_API_BEGIN
    if (!h || !max_abs_dev) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (!s.solved || !s.want_vectors) CUPPEN_THROW(CUPPEN_ERR_STATE, "no eigenvectors (solve with CUPPEN_FLAG_VECTORS)");
    if (s.G > 8) CUPPEN_THROW(CUPPEN_ERR_ARG, "the orthogonality check supports up to 8 ranks");
    const int G = s.G;
    // several ranks (every rank calls): the row slices are all-gathered, every rank evaluates its share of the tiles of
    // the Gram triangle and the maxima are combined -- n^3 / G flop per GPU, 8 n^2 bytes received per GPU
    DevBuf<double> gath, vfull;
    const double* V = s.Qcur;
    long ld = s.ldq;
    int rows = s.nloc_final;
    const double t_wall0 = wall_now();
    if (G > 1) {
        gath.alloc((size_t)G * s.n * s.ldq);
        vfull.alloc((size_t)G * s.n * s.ldq);
        s.comm.allgather(s.Qcur, gath.p, sizeof(double) * (size_t)s.n * s.ldq, s.stream);
        GramGather gg;
        gg.gath = gath.p; gg.V = vfull.p; gg.ldq = s.ldq; gg.n = s.n; gg.G = G;
        for (int r = 0; r < G; ++r) {
            int c = 0;
            for (int sub = 0; sub < s.S; ++sub) c += s.slice_lo(sub, r + 1) - s.slice_lo(sub, r);
            gg.nloc[r] = c;
        }
        launch_items(s.stream, (long)G * s.ldq * s.n, gg);
        V = vfull.p; ld = (long)G * s.ldq; rows = (int)ld;
    }
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(s.device));
    DevBuf<unsigned long long> res;
    res.alloc(8);
    dev_zero(res.p, sizeof(unsigned long long) * 8, s.stream);
    const long T = (s.n + GR_BT - 1) / GR_BT, ntiles = T * (T + 1) / 2;
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    CUDA_CHECK(cudaEventRecord(e0, s.stream));
    gram_check_kernel<<<(unsigned)std::max<long>(1, std::min<long>((ntiles + G - 1) / G, s.num_sms)), GR_THREADS, gram_smem_bytes(), s.stream>>>(
        V, ld, rows, s.n, res.p, s.comm.rank, G);
    CUDA_CHECK(cudaGetLastError());
    g_launches.launches++;
    CUDA_CHECK(cudaEventRecord(e1, s.stream));
    double worst = 0;
    if (G > 1) {
        DevBuf<unsigned long long> all;
        all.alloc(G);
        s.comm.allgather(res.p, all.p, sizeof(unsigned long long), s.stream);
        std::vector<unsigned long long> hb(G);
        dev_d2h(hb.data(), all.p, sizeof(unsigned long long) * G, s.stream);
        dev_sync(s.stream);
        for (int r = 0; r < G; ++r) { double v; memcpy(&v, &hb[r], sizeof v); worst = std::max(worst, v); }
    } else {
        unsigned long long bits = 0;
        dev_d2h(&bits, res.p, sizeof bits, s.stream);
        dev_sync(s.stream);
        memcpy(&worst, &bits, sizeof(double));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *max_abs_dev = worst;
    if (seconds) *seconds = (G > 1) ? wall_now() - t_wall0 : ms * 1e-3;
#else
    // test-only host build: every rank evaluates the whole Gram matrix of the gathered rows
    *max_abs_dev = gram_check_host(V, ld, rows, s.n);
    if (seconds) *seconds = wall_now() - t_wall0;
#endif
    CUPPEN_API_END
}

// Eigenvector file (the reference cannot emit V, SURVEY.md finding 6).  Layout, little endian:
//   char[8] "CUPPENV1" | int64 n | int64 ncols | int64 rank[ncols] (0-based, ascending lambda) |
//   double lambda[ncols] | double V[ncols][n] (one eigenvector after the other)
int cuppen_write_eigenvectors(cuppen_handle h, const char* filename) {
    CUPPEN_API_BEGIN
    if (!h) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    Solver& s = h->s;
    if (!s.solved || !(s.want_vectors || s.select_mode)) CUPPEN_THROW(CUPPEN_ERR_STATE, "no eigenvectors to write");
    // several ranks: every rank calls; rank 0 writes (selected mode: every rank holds all selected vectors)
    const bool writer = s.select_mode ? (s.sel_rank == 0) : (s.comm.rank == 0);
    if (writer && !filename) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
#if CUPPEN_CUDA
    CUDA_CHECK(cudaSetDevice(s.device));
#endif
    const long long n = s.n;
    std::vector<long long> ranks;
    if (s.select_mode) ranks.assign(s.h_sel.begin(), s.h_sel.end());
    else { ranks.resize(n); for (long long i = 0; i < n; ++i) ranks[i] = i; }
    const long long ncols = (long long)ranks.size();
    std::vector<double> lam(ncols);
    for (long long t = 0; t < ncols; ++t) lam[t] = s.h_lam_sorted[ranks[t]];
    FILE* f = nullptr;
    bool ok = true;
    if (writer) {
        f = fopen(filename, "wb");
        if (!f) fprintf(stderr, "Could not open file\n");
        ok = f != nullptr;
        ok = ok && fwrite("CUPPENV1", 1, 8, f) == 8 && fwrite(&n, 8, 1, f) == 1 && fwrite(&ncols, 8, 1, f) == 1;
        ok = ok && (ncols == 0 || (fwrite(ranks.data(), 8, ncols, f) == (size_t)ncols && fwrite(lam.data(), 8, ncols, f) == (size_t)ncols));
    }
    const bool distributed = !s.select_mode && s.G > 1;
    if (distributed) {
        // all ranks must agree before the collective part starts
        DevBuf<double> flag;
        flag.alloc(8);
        double bad = ok ? 0.0 : 1.0;
        dev_h2d(flag.p, &bad, sizeof bad, s.stream);
        s.comm.allreduce_sum(flag.p, 1, s.stream);
        dev_d2h(&bad, flag.p, sizeof bad, s.stream);
        dev_sync(s.stream);
        if (bad != 0.0) { if (f) fclose(f); CUPPEN_THROW(CUPPEN_ERR_IO, "cannot open %s", filename ? filename : "(rank 0's file)"); }
        s.materialise_sorted();
        // panels of sorted columns: all-gather the ranks' row slices (ldq x w doubles each), rank 0 puts the rows in place
        const int G = s.G;
        const long long w = std::max<long long>(1, std::min<long long>(n, (256LL << 20) / (8LL * G * s.ldq)));
        DevBuf<double> gath;
        gath.alloc((size_t)G * s.ldq * w);
        std::vector<double> hg, buf;
        std::vector<std::vector<int>> rowmap(G);
        if (writer) {
            hg.resize((size_t)G * s.ldq * w);
            buf.resize((size_t)w * n);
            for (int r = 0; r < G; ++r)
                for (int sub = 0; sub < s.S; ++sub)
                    for (int i = s.slice_lo(sub, r); i < s.slice_lo(sub, r + 1); ++i) rowmap[r].push_back(s.sub_off[sub] + i);
        }
        for (long long c0 = 0; c0 < n; c0 += w) {
            const long long wc = std::min(w, n - c0);
            s.comm.allgather(s.Qcur + c0 * s.ldq, gath.p, sizeof(double) * (size_t)s.ldq * wc, s.stream);
            if (!writer) continue;
            dev_d2h(hg.data(), gath.p, sizeof(double) * (size_t)G * s.ldq * wc, s.stream);
            dev_sync(s.stream);
            for (int r = 0; r < G; ++r)
                for (long long c = 0; c < wc; ++c) {
                    const double* src = hg.data() + ((size_t)r * wc + c) * s.ldq;
                    double* dst = buf.data() + c * n;
                    for (size_t l = 0; l < rowmap[r].size(); ++l) dst[rowmap[r][l]] = src[l];
                }
            ok = ok && fwrite(buf.data(), 8, (size_t)(wc * n), f) == (size_t)(wc * n);
        }
        dev_sync(s.stream);
    } else if (writer) {
        const double* src = nullptr;
        long ld = 0;
        if (s.select_mode) { src = s.Vsel.p; ld = s.n; }
        else { s.materialise_sorted(); src = s.Qcur; ld = s.ldq; }
        const long long panel = std::max<long long>(1, (64LL << 20) / (8 * n));        // <= 64 MB of host staging
        std::vector<double> buf((size_t)std::min(panel, std::max<long long>(ncols, 1)) * n);
        for (long long c0 = 0; ok && c0 < ncols; c0 += panel) {
            const long long w = std::min(panel, ncols - c0);
#if CUPPEN_CUDA
            CUDA_CHECK(cudaMemcpy2DAsync(buf.data(), sizeof(double) * n, src + c0 * ld, sizeof(double) * ld, sizeof(double) * n, w,
                                         cudaMemcpyDeviceToHost, s.stream));
            dev_sync(s.stream);
#else
            for (long long c = 0; c < w; ++c) memcpy(buf.data() + c * n, src + (c0 + c) * ld, sizeof(double) * n);
#endif
            ok = fwrite(buf.data(), 8, (size_t)(w * n), f) == (size_t)(w * n);
        }
    }
    if (f) ok = (fclose(f) == 0) && ok;
    if (!ok) CUPPEN_THROW(CUPPEN_ERR_IO, "write error on %s", filename ? filename : "(null)");
    CUPPEN_API_END
}

// Dense symmetric eigenproblem A = Z diag(W) Z^T on one GPU: blocked Householder
// tridiagonalisation (dense_stages.h), the tridiagonal path of this library under the accurate rule, and the
// back-transformation of its eigenvectors through the block reflectors.  A: host, column-major, symmetric (the lower
// triangle is read); W: n eigenvalues ascending; Z (may be NULL: eigenvalues only): n x n, column-major, ld >= n.
int cuppen_dense_eigh(int n, const double* A, long lda, double* W, double* Z, long ldz, int device, cuppen_dense_timers* tm) {
    CUPPEN_API_BEGIN
    if (n < 1 || !A || !W || lda < n || (Z && ldz < n)) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument");
    for (long c = 0; c < n; ++c)
        for (long r = c; r < n; ++r)
            if (!std::isfinite(A[c * lda + r])) CUPPEN_THROW(CUPPEN_ERR_ARG, "A(%ld,%ld) is not finite", r, c);
#if CUPPEN_CUDA
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        CUPPEN_THROW(CUPPEN_ERR_CUDA, "no CUDA device available; libcuppen_b200 has no CPU path");
    if (device < 0 || device >= ndev) CUPPEN_THROW(CUPPEN_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CUDA_CHECK(cudaSetDevice(device));
    set_kernel_attributes();
    cudaStream_t st;
    CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cuppen_handle hs = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    try {
        for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
        const long ldA = round_up(n, 16), ldp = round_up(n, 16) + 256;
        DevBuf<double> dA, Vp, Wp, PW, VT, part, dots, wtmp, wpart, dd, de, dtau, Tm, W1, Gp;
        DevBuf<int> iota, ntl;
        DevBuf<GemmProblem> probs;
        DevBuf<GemmTile> tiles;
        dA.alloc((size_t)ldA * n + 1024);
        for (DevBuf<double>* b : {&Vp, &Wp, &PW, &VT}) { b->alloc((size_t)ldp * DN_NB + 1024); dev_zero(b->p, b->bytes(), st); }
        part.alloc((size_t)DN_SPLIT * n); dots.alloc(2 * DN_NB); wtmp.alloc((size_t)n + 256); wpart.alloc((size_t)n / 256 + 8);
        dd.alloc((size_t)n + 8); de.alloc((size_t)n + 8); dtau.alloc((size_t)n + 8); Tm.alloc(DN_NB * DN_NB);
        W1.alloc((size_t)DN_NB * ldp + 1024);
        Gp.alloc((size_t)((n + DN_GRAM_ROWS - 1) / DN_GRAM_ROWS + 1) * DN_NB * DN_NB);
        iota.alloc((size_t)n + 256); ntl.alloc(4); probs.alloc(2);
        const long maxt = ((n + 127) / 128) * (long)((n + 127) / 128) + 1;
        tiles.alloc((size_t)maxt);
        dev_zero(W1.p, W1.bytes(), st);
        dev_zero(de.p, de.bytes(), st);
        dense_iota_kernel<<<(unsigned)((n + 256 + 255) / 256), 256, 0, st>>>(iota.p, n + 256);
        if (lda == ldA) CUDA_CHECK(cudaMemcpyAsync(dA.p, A, sizeof(double) * (size_t)lda * n, cudaMemcpyHostToDevice, st));     // one contiguous block
        else CUDA_CHECK(cudaMemcpy2DAsync(dA.p, sizeof(double) * ldA, A, sizeof(double) * lda, sizeof(double) * n, n, cudaMemcpyHostToDevice, st));
        dense_mirror_lower_kernel<<<dim3((unsigned)n, (unsigned)((n + 255) / 256)), 256, 0, st>>>(dA.p, ldA, n);
        CUDA_CHECK(cudaGetLastError());
        auto gemm_sub = [&](const double* Aop, long lda_, const double* Bop, long ldb_, double* C, long ldc_, int M, int N, int K) {
            if (M <= 0 || N <= 0) return;
            GemmProblem P;
            memset(&P, 0, sizeof P);
            P.A = Aop; P.B = Bop; P.C = C; P.colidx = iota.p; P.M = M; P.N = N; P.K = K; P.lda = lda_; P.ldb = ldb_; P.ldc = ldc_; P.c_sub = 1;
            dense_gemm_work_kernel<<<8, 256, 0, st>>>(P, probs.p, tiles.p, ntl.p);
            CUDA_CHECK(cudaGetLastError());
            const long nt = (long)((M + 127) / 128) * ((N + 127) / 128);
            launch_gemm<128, 128, 16, 2, 4, 3>(st, probs.p, tiles.p, ntl.p, std::min<long>(nt, 148L * 2));
        };
        // ---- tridiagonalisation ---------------------------------------------------------------------------------
        CUDA_CHECK(cudaEventRecord(ev[0], st));
        for (int j0 = 0; j0 < n; j0 += DN_NB) {
            const int nb = std::min((int)DN_NB, n - j0);
            dev_zero(Vp.p, sizeof(double) * (size_t)ldp * DN_NB, st);
            dev_zero(Wp.p, sizeof(double) * (size_t)ldp * DN_NB, st);
            CUDA_CHECK(cudaMemcpy2DAsync(PW.p, sizeof(double) * ldp, dA.p + (long)j0 * ldA, sizeof(double) * ldA, sizeof(double) * n, nb,
                                         cudaMemcpyDeviceToDevice, st));
            for (int c = 0; c < nb; ++c) {
                const int i = j0 + c;
                dense_house_kernel<<<1, DN_THREADS, 0, st>>>(dA.p, ldA, n, i, c, PW.p, Vp.p, ldp, dd.p, de.p, dtau.p);
                if (i >= n - 1) break;
                const int m = n - (i + 1), rb = (m + 255) / 256;
                const int rb2 = (n - ((i + 1) & ~1) + 511) / 512;                     // symv: two rows per thread
                const int nsplit = std::max(1, std::min({(int)DN_SPLIT, (4 * 148 + rb2 - 1) / rb2, (m + 63) / 64}));
                dense_symv_kernel<<<dim3((unsigned)std::max(rb2, c), (unsigned)(nsplit + 1)), 256, 0, st>>>(dA.p, ldA, n, i, c, Vp.p, Wp.p, ldp, part.p, dots.p);
                dense_w_kernel<<<rb, 256, 0, st>>>(n, i, c, nsplit, part.p, dots.p, Vp.p, Wp.p, ldp, dtau.p, wtmp.p, wpart.p);
                dense_panel_update_kernel<<<dim3((unsigned)rb, (unsigned)(nb - c)), 256, 0, st>>>(n, i, c, rb, dtau.p, wtmp.p, wpart.p, Vp.p, Wp.p, PW.p, ldp);
            }
            CUDA_CHECK(cudaGetLastError());
            const int t0 = j0 + nb;
            if (t0 < n) {
                // trailing block (both triangles): A22 -= V W^T + W V^T ; the row-major K x N operand of the GEMM is the
                // column-major panel of the other vector set
                double* C = dA.p + (long)t0 * ldA + t0;
                gemm_sub(Vp.p + t0, ldp, Wp.p + t0, ldp, C, ldA, n - t0, n - t0, DN_NB);
                gemm_sub(Wp.p + t0, ldp, Vp.p + t0, ldp, C, ldA, n - t0, n - t0, DN_NB);
            }
        }
        CUDA_CHECK(cudaEventRecord(ev[1], st));
        std::vector<double> hd(n), he(std::max(1, n - 1));
        dev_d2h(hd.data(), dd.p, sizeof(double) * n, st);
        if (n > 1) dev_d2h(he.data(), de.p, sizeof(double) * (n - 1), st);
        dev_sync(st);
        // ---- the tridiagonal path (accurate rule on every level) -----------------------------------------------------
        const double t_s0 = wall_now();
        int rc = create_common(&hs, n, 1, Z ? CUPPEN_FLAG_VECTORS | CUPPEN_FLAG_NO_RESIDUALS : 0, device, Comm());
        if (rc != 0) CUPPEN_THROW(rc, "%s", g_last_error.c_str());
        hs->s.set_matrix(hd.data(), he.data());
        hs->s.solve();
        memcpy(W, hs->s.h_lam_sorted.data(), sizeof(double) * n);
        const double t_s1 = wall_now();
        float ms_back = 0;
        if (Z) {
            // ---- back-transformation Z = H_0 ... H_{n-2} V, one block reflector per panel, last panel first ----------------
            hs->s.materialise_sorted();
            double* Zd = hs->s.Qcur;
            const long ldz_d = hs->s.ldq;
            CUDA_CHECK(cudaEventRecord(ev[2], st));
            const int last = ((n - 1) / DN_NB) * DN_NB;
            for (int j0 = last; j0 >= 0; j0 -= DN_NB) {
                const int nb = std::min((int)DN_NB, n - j0), row0 = j0 + 1;
                if (row0 >= n) continue;
                dense_extract_v_kernel<<<dim3((unsigned)((n + 255) / 256), DN_NB), 256, 0, st>>>(dA.p, ldA, n, j0, nb, Vp.p, ldp);
                const int gparts = (n - row0 + DN_GRAM_ROWS - 1) / DN_GRAM_ROWS;
                dense_gram_kernel<<<gparts, 256, 0, st>>>(n, row0, Vp.p, ldp, Gp.p);
                dense_larft_kernel<<<1, DN_THREADS, 0, st>>>(j0, nb, gparts, Gp.p, dtau.p, Tm.p);
                dense_vt_kernel<<<(unsigned)((n - row0 + 255) / 256), 256, 0, st>>>(n, row0, nb, Vp.p, ldp, Tm.p, VT.p);
                dense_vtz_kernel<<<(unsigned)((n + 63) / 64), 256, 0, st>>>(n, row0, n, Vp.p, ldp, Zd, ldz_d, W1.p, ldp);
                CUDA_CHECK(cudaGetLastError());
                gemm_sub(VT.p + row0, ldp, W1.p, ldp, Zd + row0, ldz_d, n - row0, n, DN_NB);
            }
            CUDA_CHECK(cudaEventRecord(ev[3], st));
            if (ldz == ldz_d) CUDA_CHECK(cudaMemcpyAsync(Z, Zd, sizeof(double) * (size_t)ldz * n, cudaMemcpyDeviceToHost, st));
            else CUDA_CHECK(cudaMemcpy2DAsync(Z, sizeof(double) * ldz, Zd, sizeof(double) * ldz_d, sizeof(double) * n, n, cudaMemcpyDeviceToHost, st));
            dev_sync(st);
            cudaEventElapsedTime(&ms_back, ev[2], ev[3]);
        }
        float ms_tri = 0;
        cudaEventElapsedTime(&ms_tri, ev[0], ev[1]);
        if (tm) {
            tm->tridiagonalise_s = ms_tri * 1e-3; tm->tridiagonal_solve_s = t_s1 - t_s0; tm->backtransform_s = ms_back * 1e-3;
            tm->tridiagonal_device_s = hs->s.timers.device_s;
        }
        cuppen_destroy(hs);
        hs = nullptr;
    } catch (...) {
        if (hs) cuppen_destroy(hs);
        for (auto e : ev) if (e) cudaEventDestroy(e);
        cudaStreamDestroy(st);
        throw;
    }
    for (auto e : ev) cudaEventDestroy(e);
    cudaStreamDestroy(st);
#else
    (void)Z; (void)ldz; (void)device; (void)tm;
    CUPPEN_THROW(CUPPEN_ERR_CUDA, "the dense front end has no host build");
#endif
    CUPPEN_API_END
}

const char* cuppen_last_error(void) { return g_last_error.c_str(); }

}  // extern "C"
