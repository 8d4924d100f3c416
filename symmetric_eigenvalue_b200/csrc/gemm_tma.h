// K6 (TMA version): the same contraction as gemm_dmma.h with the operand tiles staged by the TMA engine into a
// 4-stage shared-memory ring guarded by full/empty mbarriers; eight consumer warps issue mma.sync.m8n8k4.f64
// (DMMA.8x8x4).  setmaxnreg moves the producer warpgroup's registers to the consumers (128 accumulator registers per
// thread).  Two producers (template parameter):
//   TENSOR = true   (default) one tensor-map copy per operand and stage (cp.async.bulk.tensor.2d -> SASS UTMALDG), issued
//                   by one lane; the box {132, 16} is as wide as the padded shared-memory line, so the tile lands in
//                   the conflict-free layout directly and tails beyond the buffer are zero-filled;
//   TENSOR = false  32 bulk-copy lines per stage (cp.async.bulk.shared::cluster.global -> SASS UBLKCP), one 1 KiB line
//                   per k of the A tile and of the B tile, issued by the 32 lanes of the producer warp.
// Both need a 16-byte aligned source: tests/probe/tma_probe.cu shows that an fp64 tensor copy works with an EVEN start
// coordinate and raises "illegal instruction" with an odd one (round 1 had only tried odd starts and concluded that
// tensor maps were unusable; profiles/r02_tma_probe.txt).  Operands whose first row is odd (odd reference leaf sizes,
// odd slice offsets) are therefore fetched from one row earlier and read one element further (`ashift`).
//
// Shared-memory layout of one stage: A tile [16 k][128 m + 4 pad] doubles, then B tile
// [16 k][128 n + 4 pad]; the pad of 4 doubles makes the 64-bit fragment loads of mma.m8n8k4
// (m = lane>>2, k = lane&3) conflict-free.
#ifndef CUPPEN_GEMM_TMA_H
#define CUPPEN_GEMM_TMA_H

#include "gemm_dmma.h"

#if CUPPEN_CUDA
#include <cuda.h>

namespace cuppen {

enum { TMA_BM = 128, TMA_BN = 128, TMA_BK = 16, TMA_STAGES = 4, TMA_CONSUMER_WARPS = 8 };
enum { TMA_LD = TMA_BM + 4, TMA_LINE_BYTES = TMA_BM * 8, TMA_TILE_DOUBLES = TMA_BK * TMA_LD,
       TMA_STAGE_DOUBLES = 2 * TMA_TILE_DOUBLES, TMA_STAGE_TX_BYTES = 2 * TMA_BK * TMA_LINE_BYTES };
// two consumer warpgroups + one producer warpgroup
enum { TMA_THREADS = (TMA_CONSUMER_WARPS + 4) * 32, TMA_REGS_CONSUMER = 232, TMA_REGS_PRODUCER = 40 };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
// Pipeline watchdog: a lost arrival must fail, not hang the GPU.  ab[0] != 0 makes every wait return at
// once; [1..6] record who gave up first (role, stage, parity, tile, k-tile, cta).  `ab` is the caller's flag block
// (8 ints of the solver's `fail` buffer, read back with the results of the solve).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity, int role, int stage, int tile, int kt, int* ab) {
    unsigned done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && (++spins & 0x3ff) == 0) {
            if (*(volatile int*)&ab[0]) return false;
            if (spins > (1u << 22)) {
                if (atomicCAS(&ab[0], 0, 1) == 0) {
                    ab[1] = role; ab[2] = stage; ab[3] = (int)parity;
                    ab[4] = tile; ab[5] = kt; ab[6] = blockIdx.x;
                }
                return false;
            }
        }
    }
    return true;
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP)
__device__ __forceinline__ void tma_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA tensor copy of one {TMA_LD x TMA_BK} box (UTMALDG): c_in = element index along the contiguous dimension
// (must be EVEN for fp64 maps -- a 16-byte aligned start; odd coordinates raise "illegal instruction" on this pool,
// tests/probe/tma_probe.cu, profiles/r02_tma_probe.txt), c_out = line index
__device__ __forceinline__ void tma_tensor_load(void* dst, const CUtensorMap* map, int c_in, int c_out, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c_in), "r"(c_out), "r"(smem_u32(bar)) : "memory");
}

// The same copy with an L2 eviction-priority hint (createpolicy): in the super-column tile order the B panel is re-used by
// every m-tile of the super-column, so its lines are loaded `evict_last`; the C tile is written with streaming stores so
// that 0.9 GB of output per half does not push the panel out of the L2 (`hints` kernel argument; large merges only --
// on small ones the output SHOULD stay in L2 for the next kernel).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_tensor_load_hint(void* dst, const CUtensorMap* map, int c_in, int c_out, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c_in), "r"(c_out), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// Operand tiles whose first row is odd (odd reference leaf sizes, odd slice offsets): a TMA copy needs a 16-byte
// aligned source, so the A lines are fetched from one row earlier (`ashift` = 1: 130 doubles per line instead of 128;
// the 4 pad doubles of the shared-memory line absorb them) and the consumers read their fragments one element further.
// TENSOR = false: 32 bulk-copy lines per stage issued by the 32 lanes of the producer warp (UBLKCP);
// TENSOR = true:  two tensor copies per stage issued by one lane (UTMALDG), box {TMA_LD, TMA_BK}: the box is as wide
//                 as the padded shared-memory line, so the tile lands in the same conflict-free layout.
template <bool TENSOR>
__global__ void __launch_bounds__(TMA_THREADS, 1)
dgemm_tma_kernel(const GemmProblem* __restrict__ probs, const GemmTile* __restrict__ tiles, const int* __restrict__ ntiles_ptr, int* abort_flags,
                 const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int hints) {
    extern __shared__ __align__(128) double tma_smem[];
    uint64_t* full = (uint64_t*)(tma_smem + TMA_STAGES * TMA_STAGE_DOUBLES);
    uint64_t* empty = full + TMA_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = *ntiles_ptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TMA_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= TMA_CONSUMER_WARPS) {
        // ===== producer warpgroup: its first warp issues the copies =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" :: "n"((int)TMA_REGS_PRODUCER));
        if (warp != TMA_CONSUMER_WARPS) return;
        int stage = 0;
        unsigned phase = 0;
        const bool isA = lane < TMA_BK;
        const int kk = lane & (TMA_BK - 1);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const GemmTile T = tiles[tile];
            const GemmProblem& P = probs[T.prob & GEMM_TILE_PROB_MASK];
            const int ktiles = (P.K + TMA_BK - 1) / TMA_BK;
            const int ashift = P.a_row0 & 1;
            if (TENSOR) {
                if (lane != 0) continue;
                const int ain = P.a_row0 + T.m0 - ashift, bin = P.b_col0 + T.n0;
                if (hints) {
                    // hints & 1: B panel evict_last; hints & 2: A strips evict_first
                    const uint64_t polB = l2_policy_evict_last(), polA = l2_policy_evict_first();
                    for (int kt = 0; kt < ktiles; ++kt) {
                        if (!mbar_wait(&empty[stage], phase ^ 1, 0, stage, tile, kt, abort_flags)) return;
                        mbar_expect_tx(&full[stage], 2 * TMA_TILE_DOUBLES * 8);
                        double* st = tma_smem + stage * TMA_STAGE_DOUBLES;
                        if (hints & 2) tma_tensor_load_hint(st, &mapA, ain, P.a_col0 + kt * TMA_BK, &full[stage], polA);
                        else tma_tensor_load(st, &mapA, ain, P.a_col0 + kt * TMA_BK, &full[stage]);
                        tma_tensor_load_hint(st + TMA_TILE_DOUBLES, &mapB, bin, P.b_row0 + kt * TMA_BK, &full[stage], polB);
                        if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
                    }
                    continue;
                }
                for (int kt = 0; kt < ktiles; ++kt) {
                    if (!mbar_wait(&empty[stage], phase ^ 1, 0, stage, tile, kt, abort_flags)) return;
                    mbar_expect_tx(&full[stage], 2 * TMA_TILE_DOUBLES * 8);
                    double* st = tma_smem + stage * TMA_STAGE_DOUBLES;
                    tma_tensor_load(st, &mapA, ain, P.a_col0 + kt * TMA_BK, &full[stage]);
                    tma_tensor_load(st + TMA_TILE_DOUBLES, &mapB, bin, P.b_row0 + kt * TMA_BK, &full[stage]);
                    if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
                }
                continue;
            }
            // lane < 16: line kk of the A tile; lane >= 16: line kk of the B tile
            const double* src = isA ? P.A + (long)kk * P.lda + T.m0 - ashift : P.B + (long)kk * P.ldb + T.n0;
            const long step = (long)TMA_BK * (isA ? P.lda : P.ldb);
            const int dofs = (isA ? 0 : TMA_TILE_DOUBLES) + kk * TMA_LD;
            const unsigned bytes = TMA_LINE_BYTES + ((isA && ashift) ? 16u : 0u);
            for (int kt = 0; kt < ktiles; ++kt) {
                int ok = 1;
                if (lane == 0) {
                    ok = mbar_wait(&empty[stage], phase ^ 1, 0, stage, tile, kt, abort_flags) ? 1 : 0;
                    if (ok) mbar_expect_tx(&full[stage], TMA_STAGE_TX_BYTES + (ashift ? 16u * TMA_BK : 0u));
                }
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (!ok) return;
                tma_bulk_load(tma_smem + stage * TMA_STAGE_DOUBLES + dofs, src, bytes, &full[stage]);
                src += step;
                if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ===== consumer warps: 2 (m) x 4 (n), warp tile 64 x 32 =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" :: "n"((int)TMA_REGS_CONSUMER));
    const int wm = warp & 1, wn = warp >> 1;
    const int lr = lane >> 2, lk = lane & 3;
    const int aofs0 = lk * TMA_LD + wm * 64 + lr;
    const int bofs = TMA_TILE_DOUBLES + lk * TMA_LD + wn * 32 + lr;

    int stage = 0;
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const GemmTile T = tiles[tile];
        const GemmProblem P = probs[T.prob & GEMM_TILE_PROB_MASK];
        const int ktiles = (P.K + TMA_BK - 1) / TMA_BK;
        const int aofs = aofs0 + (P.a_row0 & 1);
        // half tile (64 columns): the warps of the upper two n-quarters only keep the ring moving; one warp per SM
        // sub-partition still issues its 32 independent DMMAs per k-step back to back, so the tile takes about half the time
        const int nw = (T.prob & GEMM_TILE_HALF) ? 64 : TMA_BN;
        const bool active = wn * 32 < nw;
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < ktiles; ++kt) {
            if (!mbar_wait(&full[stage], phase, 1, stage, tile, kt, abort_flags)) return;
            const double* sA = tma_smem + stage * TMA_STAGE_DOUBLES + aofs;
            const double* sB = tma_smem + stage * TMA_STAGE_DOUBLES + bofs;
            if (active)
#pragma unroll
            for (int k4 = 0; k4 < TMA_BK / 4; ++k4) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = sA[k4 * 4 * TMA_LD + i * 8];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = sB[k4 * 4 * TMA_LD + j * 8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
        }
        // epilogue: C[m = lr][n = 2*lk + {0,1}] of every 8x8 sub-tile, column-scattered
        if (active)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int nn = T.n0 + wn * 32 + j * 8 + 2 * lk + h;
                if (nn >= P.N) continue;
                double* ccol = P.C + (long)P.colidx[nn] * P.ldc;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int mm = T.m0 + wm * 64 + i * 8 + lr;
                    if (mm < P.M) {
                        if (hints) __stcs(&ccol[mm], acc[i][j][h]);      // streaming: the output must not evict the B panel
                        else ccol[mm] = acc[i][j][h];
                    }
                }
            }
        }
    }
}

inline size_t tma_smem_bytes() { return (size_t)TMA_STAGES * TMA_STAGE_DOUBLES * sizeof(double) + 2 * TMA_STAGES * sizeof(uint64_t); }

// Tensor map of a whole operand buffer (fp64, 2-D, box {TMA_LD, TMA_BK}, no swizzle): `inner` contiguous elements per
// line, `lines` lines `ld` elements apart.  The driver entry point is resolved through the runtime (no -lcuda).
inline bool tma_encode_map(CUtensorMap* map, const double* base, long inner, long lines, long ld) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { cudaGetLastError(); return false; }
        enc = (EncodeFn)fn;
    }
    memset(map, 0, sizeof *map);
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)lines};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {TMA_LD, TMA_BK}, es[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// (the >48 KB dynamic shared memory opt-in is a per-device attribute: set_kernel_attributes() in solver.cu)
// mapA / mapB == nullptr: bulk-copy lines; else one tensor copy per operand and stage through the two maps
inline void launch_gemm_tma(Stream s, const GemmProblem* probs, const GemmTile* tiles, const int* ntiles_ptr, int grid, int* abort_flags,
                            const CUtensorMap* mapA = nullptr, const CUtensorMap* mapB = nullptr, int hints = 0) {
    if (grid < 1) return;
    if (mapA && mapB) dgemm_tma_kernel<true><<<grid, TMA_THREADS, tma_smem_bytes(), s>>>(probs, tiles, ntiles_ptr, abort_flags, *mapA, *mapB, hints);
    else {
        static const CUtensorMap none = {};
        dgemm_tma_kernel<false><<<grid, TMA_THREADS, tma_smem_bytes(), s>>>(probs, tiles, ntiles_ptr, abort_flags, none, none, 0);
    }
    CUDA_CHECK(cudaGetLastError());
}

// host copy of the 8 watchdog ints: throws if the pipeline watchdog fired
inline void tma_check_abort(const int* h) {
    if (h[0])
        CUPPEN_THROW(-10, "TMA GEMM pipeline timeout: role=%d (0 producer, 1 consumer) stage=%d parity=%d tile=%d ktile=%d cta=%d",
                     h[1], h[2], h[3], h[4], h[5], h[6]);
}

}  // namespace cuppen
#endif  // CUPPEN_CUDA
#endif
