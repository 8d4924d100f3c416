// K6 (TMA version): the same contraction as gemm_dmma.h, with the operand tiles staged by the
// Tensor Memory Accelerator (cp.async.bulk.tensor.2d -> SASS UTMALDG) into 128B-swizzled shared
// memory behind a ring of mbarriers: one producer warp issues the bulk tensor copies, eight
// consumer warps issue mma.sync.m8n8k4.f64 (DMMA.8x8x4).  TMA takes arbitrary tile coordinates
// inside the Apack / U-arena tensors, so odd row offsets of a half (n1 odd) need no special case.
//
// Shared-memory layout of one stage: A tile 128(m) x 16(k) as 8 boxes of [16 k-rows][16 m] doubles
// (row = 128 B, 16-byte chunks XOR-swizzled with the row index), then the B tile 128(n) x 16(k)
// likewise.  A fragment load of mma.m8n8k4 touches (m = lane>>2, k = lane&3); assigning
// k = 2*(lane&3) + (step&1) + 8*(step>>1) makes the 16 lanes of a half-warp hit 16 different
// 8-byte bank pairs under the 128B swizzle (conflict-free), and since A and B use the same k
// permutation the sum over k is unchanged.
#ifndef CUPPEN_GEMM_TMA_H
#define CUPPEN_GEMM_TMA_H

#include "gemm_dmma.h"

#if CUPPEN_CUDA
#include <cuda.h>

namespace cuppen {

enum { TMA_BM = 128, TMA_BN = 128, TMA_BK = 16, TMA_STAGES = 4, TMA_CONSUMER_WARPS = 8 };
// two consumer warpgroups + one producer warpgroup; setmaxnreg moves the producer's registers to the consumers
enum { TMA_THREADS = (TMA_CONSUMER_WARPS + 4) * 32, TMA_REGS_CONSUMER = 232, TMA_REGS_PRODUCER = 40 };
enum { TMA_BOX_BYTES = 16 * 16 * 8, TMA_TILE_BYTES = 8 * TMA_BOX_BYTES, TMA_STAGE_BYTES = 2 * TMA_TILE_BYTES };

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1ull << 26)) __trap();      // a lost arrival must fail, not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
        :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const GemmProblem* __restrict__ probs, const GemmTile* __restrict__ tiles, int ntiles) {
    extern __shared__ uint8_t tma_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)tma_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + TMA_STAGES * TMA_STAGE_BYTES);
    uint64_t* empty = full + TMA_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TMA_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= TMA_CONSUMER_WARPS) {
        // ===== producer warpgroup: one lane of its first warp issues the bulk tensor copies =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" :: "n"((int)TMA_REGS_PRODUCER));
        if (warp == TMA_CONSUMER_WARPS && lane == 0) {
            int stage = 0;
            unsigned phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const GemmTile T = tiles[tile];
                const GemmProblem& P = probs[T.prob];
                const int ktiles = (P.K + TMA_BK - 1) / TMA_BK;
                const int ar = P.a_row0 + T.m0, ac = P.a_col0, br = P.b_row0, bc = P.b_col0 + T.n0;
                for (int kt = 0; kt < ktiles; ++kt) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], TMA_STAGE_BYTES);
                    uint8_t* sA = smem + stage * TMA_STAGE_BYTES;
                    uint8_t* sB = sA + TMA_TILE_BYTES;
#pragma unroll
                    for (int b = 0; b < 8; ++b) tma_load_2d(sA + b * TMA_BOX_BYTES, &mapA, ar + 16 * b, ac + kt * TMA_BK, &full[stage]);
#pragma unroll
                    for (int b = 0; b < 8; ++b) tma_load_2d(sB + b * TMA_BOX_BYTES, &mapB, bc + 16 * b, br + kt * TMA_BK, &full[stage]);
                    if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumer warps: 2 (m) x 4 (n), warp tile 64 x 32 =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" :: "n"((int)TMA_REGS_CONSUMER));
    const int wm = warp & 1, wn = warp >> 1;
    const int lr = lane >> 2, lk = lane & 3;
    // byte offsets of this lane's fragment elements inside a tile, for k-step parity 0 (even k) --
    // the odd step flips chunk bit 0 (xor 16 bytes) and adds one k row (128 bytes)
    unsigned offA[8], offB[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int box = wm * 4 + (i >> 1), c = ((i & 1) * 4 + (lr >> 1)) ^ (2 * lk);
        offA[i] = box * TMA_BOX_BYTES + (2 * lk) * 128 + c * 16 + (lr & 1) * 8;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int box = wn * 2 + (j >> 1), c = ((j & 1) * 4 + (lr >> 1)) ^ (2 * lk);
        offB[j] = box * TMA_BOX_BYTES + (2 * lk) * 128 + c * 16 + (lr & 1) * 8;
    }

    int stage = 0;
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const GemmTile T = tiles[tile];
        const GemmProblem P = probs[T.prob];
        const int ktiles = (P.K + TMA_BK - 1) / TMA_BK;
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(&full[stage], phase);
            const uint8_t* sA = smem + stage * TMA_STAGE_BYTES;
            const uint8_t* sB = sA + TMA_TILE_BYTES;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                // k = 2*lk + (s&1) + 8*(s>>1): row offset (s&1)*128 + (s>>1)*1024, chunk bit 0 flipped for odd s
                const unsigned kofs = (s & 1) * 128 + (s >> 1) * 1024, flip = (s & 1) * 16;
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = *(const double*)(sA + ((offA[i] ^ flip) + kofs));
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = *(const double*)(sB + ((offB[j] ^ flip) + kofs));
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
                    if (mm < P.M) ccol[mm] = acc[i][j][h];
                }
            }
        }
    }
}

// ---- host side: tensor maps -----------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tma_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) CUPPEN_THROW(-10, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// fp64 matrix with `inner` contiguous elements per line, `outer` lines, line stride `ld` elements;
// boxes of 16 x 16 elements, 128-byte swizzle, zero fill outside
inline CUtensorMap make_tma_map(const double* base, uint64_t inner, uint64_t outer, uint64_t ld) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(double)};
    cuuint32_t box[2] = {16, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = tma_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CUPPEN_THROW(-10, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return m;
}

inline size_t tma_smem_bytes() { return (size_t)TMA_STAGES * TMA_STAGE_BYTES + 2 * TMA_STAGES * sizeof(uint64_t) + 1024; }

inline void launch_gemm_tma(Stream s, const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmProblem* probs,
                            const GemmTile* tiles, int ntiles, int max_ctas) {
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes()));
        attr_set = true;
    }
    int grid = ntiles < max_ctas ? ntiles : max_ctas;
    dgemm_tma_kernel<<<grid, TMA_THREADS, tma_smem_bytes(), s>>>(mapA, mapB, probs, tiles, ntiles);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace cuppen
#endif  // CUPPEN_CUDA
#endif
