// K6: FP64 tensor-core GEMM for the back-transformation  Q'[rows of a half, live columns] = A * B
//   A  packed live columns of diag(Q1,Q2) after the inverse Givens rotations, column-major (M x K)
//   B  eigenvector matrix of the rank-one update, row-major (K x N)   (ugen_kernel)
//   C  written column-scattered into the parent's Q (column-major): column n -> colidx[n]
// This is the dense contraction that the reference evaluates lazily, one row and one eigenvector
// at a time, in writeResults (/root/reference/src/filehandling.c:443-494) and for the two
// boundary rows in src/main.c:613-639.
//
// sm_100a has no f64 tcgen05/wgmma: the FP64 tensor instruction is mma.sync.m8n8k4.f64
// (SASS DMMA.8x8x4).  One launch covers a whole list of (problem, tile) work items so that all
// merges of a tree level share a launch.
#ifndef CUPPEN_GEMM_DMMA_H
#define CUPPEN_GEMM_DMMA_H

#include "platform.h"

namespace cuppen {

struct GemmProblem {
    const double* A;      // M x K, column-major, lda
    const double* B;      // K x N, row-major, ldb
    double* C;            // base of the output rows; column n lives at C + colidx[n]*ldc
    const int* colidx;    // N entries
    int M, N, K;
    long lda, ldb, ldc;
    // coordinates of the same operands inside the Apack / U-arena tensors (TMA version)
    int a_row0, a_col0, b_row0, b_col0;
    int c_sub;            // 0: C = A B (the merges), 1: C -= A B (rank-2k / block-reflector updates of the dense front end; cp.async kernels)
};

struct GemmTile { int prob, m0, n0; };
// `prob` carries a flag: a HALF tile covers only the 64 columns [n0, n0 + 64) -- the tiles of an under-filled last wave are
// emitted as two halves each (work list builder, matrix_stages.h; tensor-map TMA kernel only)
enum { GEMM_TILE_HALF = 0x40000000, GEMM_TILE_PROB_MASK = 0x3fffffff };

#if CUPPEN_CUDA

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// FP64 yardsticks: register-resident issue loops (no memory traffic)
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double acc[16][2];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = 0.0;
    const double a = 1e-9 * threadIdx.x, b = 1.0 + 1e-9 * blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dmma884(acc[j][0], acc[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j][0] + acc[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 1e-3 * j;
    const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-9 * blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma(acc[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA issued from the same SM at the same time (warps 0-3 DMMA, warps 4-7 DFMA): if the two
// instruction kinds share the FP64 units, the mixed rates add up to one peak -- the evidence behind the decision
// not to generate U inside the GEMM producer.
__global__ void __launch_bounds__(256) fp64_mix_kernel(double* out, int iters_dmma, int iters_dfma) {
    const bool is_dmma = (threadIdx.x >> 5) < 4;
    double s = 0;
    if (is_dmma) {
        double acc[16][2];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = 0.0;
        const double a = 1e-9 * threadIdx.x, b = 1.0 + 1e-9 * blockIdx.x;
        for (int it = 0; it < iters_dmma; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma884(acc[j][0], acc[j][1], a, b);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) s += acc[j][0] + acc[j][1];
    } else {
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 1e-3 * j;
        const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-9 * blockIdx.x;
        for (int it = 0; it < iters_dfma; ++it) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fma(acc[j], a, b);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) s += acc[j];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// CTA tile BM x BN x BK, WARPS_M x WARPS_N warps, each warp (BM/WARPS_M) x (BN/WARPS_N).
template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
struct DmmaCfg {
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    static constexpr int MI = WM / 8, NI = WN / 8;
    static constexpr int LDA_S = BM + 4, LDB_S = BN + 4;         // +4 doubles: conflict-free 64-bit fragment loads
    static constexpr int STAGE_DOUBLES = BK * (LDA_S + LDB_S);
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
};

template <class Cfg, int BM, int BN, int BK, int STAGES>
__device__ __forceinline__ void gemm_load_stage(double* sA, double* sB, const GemmProblem& P, int m0, int n0, int k0) {
    // A tile: BK rows (k) of BM contiguous m; out-of-range k rows are not needed (A is zero padded to K_PAD)
    constexpr int A_ELEMS = BK * BM;
    for (int idx = threadIdx.x; idx < A_ELEMS; idx += Cfg::THREADS) {
        int kk = idx / BM, mm = idx - kk * BM;
        cp_async8(sA + kk * Cfg::LDA_S + mm, P.A + (long)(k0 + kk) * P.lda + (m0 + mm));
    }
    constexpr int B_ELEMS = BK * BN;
    for (int idx = threadIdx.x; idx < B_ELEMS; idx += Cfg::THREADS) {
        int kk = idx / BN, nn = idx - kk * BN;
        cp_async8(sB + kk * Cfg::LDB_S + nn, P.B + (long)(k0 + kk) * P.ldb + (n0 + nn));
    }
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32)
dgemm_dmma_kernel(const GemmProblem* __restrict__ probs, const GemmTile* __restrict__ tiles, const int* __restrict__ ntiles_ptr) {
    const int ntiles = *ntiles_ptr;
    using Cfg = DmmaCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
    extern __shared__ __align__(16) double gemm_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lr = lane >> 2, lk = lane & 3;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const GemmTile T = tiles[tile];
        const GemmProblem P = probs[T.prob & GEMM_TILE_PROB_MASK];
        const int m0 = T.m0, n0 = T.n0;
        const int ktiles = (P.K + BK - 1) / BK;

        double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
            for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        __syncthreads();      // previous tile's smem reads are done
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < ktiles) {
                double* sA = gemm_smem + s * Cfg::STAGE_DOUBLES;
                gemm_load_stage<Cfg, BM, BN, BK, STAGES>(sA, sA + BK * Cfg::LDA_S, P, m0, n0, s * BK);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < ktiles; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            {   // prefetch tile kt+STAGES-1 into the slot that was consumed in iteration kt-1
                const int nk = kt + STAGES - 1;
                if (nk < ktiles) {
                    double* sA = gemm_smem + (nk % STAGES) * Cfg::STAGE_DOUBLES;
                    gemm_load_stage<Cfg, BM, BN, BK, STAGES>(sA, sA + BK * Cfg::LDA_S, P, m0, n0, nk * BK);
                }
                cp_async_commit();
            }
            const double* sA = gemm_smem + (kt % STAGES) * Cfg::STAGE_DOUBLES;
            const double* sB = sA + BK * Cfg::LDA_S;
#pragma unroll
            for (int k4 = 0; k4 < BK / 4; ++k4) {
                double af[Cfg::MI], bf[Cfg::NI];
                const double* pa = sA + (k4 * 4 + lk) * Cfg::LDA_S + wm * Cfg::WM + lr;
                const double* pb = sB + (k4 * 4 + lk) * Cfg::LDB_S + wn * Cfg::WN + lr;
#pragma unroll
                for (int i = 0; i < Cfg::MI; ++i) af[i] = pa[i * 8];
#pragma unroll
                for (int j = 0; j < Cfg::NI; ++j) bf[j] = pb[j * 8];
#pragma unroll
                for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
                    for (int j = 0; j < Cfg::NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        cp_async_wait<0>();
        // epilogue: thread holds C[m = lr][n = 2*lk + {0,1}] of every 8x8 sub-tile
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int nn = n0 + wn * Cfg::WN + j * 8 + 2 * lk + h;
                if (nn >= P.N) continue;
                double* ccol = P.C + (long)P.colidx[nn] * P.ldc;
#pragma unroll
                for (int i = 0; i < Cfg::MI; ++i) {
                    const int mm = m0 + wm * Cfg::WM + i * 8 + lr;
                    if (mm < P.M) ccol[mm] = P.c_sub ? ccol[mm] - acc[i][j][h] : acc[i][j][h];
                }
            }
        }
    }
}
#endif  // CUPPEN_CUDA

}  // namespace cuppen
#endif
