// On-GPU orthogonality check  max_ij |(V^T V - I)_ij|  of the computed eigenvector matrix
// (BASELINE's criterion "orthogonality ||V^T V - I|| at or below the reference's"; the reference has
// no such check and cannot even emit V, SURVEY.md finding 6).
//
// V is column-major (rows x n, leading dimension ld), so both operands of the Gram product are
// K-contiguous ("TN" GEMM) -- a different shared-memory layout from the back-transformation GEMMs:
//   stage = A tile [128 columns i][16 k + 4 pad] and B tile [128 columns j][16 k + 4 pad], filled with
//   16-byte cp.async (zero-filled beyond the last row / column), 4-stage ring;
//   fragments of mma.sync.m8n8k4.f64 (DMMA.8x8x4) read as s[(col)*20 + k]: the row stride of 20 doubles
//   makes the 64-bit fragment loads conflict-free.
// Only the tiles on or above the diagonal are computed (the Gram matrix is symmetric); the product
// is never written to memory: the epilogue subtracts the identity and reduces max|.| into one word.
#ifndef CUPPEN_ORTH_CHECK_H
#define CUPPEN_ORTH_CHECK_H

#include "gemm_dmma.h"

namespace cuppen {

enum { GR_BT = 128, GR_BK = 16, GR_LDK = GR_BK + 4, GR_STAGES = 4, GR_THREADS = 256 };
enum { GR_TILE_DOUBLES = GR_BT * GR_LDK, GR_STAGE_DOUBLES = 2 * GR_TILE_DOUBLES };

// gathered row slices [rank][column][ldq] -> one column-major matrix with G*ldq rows per column (rank r's rows at
// r*ldq, zero beyond its row count: V^T V does not care about the order of the rows)
struct GramGather {
    const double* gath;
    double* V;
    long ldq;
    int n, G;
    int nloc[8];
    CUPPEN_HD void operator()(long t) const {
        const long per_col = (long)G * ldq;
        const long col = t / per_col, rem = t - col * per_col;
        const int r = (int)(rem / ldq), i = (int)(rem - (long)r * ldq);
        V[t] = (i < nloc[r]) ? gath[((long)r * n + col) * ldq + i] : 0.0;
    }
};

#if CUPPEN_CUDA
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(src_bytes));
}

// linear index over the upper triangle (ti <= tj) of a T x T tile grid, row by row
__device__ __forceinline__ void gram_tile_coords(long t, int T, int& ti, int& tj) {
    // row ti starts at ti*T - ti*(ti-1)/2
    double tt = (double)T + 0.5;
    int r = (int)(tt - sqrt(tt * tt - 2.0 * (double)t));
    if (r < 0) r = 0;
    if (r > T - 1) r = T - 1;
    while (r > 0 && (long)r * T - (long)r * (r - 1) / 2 > t) --r;
    while ((long)(r + 1) * T - (long)(r + 1) * r / 2 <= t) ++r;
    ti = r;
    tj = r + (int)(t - ((long)r * T - (long)r * (r - 1) / 2));
}

__device__ __forceinline__ void gram_load_stage(double* stage, const double* __restrict__ V, long ld, int rows, int n,
                                                int i0, int j0, int k0) {
    // 2 tiles x 128 columns x 8 chunks of 16 bytes = 2048 chunks, 8 per thread
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int idx = threadIdx.x + u * GR_THREADS;
        const int which = idx >> 10;                    // 0: A tile, 1: B tile
        const int cc = (idx >> 3) & 127;                // column inside the tile
        const int ch = idx & 7;                         // 16-byte chunk (2 doubles) along k
        const int col = (which ? j0 : i0) + cc;
        const int k = k0 + 2 * ch;
        int bytes = (rows - k) * 8;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        if (col >= n) bytes = 0;
        const double* src = V + (long)(col < n ? col : 0) * ld + (bytes > 0 ? k : 0);
        cp_async16_zfill(stage + which * GR_TILE_DOUBLES + cc * GR_LDK + 2 * ch, src, bytes);
    }
}

__global__ void __launch_bounds__(GR_THREADS, 1)
gram_check_kernel(const double* __restrict__ V, long ld, int rows, int n, unsigned long long* __restrict__ result, int part, int nparts) {
    extern __shared__ __align__(16) double gram_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1;           // 2 (i) x 4 (j) warps, warp tile 64 x 32
    const int lr = lane >> 2, lk = lane & 3;
    const int T = (n + GR_BT - 1) / GR_BT;
    const long ntiles = (long)T * (T + 1) / 2;
    const int ktiles = (rows + GR_BK - 1) / GR_BK;
    double worst = 0.0;

    // several GPUs: every rank holds the gathered V and takes the tiles part, part + nparts, ... of the triangle
    for (long tile = (long)blockIdx.x * nparts + part; tile < ntiles; tile += (long)gridDim.x * nparts) {
        int ti, tj;
        gram_tile_coords(tile, T, ti, tj);
        const int i0 = ti * GR_BT, j0 = tj * GR_BT;
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        __syncthreads();
#pragma unroll
        for (int s = 0; s < GR_STAGES - 1; ++s) {
            if (s < ktiles) gram_load_stage(gram_smem + s * GR_STAGE_DOUBLES, V, ld, rows, n, i0, j0, s * GR_BK);
            cp_async_commit();
        }
        for (int kt = 0; kt < ktiles; ++kt) {
            cp_async_wait<GR_STAGES - 2>();
            __syncthreads();
            {
                const int nk = kt + GR_STAGES - 1;
                if (nk < ktiles) gram_load_stage(gram_smem + (nk % GR_STAGES) * GR_STAGE_DOUBLES, V, ld, rows, n, i0, j0, nk * GR_BK);
                cp_async_commit();
            }
            const double* sA = gram_smem + (kt % GR_STAGES) * GR_STAGE_DOUBLES + (wm * 64 + lr) * GR_LDK + lk;
            const double* sB = gram_smem + (kt % GR_STAGES) * GR_STAGE_DOUBLES + GR_TILE_DOUBLES + (wn * 32 + lr) * GR_LDK + lk;
#pragma unroll
            for (int k4 = 0; k4 < GR_BK / 4; ++k4) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = sA[i * 8 * GR_LDK + k4 * 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = sB[j * 8 * GR_LDK + k4 * 4];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        cp_async_wait<0>();
        // epilogue: thread holds G[i = lr][j = 2*lk + {0,1}] of every 8x8 sub-tile
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gi = i0 + wm * 64 + i * 8 + lr;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gj = j0 + wn * 32 + j * 8 + 2 * lk + h;
                    if (gi < n && gj < n) {
                        double dev = fabs(acc[i][j][h] - (gi == gj ? 1.0 : 0.0));
                        if (!(dev <= 1.7e308)) dev = 1.7e308;         // NaN / inf must not hide in the max
                        worst = fmax(worst, dev);
                    }
                }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane == 0 && worst > 0.0) atomicMax(result, (unsigned long long)__double_as_longlong(worst));
}

inline size_t gram_smem_bytes() { return (size_t)GR_STAGES * GR_STAGE_DOUBLES * sizeof(double); }
#else
// TEST-ONLY host twin (CUPPEN_HOST_EMULATION)
inline double gram_check_host(const double* V, long ld, int rows, int n) {
    double worst = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < rows; ++k) s += V[(long)i * ld + k] * V[(long)j * ld + k];
            double dev = fabs(s - (i == j ? 1.0 : 0.0));
            if (!(dev <= 1.7e308)) dev = 1.7e308;
            if (dev > worst) worst = dev;
        }
    return worst;
}
#endif

}  // namespace cuppen
#endif
