// Dense symmetric front end: the step BEFORE the tridiagonal path in a real eigen-pipeline --
// Householder tridiagonalisation A = Q T Q^T (LAPACK dsytrd / dlatrd, lower triangle, blocked) and the
// back-transformation Z = Q V of the tridiagonal eigenvectors (dormtr with compact-WY block reflectors, dlarft).  The
// reference has no counterpart (its input is tridiagonal, src/filehandling.c:76-153); the tridiagonal eigenproblem in
// the middle is the path this library is about (ref_leaves = 1: LAPACK-grade tolerances).
//
// Layout: A is kept as a FULL symmetric matrix, column-major, ld = lda (both triangles are updated, so the matrix-
// vector product of every reflector reads whole columns, coalesced).  Per panel of DN_NB columns:
//   per column i:   dense_house_kernel          reflector v_i, tau_i, d_i, e_i from the up-to-date panel workspace column
//                   dense_symv_kernel           partial sums of A22 v over DN_SPLIT column ranges (HBM bound: the n^3/3 * 8 bytes
//                                               that a one-stage reduction must read) + the dots V^T v, W^T v of the panel
//                   dense_w_kernel              w' = tau (A22 v - V W^T v - W V^T v) and the partial sums of w'^T v
//                   dense_panel_update_kernel   w = w' - (tau/2)(w'^T v) v; rank-2 update of the remaining panel columns
//   per panel:      A22 -= [V W] [W V]^T                                  -> the DMMA GEMM of gemm_dmma.h (subtract epilogue)
// Back-transformation, panels in reverse: V^T V (dense_gram_kernel), T (dense_larft_kernel), VT = V T (dense_vt_kernel), W1 = V^T Z
// (dense_vtz_kernel), Z -= VT W1 (DMMA GEMM, subtract epilogue).
#ifndef CUPPEN_DENSE_STAGES_H
#define CUPPEN_DENSE_STAGES_H

#include "gemm_dmma.h"

namespace cuppen {

enum { DN_NB = 64, DN_SPLIT = 64, DN_THREADS = 1024 };

#if CUPPEN_CUDA
__device__ __forceinline__ double dn_block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) s += sh[w];      // fixed order
    return s;
}

// Column i = j0 + c of the panel: the panel workspace PW (n x nb copy of A[:, j0:j0+nb], kept up to date by
// dense_panel_update_kernel) holds the current column; build the reflector that annihilates PW[i+2:n, c] (dlarfg).
// One block.  out: d[i], e[i], tau[i]; v in Vp[:, c] (v[i+1] = 1, zero above) and below the sub-diagonal of A[:, i].
__global__ void __launch_bounds__(DN_THREADS) dense_house_kernel(double* __restrict__ A, long lda, int n, int i, int c,
                                                                 const double* __restrict__ PW, double* __restrict__ Vp, long ldp,
                                                                 double* __restrict__ d, double* __restrict__ e, double* __restrict__ tau) {
    __shared__ double sh[32];
    const double* col = PW + (long)c * ldp;
    double* acol = A + (long)i * lda;
    double* vcol = Vp + (long)c * ldp;
    double ss = 0;                                   // sum of squares of col[i+2:n]
    for (int r = i + 2 + threadIdx.x; r < n; r += blockDim.x) ss = fma(col[r], col[r], ss);
    ss = dn_block_sum(ss, sh);
    if (i >= n - 1) {                                // last column: nothing to annihilate
        if (threadIdx.x == 0) { d[i] = col[i]; tau[i] = 0.0; }
        for (int r = threadIdx.x; r < n; r += blockDim.x) vcol[r] = 0.0;
        return;
    }
    const double alpha = col[i + 1];
    double beta, t_i, scal;
    if (ss == 0.0) { beta = alpha; t_i = 0.0; scal = 0.0; }
    else {
        const double nrm = sqrt(alpha * alpha + ss);
        beta = (alpha >= 0) ? -nrm : nrm;
        t_i = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
    }
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        double v = 0.0;
        if (r == i + 1) v = 1.0;
        else if (r > i + 1) v = col[r] * scal;
        vcol[r] = v;
        if (r > i + 1) acol[r] = v;                  // reflector kept below the sub-diagonal (LAPACK layout)
    }
    if (threadIdx.x == 0) { d[i] = col[i]; e[i] = beta; tau[i] = t_i; acol[i] = col[i]; acol[i + 1] = beta; }
}

// grid (max(row blocks of 512, c), nsplit + 1), nsplit <= DN_SPLIT chosen by the host so that >= 4 blocks per SM exist:
//   blockIdx.y < nsplit    partial sums of y = A[i+1:n, i+1:n] v over a range of columns: part[s][r]   (HBM bound)
//   blockIdx.y == nsplit   block t < c: the dots a_t = V[:,t]^T v and b_t = W[:,t]^T v of the panel correction
__global__ void __launch_bounds__(256) dense_symv_kernel(const double* __restrict__ A, long lda, int n, int i, int c,
                                                         const double* __restrict__ Vp, const double* __restrict__ Wp, long ldp,
                                                         double* __restrict__ part, double* __restrict__ dots) {
    const double* v = Vp + (long)c * ldp;
    __shared__ double sv[256];
    if (blockIdx.y == gridDim.y - 1) {
        const int t = blockIdx.x;
        if (t >= c) return;
        double a = 0, b = 0;
        for (int r = i + 1 + threadIdx.x; r < n; r += 256) {
            const double vr = v[r];
            a = fma(Vp[(long)t * ldp + r], vr, a);
            b = fma(Wp[(long)t * ldp + r], vr, b);
        }
        __shared__ double sh[32];
        a = dn_block_sum(a, sh);
        b = dn_block_sum(b, sh);
        if (threadIdx.x == 0) { dots[t] = a; dots[DN_NB + t] = b; }
        return;
    }
    // two rows per thread (16-byte loads, even global row r), eight columns in flight: enough bytes in flight per SM to
    // approach the HBM roof (one row and two columns per thread left the kernel latency-bound at ~2 TB/s)
    const int rbase = (i + 1) & ~1;
    const int r = rbase + 2 * (blockIdx.x * 256 + threadIdx.x);
    const int m = n - rbase;
    if ((int)blockIdx.x * 512 >= m) return;
    const int nsplit = gridDim.y - 1;
    const int mc = n - (i + 1);
    const int per = (mc + nsplit - 1) / nsplit;
    const int c0 = i + 1 + blockIdx.y * per, c1 = min(n, c0 + per);
    double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
    const bool pair = (r + 1 < n) && ((lda & 1) == 0);
    for (int cb = c0; cb < c1; cb += 256) {
        const int cnt = min(256, c1 - cb);
        __syncthreads();
        if (threadIdx.x < cnt) sv[threadIdx.x] = v[cb + threadIdx.x];
        __syncthreads();
        if (r >= n) continue;
        const double* a = A + (long)cb * lda + r;
        int t = 0;
        if (pair) {
            for (; t + 3 < cnt; t += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double2 x = *reinterpret_cast<const double2*>(a + (long)(t + u) * lda);
                    a0[u] = fma(x.x, sv[t + u], a0[u]);
                    a1[u] = fma(x.y, sv[t + u], a1[u]);
                }
            }
            for (; t < cnt; ++t) {
                const double2 x = *reinterpret_cast<const double2*>(a + (long)t * lda);
                a0[0] = fma(x.x, sv[t], a0[0]);
                a1[0] = fma(x.y, sv[t], a1[0]);
            }
        } else {
            for (; t < cnt; ++t) {
                a0[0] = fma(a[(long)t * lda], sv[t], a0[0]);
                if (r + 1 < n) a1[0] = fma(a[(long)t * lda + 1], sv[t], a1[0]);
            }
        }
    }
    if (r < n && r >= i + 1) part[(long)blockIdx.y * n + r] = (a0[0] + a0[1]) + (a0[2] + a0[3]);
    if (r + 1 < n) part[(long)blockIdx.y * n + r + 1] = (a1[0] + a1[1]) + (a1[2] + a1[3]);
}

// w' = tau (y - V b - W a) for the rows of this block (a_t = V[:,t]^T v, b_t = W[:,t]^T v), stored in wtmp;
// wpart[block] = sum over the block's rows of w' v   (for alpha = -(tau/2) w'^T v)
__global__ void __launch_bounds__(256) dense_w_kernel(int n, int i, int c, int nsplit, const double* __restrict__ part, const double* __restrict__ dots,
                                                      const double* __restrict__ Vp, const double* __restrict__ Wp, long ldp,
                                                      const double* __restrict__ tau, double* __restrict__ wtmp, double* __restrict__ wpart) {
    __shared__ double sh[32];
    __shared__ double sa[DN_NB], sb[DN_NB];
    if (threadIdx.x < c) { sa[threadIdx.x] = dots[threadIdx.x]; sb[threadIdx.x] = dots[DN_NB + threadIdx.x]; }
    __syncthreads();
    const int r = i + 1 + blockIdx.x * 256 + threadIdx.x;
    const double t_i = tau[i];
    double wv = 0;
    if (r < n) {
        double y = 0;
        for (int s = 0; s < nsplit; ++s) y += part[(long)s * n + r];
        for (int t = 0; t < c; ++t) y -= Vp[(long)t * ldp + r] * sb[t] + Wp[(long)t * ldp + r] * sa[t];
        y *= t_i;
        wtmp[r] = y;
        wv = y * Vp[(long)c * ldp + r];
    }
    wv = dn_block_sum(wv, sh);
    if (threadIdx.x == 0) wpart[blockIdx.x] = wv;
}

// grid (row blocks of 256 over rows i+1..n-1, 1 + remaining panel columns):
//   w = w' + alpha v with alpha = -(tau/2) sum(wpart)  (every block re-derives alpha from the partial sums, fixed order)
//   blockIdx.y == 0   stores w into Wp[:, c]
//   blockIdx.y == q   panel column col = i + q (<  j0 + nb): PW[r, col] -= v[r] w[col] + w[r] v[col] for rows r >= col
__global__ void __launch_bounds__(256) dense_panel_update_kernel(int n, int i, int c, int nblocks, const double* __restrict__ tau,
                                                                 const double* __restrict__ wtmp, const double* __restrict__ wpart,
                                                                 const double* __restrict__ Vp, double* __restrict__ Wp, double* __restrict__ PW, long ldp) {
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += wpart[b];
    const double alpha = -0.5 * tau[i] * s;
    const double* v = Vp + (long)c * ldp;
    const int r = i + 1 + blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    const double vr = v[r], wr = fma(alpha, vr, wtmp[r]);
    if (blockIdx.y == 0) { Wp[(long)c * ldp + r] = wr; return; }
    const int col = i + blockIdx.y;                  // global column, panel-local index c + blockIdx.y
    if (r < col) return;
    const double vc = v[col], wc = fma(alpha, vc, wtmp[col]);
    double* p = PW + (long)(c + blockIdx.y) * ldp + r;
    *p -= vr * wc + wr * vc;
}

// Partial Gram matrices of a panel: block b sums V[r, :]^T V[r, :] over its rows r (DN_GRAM_ROWS each) into
// Gp[b][q][t] (64 x 64, row-major).  256 threads, thread (tx, ty) of 16 x 16 owns a 4 x 4 patch; rows go through shared
// memory 32 at a time.  (The first version computed the 2016 dot products of dlarft one after the other inside a single
// block: 6 ms per panel, 0.67 s of back-transformation at n = 8192.)
enum { DN_GRAM_ROWS = 512 };
__global__ void __launch_bounds__(256) dense_gram_kernel(int n, int row0, const double* __restrict__ Vp, long ldp, double* __restrict__ Gp) {
    __shared__ double sV[32][DN_NB + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r_begin = row0 + blockIdx.x * DN_GRAM_ROWS, r_end = min(n, r_begin + DN_GRAM_ROWS);
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int rb = r_begin; rb < r_end; rb += 32) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * DN_NB; idx += 256) {
            const int rr = idx & 31, q = idx >> 5;
            sV[rr][q] = (rb + rr < r_end) ? Vp[(long)q * ldp + rb + rr] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            double va[4], vb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) va[a] = sV[rr][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) vb[b] = sV[rr][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(va[a], vb[b], acc[a][b]);
        }
    }
    double* g = Gp + (size_t)blockIdx.x * DN_NB * DN_NB;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) g[(ty * 4 + a) * DN_NB + tx * 4 + b] = acc[a][b];
}

// T factor of the block reflector H = I - V T V^T of a panel of nb reflectors (dlarft, forward, columnwise):
// T[t,t] = tau_t ; T[0:t, t] = -tau_t T[0:t, 0:t] (V[:, 0:t]^T v_t), with V^T V summed from the partial Gram matrices in a
// fixed order.  One block; T is nb x nb, column-major, ld = DN_NB.
__global__ void __launch_bounds__(DN_THREADS) dense_larft_kernel(int j0, int nb, int nparts, double* __restrict__ Gp,
                                                                 const double* __restrict__ tau, double* __restrict__ T) {
    __shared__ double sT[DN_NB][DN_NB + 1];
    double* sG = Gp + (size_t)nparts * DN_NB * DN_NB;          // the summed Gram matrix goes to the slot after the partial ones
    for (int idx = threadIdx.x; idx < DN_NB * DN_NB; idx += blockDim.x) {
        double g = 0;
        for (int b = 0; b < nparts; ++b) g += Gp[(size_t)b * DN_NB * DN_NB + idx];
        sG[idx] = g;
        sT[idx / DN_NB][idx % DN_NB] = 0.0;
    }
    __syncthreads();
    for (int t = 0; t < nb; ++t) {
        const double tt = tau[j0 + t];
        if (threadIdx.x < t) {                        // T[0:t, t] = -tau_t * T[0:t, 0:t] G[0:t, t]   (T upper triangular)
            double s = 0;
            for (int q = threadIdx.x; q < t; ++q) s += sT[threadIdx.x][q] * sG[q * DN_NB + t];
            sT[threadIdx.x][t] = -tt * s;
        }
        if (threadIdx.x == 0) sT[t][t] = tt;
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < DN_NB * DN_NB; idx += blockDim.x) {
        const int r = idx % DN_NB, cc = idx / DN_NB;
        T[idx] = (r < nb && cc < nb) ? sT[r][cc] : 0.0;
    }
}

// VT = V T (n x nb, column-major ld = ldp): thread per row
__global__ void __launch_bounds__(256) dense_vt_kernel(int n, int row0, int nb, const double* __restrict__ Vp, long ldp,
                                                       const double* __restrict__ T, double* __restrict__ VT) {
    __shared__ double sT[DN_NB * DN_NB];
    for (int idx = threadIdx.x; idx < DN_NB * DN_NB; idx += blockDim.x) sT[idx] = T[idx];
    __syncthreads();
    const int r = row0 + blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    double vr[DN_NB];
#pragma unroll
    for (int q = 0; q < DN_NB; ++q) vr[q] = (q < nb) ? Vp[(long)q * ldp + r] : 0.0;
#pragma unroll 4
    for (int cc = 0; cc < DN_NB; ++cc) {
        double s = 0;
#pragma unroll
        for (int q = 0; q < DN_NB; ++q) s = fma(vr[q], sT[q + cc * DN_NB], s);     // T upper triangular: zeros below
        VT[(long)cc * ldp + r] = s;
    }
}

// W1 = V^T Z  (nb x ncols, ROW-major with leading dimension ldw: the B operand layout of the DMMA GEMM), rows row0..n-1.
// Block: 64 columns of Z x all DN_NB reflectors, rows streamed through shared memory 32 at a time;
// thread (tx, ty) of 16 x 16 accumulates a 4 (reflectors) x 4 (columns) patch.
__global__ void __launch_bounds__(256) dense_vtz_kernel(int n, int row0, int ncols, const double* __restrict__ Vp, long ldp,
                                                        const double* __restrict__ Z, long ldz, double* __restrict__ W1, long ldw) {
    __shared__ double sV[32][DN_NB + 1];
    __shared__ double sZ[32][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = blockIdx.x * 64;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int rb = row0; rb < n; rb += 32) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * DN_NB; idx += 256) {
            const int rr = idx & 31, q = idx >> 5;
            sV[rr][q] = (rb + rr < n) ? Vp[(long)q * ldp + rb + rr] : 0.0;
        }
        for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
            const int rr = idx & 31, cc = idx >> 5;
            sZ[rr][cc] = (rb + rr < n && c0 + cc < ncols) ? Z[(long)(c0 + cc) * ldz + rb + rr] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            double vv[4], zz[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) vv[a] = sV[rr][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) zz[b] = sZ[rr][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(vv[a], zz[b], acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (c0 + tx * 4 + b < ncols) W1[(long)(ty * 4 + a) * ldw + c0 + tx * 4 + b] = acc[a][b];
}

// reflectors of the panel j0 .. j0+nb-1 back out of A's lower triangle (LAPACK layout): Vp[r, t] = 1 at r = j0+t+1, A[r, j0+t]
// below, 0 above and in the unused columns t >= nb
__global__ void __launch_bounds__(256) dense_extract_v_kernel(const double* __restrict__ A, long lda, int n, int j0, int nb,
                                                              double* __restrict__ Vp, long ldp) {
    const int t = blockIdx.y, r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    double v = 0.0;
    if (t < nb && j0 + t + 1 < n) {
        if (r == j0 + t + 1) v = 1.0;
        else if (r > j0 + t + 1) v = A[(long)(j0 + t) * lda + r];
    }
    Vp[(long)t * ldp + r] = v;
}

// one GEMM problem + its 128 x 128 tile list for the DMMA kernel (the merges build theirs from the descriptors)
__global__ void dense_gemm_work_kernel(GemmProblem P, GemmProblem* probs, GemmTile* tiles, int* ntiles) {
    const int ntm = (P.M + 127) / 128, ntn = (P.N + 127) / 128;
    if (blockIdx.x == 0 && threadIdx.x == 0) { probs[0] = P; ntiles[0] = ntm * ntn; }
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ntm * ntn; q += gridDim.x * blockDim.x)
        tiles[q] = GemmTile{0, (q % ntm) * 128, (q / ntm) * 128};
}

__global__ void dense_iota_kernel(int* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
// symmetrise from the lower triangle (the caller may hand over only the lower triangle, like dsytrd 'L')
__global__ void dense_mirror_lower_kernel(double* A, long lda, int n) {
    const int c = blockIdx.x, r0 = blockIdx.y * 256 + threadIdx.x;
    if (r0 < c && r0 < n) A[(long)c * lda + r0] = A[(long)r0 * lda + c];
}
#endif  // CUPPEN_CUDA

}  // namespace cuppen
#endif
