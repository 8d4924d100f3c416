// TEST / BENCH ONLY: self-tests and FP64 yardsticks of the kernels of libcuppen_b200, built into their own
// library (lib/libcuppen_selftest.so) so that the product ABI (include/cuppen_b200.h) exports nothing but the path.
// Declarations: cuppen_selftest.h.  Loaded by tests/ (kernel parity against plain loops) and by bench.py (the FP64
// peak that the GEMM roofline is quoted against; MEASURED_PEAKS.json has no FP64 entry).
#include <math.h>
#include <string.h>
#include <algorithm>

#include "../../include/cuppen_b200.h"
#include "cuppen_selftest.h"
#include "platform.h"
#include "merge_stages.h"
#include "matrix_stages.h"
#include "gemm_dmma.h"
#include "gemm_tma.h"

namespace cuppen {
LaunchCounter g_launches;
static thread_local std::string g_selftest_error;
template <int BM, int BN, int BK, int WMs, int WNs, int STAGES>
static void launch_gemm(Stream s, const GemmProblem* probs, const GemmTile* tiles, const int* ntiles_ptr, long grid_want) {
    using Cfg = DmmaCfg<BM, BN, BK, WMs, WNs, STAGES>;
    auto kern = dgemm_dmma_kernel<BM, BN, BK, WMs, WNs, STAGES>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    int grid = (int)std::max<long>(1, grid_want);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(probs, tiles, ntiles_ptr);
    CUDA_CHECK(cudaGetLastError());
}
}  // namespace cuppen
using namespace cuppen;

#define CUPPEN_API_BEGIN try {
#define CUPPEN_API_END                                                            \
    }                                                                             \
    catch (const cuppen::Error& e) { g_selftest_error = e.msg; return e.code; }   \
    catch (...) { g_selftest_error = "unknown error"; return CUPPEN_ERR_ARG; }    \
    return CUPPEN_OK;

#if CUPPEN_CUDA
namespace cuppen {
__global__ void fill_kernel(double* p, long count, unsigned seed) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long x = (unsigned long long)i * 6364136223846793005ull + seed * 1442695040888963407ull + 1013904223ull;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 32;
    p[i] = (double)(x >> 11) * (1.0 / 4503599627370496.0) - 1.0;        // 53 random mantissa bits in [-1, 1)
}
__global__ void sample_check_kernel(const GemmProblem P, int samples, double* err) {
    int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= samples) return;
    int m = (int)(((unsigned long long)sidx * 2654435761ull) % (unsigned)P.M);
    int nn = (int)(((unsigned long long)sidx * 40503ull + 17) % (unsigned)P.N);
    double s = 0;
    for (int k = 0; k < P.K; ++k) s = fma(P.A[(long)k * P.lda + m], P.B[(long)k * P.ldb + nn], s);
    double got = P.C[(long)P.colidx[nn] * P.ldc + m];
    atomicMax((unsigned long long*)err, (unsigned long long)__double_as_longlong(fabs(got - s)));
}
__global__ void iota_rev_kernel(int* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = n - 1 - i; }
// plain one-thread-per-column restatement of the residual, for cuppen_selftest_residual
__global__ void residual_check_kernel(const double* V, long ldq, int n, int ncols, int g0, int l0, int cnt, const double* OD, const double* OE,
                                      const double* lam, const int* perm, const double* hlo, const double* hhi, const double* got, double* err) {
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const double* x = V + (long)perm[col] * ldq + l0 - g0;
    double acc = 0;
    for (int r = g0; r < g0 + cnt; ++r) {
        double y = OD[r] * x[r] - lam[col] * x[r];
        if (r > 0) y += OE[r - 1] * (r > g0 ? x[r - 1] : hlo[col]);
        if (r < n - 1) y += OE[r] * (r + 1 < g0 + cnt ? x[r + 1] : hhi[col]);
        acc += y * y;
    }
    double rel = fabs(got[col] - acc) / fmax(acc, 1e-300);
    atomicMax((unsigned long long*)err, (unsigned long long)__double_as_longlong(rel));
}
// cuppen_selftest_rcp: seed and results of the two reciprocal forms of platform.h against the IEEE quotient, over random
// mantissas, both signs and binary exponents in [-500, 500].  out[0] max |1 - x seed|, out[1] / out[2] max distance in
// ulps (difference of the bit patterns) of the two-Newton-step / cubic-step result from __drcp_rn(x).
__global__ void rcp_check_kernel(long count, unsigned seed, unsigned long long* out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long h = (unsigned long long)i * 6364136223846793005ull + seed * 1442695040888963407ull + 1013904223ull;
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    const unsigned long long mant = h & 0x000fffffffffffffull;
    const int ex = 1023 - 500 + (int)((h >> 52) % 1001);
    const unsigned long long sign = (h >> 63) << 63;
    const double x = __longlong_as_double((long long)(sign | ((unsigned long long)ex << 52) | mant));
    const double exact = __drcp_rn(x);
    const double e0 = fabs(fma(-x, rcp_seed(x), 1.0));
    const long long d2 = llabs(__double_as_longlong(rcp_newton2(x)) - __double_as_longlong(exact));
    const long long d3 = llabs(__double_as_longlong(rcp_cubic(x)) - __double_as_longlong(exact));
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(e0));
    atomicMax(&out[1], (unsigned long long)d2);
    atomicMax(&out[2], (unsigned long long)d3);
}
}  // namespace cuppen
#endif

extern "C" {
const char* cuppen_selftest_last_error(void) { return g_selftest_error.c_str(); }

int cuppen_selftest_gemm(int device, int variant, int M, int N, int K, int reps, double* max_abs_err, double* tflops) {
    CUPPEN_API_BEGIN
#if CUPPEN_CUDA
    if (M < 1 || N < 1 || K < 0 || !max_abs_err || !tflops) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    // row offset of the operand inside its buffer: odd for the cp.async kernels and for variants 3 / 5 (the TMA kernels
    // fetch such tiles from one row earlier), even for variants 1 / 4
    const int row0 = (variant == 1 || variant == 4) ? 2 : 3;
    const bool tma = (variant == 1 || variant == 3 || variant == 4 || variant == 5), tensor = (variant == 4 || variant == 5);
    const long lda = round_up(M + row0 + 128, 16), ldb = round_up(N, 16) + 16, ldc = lda;
    const long Kp = round_up(K, K_PAD);
    DevBuf<double> A, Bm, C, err;
    DevBuf<int> colidx;
    DevBuf<GemmProblem> dprob;
    DevBuf<GemmTile> dtiles;
    A.alloc((size_t)lda * (Kp + K_PAD + 1) + 4096); Bm.alloc((size_t)ldb * (Kp + 2 * K_PAD + 2) + 4096);
    C.alloc((size_t)ldc * (N + 1) + 4096); err.alloc(1); colidx.alloc(N);
    Stream s = 0;
    fill_kernel<<<(unsigned)((A.n + 255) / 256), 256>>>(A.p, (long)A.n, 1u);
    fill_kernel<<<(unsigned)((Bm.n + 255) / 256), 256>>>(Bm.p, (long)Bm.n, 2u);
    // zero the K tail of A (columns K..Kp) as pack_kernel does
    if (Kp > K) CUDA_CHECK(cudaMemset(A.p + (size_t)K * lda, 0, sizeof(double) * (size_t)(Kp - K) * lda));
    CUDA_CHECK(cudaMemset(C.p, 0, C.bytes()));
    CUDA_CHECK(cudaMemset(err.p, 0, sizeof(double)));
    iota_rev_kernel<<<(N + 255) / 256, 256>>>(colidx.p, N);
    GemmProblem P;
    P.A = A.p + row0; P.B = Bm.p; P.C = C.p + row0; P.colidx = colidx.p; P.M = M; P.N = N; P.K = K;
    P.lda = lda; P.ldb = ldb; P.ldc = ldc; P.a_row0 = row0; P.a_col0 = 0; P.b_row0 = 0; P.b_col0 = 0; P.c_sub = 0;
    std::vector<GemmTile> ht;
    const int BM = variant == 2 ? 64 : 128, BN = BM;
    CUtensorMap mA, mB;
    if (tensor && (!tma_encode_map(&mA, A.p, lda, (long)(A.n / lda), lda) || !tma_encode_map(&mB, Bm.p, ldb, (long)(Bm.n / ldb), ldb)))
        CUPPEN_THROW(CUPPEN_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    for (int m0 = 0; m0 < M; m0 += BM)
        for (int n0 = 0; n0 < N; n0 += BN) ht.push_back(GemmTile{0, m0, n0});
    dprob.alloc(1); dtiles.alloc(ht.size());
    CUDA_CHECK(cudaMemcpy(dprob.p, &P, sizeof P, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dtiles.p, ht.data(), sizeof(GemmTile) * ht.size(), cudaMemcpyHostToDevice));
    DevBuf<int> dnt, dabort;
    dnt.alloc(4); dabort.alloc(8);
    CUDA_CHECK(cudaMemset(dabort.p, 0, sizeof(int) * 8));
    int hnt[4] = {(int)ht.size(), 0, 0, 0};
    CUDA_CHECK(cudaMemcpy(dnt.p, hnt, sizeof hnt, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < std::max(1, reps) + 1; ++r) {
        CUDA_CHECK(cudaEventRecord(e0, s));
        const long nt = (long)ht.size();
        if (tma) {
            CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes()));
            CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes()));
            launch_gemm_tma(s, dprob.p, dtiles.p, dnt.p, (int)std::min<long>(nt, prop.multiProcessorCount), dabort.p, tensor ? &mA : nullptr, tensor ? &mB : nullptr);
        }
        else if (variant == 2) launch_gemm<64, 64, 16, 2, 2, 3>(s, dprob.p, dtiles.p, dnt.p, std::min<long>(nt, prop.multiProcessorCount * 8L));
        else launch_gemm<128, 128, 16, 2, 4, 3>(s, dprob.p, dtiles.p, dnt.p, std::min<long>(nt, prop.multiProcessorCount * 2L));
        CUDA_CHECK(cudaEventRecord(e1, s));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 || reps <= 0) best = std::min(best, ms);
    }
    { int hab[8]; CUDA_CHECK(cudaMemcpy(hab, dabort.p, sizeof hab, cudaMemcpyDeviceToHost)); tma_check_abort(hab); }
    const int samples = 8192;
    sample_check_kernel<<<(samples + 127) / 128, 128>>>(P, samples, err.p);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(max_abs_err, err.p, sizeof(double), cudaMemcpyDeviceToHost));
    *tflops = 2.0 * M * (double)N * K / (best * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
#else
    (void)device; (void)variant; (void)M; (void)N; (void)K; (void)reps; (void)max_abs_err; (void)tflops;
    CUPPEN_THROW(CUPPEN_ERR_CUDA, "no GPU in the host test build");
#endif
    CUPPEN_API_END
}

int cuppen_selftest_residual(int device, int n, int variant, int g0, int l0, int cnt, double* max_rel_err, double* seconds) {
    CUPPEN_API_BEGIN
#if CUPPEN_CUDA
    if (n < 1 || g0 < 0 || l0 < 0 || cnt < 1 || g0 + cnt > n || !max_rel_err) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument");
    const int ncols = n;
    CUDA_CHECK(cudaSetDevice(device));
    long ldq = round_up(l0 + cnt, 16);
    if (getenv("CUPPEN_LDPAD")) ldq += atoi(getenv("CUPPEN_LDPAD"));      // experiment: column stride away from a power of two
    DevBuf<double> V, OD, OE, lam, hlo, hhi, res, err;
    DevBuf<int> perm;
    V.alloc((size_t)ldq * ncols + 64); OD.alloc(n + 64); OE.alloc(n + 64); lam.alloc(ncols); hlo.alloc(ncols); hhi.alloc(ncols);
    res.alloc(ncols); err.alloc(1); perm.alloc(ncols);
    fill_kernel<<<(unsigned)((V.n + 255) / 256), 256>>>(V.p, (long)V.n, 11u);
    fill_kernel<<<(unsigned)((OD.n + 255) / 256), 256>>>(OD.p, (long)OD.n, 12u);
    fill_kernel<<<(unsigned)((OE.n + 255) / 256), 256>>>(OE.p, (long)OE.n, 13u);
    fill_kernel<<<(unsigned)((ncols + 255) / 256), 256>>>(lam.p, ncols, 14u);
    fill_kernel<<<(unsigned)((ncols + 255) / 256), 256>>>(hlo.p, ncols, 15u);
    fill_kernel<<<(unsigned)((ncols + 255) / 256), 256>>>(hhi.p, ncols, 16u);
    iota_rev_kernel<<<(ncols + 255) / 256, 256>>>(perm.p, ncols);
    CUDA_CHECK(cudaMemset(err.p, 0, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_CHECK(cudaEventRecord(e0, 0));
        ResSlices rs;
        memset(&rs, 0, sizeof rs);
        rs.ns = 1; rs.g0[0] = g0; rs.l0[0] = l0; rs.cnt[0] = cnt; rs.lo[0] = hlo.p; rs.hi[0] = hhi.p;
        launch_residual(0, variant, V.p, ldq, n, rs, OD.p, OE.p, lam.p, perm.p, res.p);
        CUDA_CHECK(cudaEventRecord(e1, 0));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (seconds) *seconds = best * 1e-3;
    residual_check_kernel<<<(ncols + 127) / 128, 128>>>(V.p, ldq, n, ncols, g0, l0, cnt, OD.p, OE.p, lam.p, perm.p, hlo.p, hhi.p, res.p, err.p);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(max_rel_err, err.p, sizeof(double), cudaMemcpyDeviceToHost));
#else
    (void)device; (void)n; (void)variant; (void)g0; (void)l0; (void)cnt; (void)max_rel_err; (void)seconds;
    CUPPEN_THROW(CUPPEN_ERR_CUDA, "no GPU in the host test build");
#endif
    CUPPEN_API_END
}

int cuppen_selftest_rcp(int device, long count, double* seed_max_rel_err, double* newton2_max_ulp, double* cubic_max_ulp) {
    CUPPEN_API_BEGIN
#if CUPPEN_CUDA
    if (!seed_max_rel_err || !newton2_max_ulp || !cubic_max_ulp || count < 1) CUPPEN_THROW(CUPPEN_ERR_ARG, "bad argument");
    CUDA_CHECK(cudaSetDevice(device));
    unsigned long long* out = nullptr;
    CUDA_CHECK(cudaMalloc(&out, 3 * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemset(out, 0, 3 * sizeof(unsigned long long)));
    rcp_check_kernel<<<(unsigned)((count + 255) / 256), 256>>>(count, 12345u, out);
    CUDA_CHECK(cudaGetLastError());
    unsigned long long h[3];
    CUDA_CHECK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
    cudaFree(out);
    double e0;
    memcpy(&e0, &h[0], sizeof e0);
    *seed_max_rel_err = e0;
    *newton2_max_ulp = (double)h[1];
    *cubic_max_ulp = (double)h[2];
#else
    (void)device; (void)count; (void)seed_max_rel_err; (void)newton2_max_ulp; (void)cubic_max_ulp;
    CUPPEN_THROW(CUPPEN_ERR_CUDA, "no GPU in the host test build");
#endif
    CUPPEN_API_END
}

int cuppen_measure_fp64_peak(int device, int ms, double* dmma_tflops, double* dfma_tflops) {
    CUPPEN_API_BEGIN
#if CUPPEN_CUDA
    if (!dmma_tflops || !dfma_tflops) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 256;
    double* out = nullptr;
    CUDA_CHECK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    for (int which = 0; which < 2; ++which) {
        int iters = 2000;
        double best = 0;
        for (int rep = 0; rep < 6; ++rep) {
            CUDA_CHECK(cudaEventRecord(a));
            if (which == 0) dmma_peak_kernel<<<blocks, threads>>>(out, iters);
            else dfma_peak_kernel<<<blocks, threads>>>(out, iters);
            CUDA_CHECK(cudaEventRecord(b));
            CUDA_CHECK(cudaEventSynchronize(b));
            float t = 0; cudaEventElapsedTime(&t, a, b);
            const double per_thread_iter = which == 0 ? 16.0 * 512.0 / 32.0 : 32.0 * 2.0;
            const double fl = (double)blocks * threads * iters * per_thread_iter;
            best = std::max(best, fl / (t * 1e-3) * 1e-12);
            if (t < ms && rep < 3) iters = (int)std::min(2.0e8, iters * std::max(2.0, (double)ms / std::max(t, 0.01f)));
        }
        *(which == 0 ? dmma_tflops : dfma_tflops) = best;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
#else
    (void)device; (void)ms; (void)dmma_tflops; (void)dfma_tflops;
    CUPPEN_THROW(CUPPEN_ERR_CUDA, "no GPU in the host test build");
#endif
    CUPPEN_API_END
}

// DMMA and DFMA loops sharing every SM: rates of each kind alone (half of the warps idle) and together.
int cuppen_measure_fp64_mix(int device, double* dmma_alone, double* dfma_alone, double* dmma_mixed, double* dfma_mixed) {
    CUPPEN_API_BEGIN
#if CUPPEN_CUDA
    if (!dmma_alone || !dfma_alone || !dmma_mixed || !dfma_mixed) CUPPEN_THROW(CUPPEN_ERR_ARG, "null argument");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 256;
    double* out = nullptr;
    CUDA_CHECK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    // iteration counts sized so that each kind alone runs ~20 ms: DMMA 16 x 512 flop per warp-iteration,
    // DFMA 32 x 2 flop per thread-iteration
    const int it_dmma = 100000, it_dfma = 400000;           // ~26 ms each when alone
    auto run = [&](int i1, int i2) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CUDA_CHECK(cudaEventRecord(a));
            fp64_mix_kernel<<<blocks, threads>>>(out, i1, i2);
            CUDA_CHECK(cudaEventRecord(b));
            CUDA_CHECK(cudaEventSynchronize(b));
            float t = 0; cudaEventElapsedTime(&t, a, b);
            best = std::min(best, t);
        }
        return (double)best * 1e-3;
    };
    const double fl_dmma = (double)blocks * 4 * it_dmma * 16.0 * 512.0;          // 4 DMMA warps per block
    const double fl_dfma = (double)blocks * 128 * it_dfma * 64.0;                // 128 DFMA threads per block
    const double t1 = run(it_dmma, 0), t2 = run(0, it_dfma), t3 = run(it_dmma, it_dfma);
    *dmma_alone = fl_dmma / t1 * 1e-12;
    *dfma_alone = fl_dfma / t2 * 1e-12;
    // together: both finish inside t3 (the slower kind defines it); report the rates over the common window
    *dmma_mixed = fl_dmma / t3 * 1e-12;
    *dfma_mixed = fl_dfma / t3 * 1e-12;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
#else
    (void)device; (void)dmma_alone; (void)dfma_alone; (void)dmma_mixed; (void)dfma_mixed;
    CUPPEN_THROW(CUPPEN_ERR_CUDA, "no GPU in the host test build");
#endif
    CUPPEN_API_END
}

}  // extern "C"
