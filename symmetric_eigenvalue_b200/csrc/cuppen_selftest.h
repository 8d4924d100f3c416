/* TEST / BENCH ONLY -- entry points of lib/libcuppen_selftest.so (csrc/selftest.cu): kernel self-tests on random
 * data and FP64 yardsticks.  Not part of the product ABI (include/cuppen_b200.h). */
#ifndef CUPPEN_SELFTEST_H
#define CUPPEN_SELFTEST_H
#ifdef __cplusplus
extern "C" {
#endif

/* FP64 yardsticks measured on the device: register-resident DMMA.8x8x4 issue loop and DFMA loop,
 * TFLOP/s with all SMs busy for ~`ms` milliseconds each (the FP64 peak is not in MEASURED_PEAKS.json). */
int cuppen_measure_fp64_peak(int device, int ms, double* dmma_tflops, double* dfma_tflops);

/* DMMA and DFMA issue loops sharing every SM (4 warps each per block): TFLOP/s of each kind alone and over the common
 * window when both run together -- do the two instruction kinds share the FP64 units? */
int cuppen_measure_fp64_mix(int device, double* dmma_alone, double* dfma_alone, double* dmma_mixed, double* dfma_mixed);

/* GEMM self-test / micro-benchmark of the back-transformation kernels on random data:
 * variant 0 = cp.async DMMA kernel 128x128, 1 = TMA DMMA kernel 128x128, 2 = cp.async 64x64.
 * max_abs_err: against an fp64 FMA dot product on 8192 sampled entries; tflops: best of `reps`. */
int cuppen_selftest_gemm(int device, int variant, int M, int N, int K, int reps, double* max_abs_err, double* tflops);

/* Residual-kernel self-test / micro-benchmark on random data (n columns): one slice of rows [g0, g0+cnt) of an n-row
 * problem stored at local rows [l0, l0+cnt) with halo rows (the multi-GPU slice layout), against a plain per-column
 * loop.  variant: 0 = the default, else 10*columns-per-block + min-blocks-per-SM.  seconds (may be NULL): best of 3. */
int cuppen_selftest_residual(int device, int n, int variant, int g0, int l0, int cnt, double* max_rel_err, double* seconds);

/* Reciprocal of the Cauchy-like inner loops (platform.h) on `count` random operands: max |1 - x*seed| of the hardware
 * seed, and the max distance in ulps from the correctly rounded 1/x of the two-Newton-step and of the cubic-step form. */
int cuppen_selftest_rcp(int device, long count, double* seed_max_rel_err, double* newton2_max_ulp, double* cubic_max_ulp);

const char* cuppen_selftest_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
