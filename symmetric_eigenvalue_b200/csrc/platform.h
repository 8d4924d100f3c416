// Device plumbing for libcuppen_b200: error checks, device buffers and the generic
// one-thread-per-item launcher used by the O(m) "vector" stages of a merge.
//
// The product is compiled by nvcc for sm_100a only.  CUPPEN_HOST_EMULATION is a TEST-ONLY build
// mode (tests/host/Makefile, g++): the same sources run the per-item functors in a serial loop on
// the host so the tree planning, index bookkeeping and numerics can be unit-tested in a container
// without a GPU.  It is never compiled into libcuppen_b200.so and there is no runtime switch:
// the shipped library has no CPU path.
#ifndef CUPPEN_PLATFORM_H
#define CUPPEN_PLATFORM_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>

#if defined(__CUDACC__) && !defined(CUPPEN_HOST_EMULATION)
#define CUPPEN_CUDA 1
#include <cuda_runtime.h>
#else
#ifndef CUPPEN_HOST_EMULATION
#error "libcuppen_b200 must be compiled with nvcc (sm_100a); the host build is test-only (-DCUPPEN_HOST_EMULATION)"
#endif
#define CUPPEN_CUDA 0
#endif

#if CUPPEN_CUDA
#define CUPPEN_HD __host__ __device__ __forceinline__
#define CUPPEN_D __device__ __forceinline__
#else
#define CUPPEN_HD inline
#define CUPPEN_D inline
#endif

namespace cuppen {

// Reciprocal for the Cauchy-like inner loops (one per pole/root pair in the secular, Loewner, norm and U kernels): the
// hardware seed (MUFU.RCP64H: the low 32 mantissa bits of the operand are ignored, ~20 bits) and ONE cubic step
// r (1 + e + e^2), e = 1 - x r -- 4 FP64 instructions, within 1 ulp of the correctly rounded value (relative error e^3 ~ 2^-60
// before the final rounding; `cuppen_selftest_rcp` measures seed and result on the device, tests/test_gpu_parity.py) --
// instead of the ~20-instruction IEEE division sequence with its slow-path checks, which made those kernels
// issue-bound (profiles/r02_ncu_full_vector_kernels_goe_n16384_raw.csv).  Two Newton steps (5 instructions) were the
// first version; -DCUPPEN_RCP_NEWTON2 brings them back.  x == 0 or subnormal gives NaN / inf: the
// callers treat a non-finite quotient like the IEEE +-inf (a bracket step, a clamped matrix entry).
#if CUPPEN_CUDA
__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rcp_newton2(double x) {
    double r = rcp_seed(x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double rcp_cubic(double x) {
    const double r = rcp_seed(x);
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__host__ __device__ __forceinline__ double fast_rcp(double x) {
#ifdef __CUDA_ARCH__
#ifdef CUPPEN_RCP_NEWTON2
    return rcp_newton2(x);
#else
    return rcp_cubic(x);
#endif
#else
    return 1.0 / x;
#endif
}
// a / x for the stage functors that the host test build shares with the device (merge_stages.h): the reciprocal form on
// the device, the plain IEEE quotient on the host -- where the eigenvalue-only path (RowGemv) and the matrix path
// (ugen + GEMM twins) are required to agree to the last bit (tests/test_host_logic.py)
__host__ __device__ __forceinline__ double fast_div(double a, double x) {
#ifdef __CUDA_ARCH__
    return a * fast_rcp(x);
#else
    return a / x;
#endif
}
#define CUPPEN_RCP(x) ::cuppen::fast_rcp(x)
#define CUPPEN_DIV(a, x) ::cuppen::fast_div((a), (x))
#else
#define CUPPEN_RCP(x) (1.0 / (x))
#define CUPPEN_DIV(a, x) ((a) / (x))
#endif

struct Error {
    int code;
    std::string msg;
};

#define CUPPEN_THROW(code_, ...)                                   \
    do {                                                           \
        char buf_[512];                                            \
        snprintf(buf_, sizeof buf_, __VA_ARGS__);                  \
        throw ::cuppen::Error{(code_), std::string(buf_)};         \
    } while (0)

#if CUPPEN_CUDA
#define CUDA_CHECK(expr)                                                                        \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            CUPPEN_THROW(-10, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__,      \
                         __LINE__, cudaGetErrorString(e_));                                     \
    } while (0)
typedef cudaStream_t Stream;
#else
typedef int Stream;
#endif

// ---- device memory ---------------------------------------------------------------------------
inline void* dev_alloc_bytes(size_t bytes) {
    if (bytes == 0) bytes = 8;
#if CUPPEN_CUDA
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, bytes));
    return p;
#else
    void* p = malloc(bytes);
    if (!p) CUPPEN_THROW(-11, "host emulation: out of memory (%zu bytes)", bytes);
    memset(p, 0xff, bytes);   // poison (NaN pattern) so that reads of unwritten memory show up
    return p;
#endif
}
inline void dev_free(void* p) {
    if (!p) return;
#if CUPPEN_CUDA
    cudaFree(p);
#else
    free(p);
#endif
}
inline void dev_h2d(void* d, const void* h, size_t bytes, Stream s) {
#if CUPPEN_CUDA
    CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
#else
    (void)s; memcpy(d, h, bytes);
#endif
}
inline void dev_d2h(void* h, const void* d, size_t bytes, Stream s) {
#if CUPPEN_CUDA
    CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
#else
    (void)s; memcpy(h, d, bytes);
#endif
}
inline void dev_d2d(void* dst, const void* src, size_t bytes, Stream s) {
#if CUPPEN_CUDA
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
#else
    (void)s; memmove(dst, src, bytes);
#endif
}
inline void dev_zero(void* d, size_t bytes, Stream s) {
#if CUPPEN_CUDA
    CUDA_CHECK(cudaMemsetAsync(d, 0, bytes, s));
#else
    (void)s; memset(d, 0, bytes);
#endif
}
inline void dev_sync(Stream s) {
#if CUPPEN_CUDA
    CUDA_CHECK(cudaStreamSynchronize(s));
#else
    (void)s;
#endif
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;       // false: a view into memory owned by somebody else (the symmetric heap of p2p.h)
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (owned) dev_free(p); }
    void alloc(size_t count) {
        if (owned) dev_free(p);
        p = nullptr;
        n = count;
        owned = true;
        p = static_cast<T*>(dev_alloc_bytes(count * sizeof(T)));
    }
    void attach(T* ptr, size_t count) {
        if (owned) dev_free(p);
        p = ptr; n = count; owned = false;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// ---- generic per-item launcher -----------------------------------------------------------------
// Functor F: `CUPPEN_HD void operator()(long i) const`.  Kernel names in profiles read
// cuppen::per_item_kernel<cuppen::ZAssemble> etc.
#if CUPPEN_CUDA
template <class F>
__global__ void __launch_bounds__(256) per_item_kernel(long n, F f) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}
#endif

struct LaunchCounter {
    long launches = 0;
};
extern LaunchCounter g_launches;

template <class F>
inline void launch_items(Stream s, long n, const F& f) {
    if (n <= 0) return;
#if CUPPEN_CUDA
    long blocks = (n + 255) / 256;
    per_item_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(n, f);
    CUDA_CHECK(cudaGetLastError());
#else
    (void)s;
    for (long i = 0; i < n; ++i) f(i);
#endif
    g_launches.launches++;
}

// ---- one-warp-per-item launcher ------------------------------------------------------------------
// Functor F: `template <class L> CUPPEN_HD void operator()(long i, const L& lanes) const`; the
// inner loops stride over lanes.lane()/lanes.lanes() and combine with lanes.sum/isum/max/prod.
struct SerialLanes {
    CUPPEN_HD int lane() const { return 0; }
    CUPPEN_HD int lanes() const { return 1; }
    CUPPEN_HD double sum(double v) const { return v; }
    CUPPEN_HD double max(double v) const { return v; }
    CUPPEN_HD double prod(double v) const { return v; }
    CUPPEN_HD int isum(int v) const { return v; }
};
#if CUPPEN_CUDA
struct WarpLanes {
    CUPPEN_D int lane() const { return threadIdx.x & 31; }
    CUPPEN_D int lanes() const { return 32; }
    CUPPEN_D double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    CUPPEN_D double max(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    CUPPEN_D double prod(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    CUPPEN_D int isum(int v) const { return __reduce_add_sync(0xffffffffu, v); }
};
template <class F>
__global__ void __launch_bounds__(256) per_warp_kernel(long n, F f) {
    long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i < n) f(i, WarpLanes());
}
#endif

template <class F>
inline void launch_warps(Stream s, long n, const F& f) {
    if (n <= 0) return;
#if CUPPEN_CUDA
    long blocks = (n + 7) / 8;
    per_warp_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(n, f);
    CUDA_CHECK(cudaGetLastError());
#else
    (void)s;
    for (long i = 0; i < n; ++i) f(i, SerialLanes());
#endif
    g_launches.launches++;
}

static inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

}  // namespace cuppen
#endif
