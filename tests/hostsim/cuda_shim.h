// cuda_shim.h -- TEST-ONLY: just enough of the CUDA device vocabulary to compile
// mjpdes_b200/csrc/mjp_line_kernel.cuh as host C++ with ONE thread per block, so the
// kernel's control logic can be checked against the oracle on a machine without a
// GPU (pytest -m "not gpu").  Not part of the product: mjpdes_b200 never builds,
// loads or falls back to this; the numbers that count come from the real kernel.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n)
#define __shared__
struct shim_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local shim_dim3 threadIdx_v, blockIdx, gridDim;
static const shim_dim3 blockDim;   // always 1 thread
#define threadIdx shim_tid
static const struct { unsigned x = 0, y = 0, z = 0; } shim_tid;
// the kernels declare their dynamic shared memory `extern` at function scope inside namespace mjp
namespace mjp {
alignas(16) unsigned char smem_raw[1 << 20];
alignas(16) double red[4096];
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline unsigned atomicAdd(unsigned *a, unsigned v) { unsigned o = *a; *a = o + v; return o; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __ull2double_rn(unsigned long long k) { return (double)k; }
static inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
static inline int __ffs(int m) { return __builtin_ffs(m); }
static inline int __ffsll(long long m) { return __builtin_ffsll(m); }
static inline int __popc(unsigned m) { return __builtin_popcount(m); }
static inline int __popcll(unsigned long long m) { return __builtin_popcountll(m); }
static inline double atomicAdd(double *a, double v) { double o = *a; *a = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *a, unsigned long long v) { unsigned long long o = *a; *a = o + v; return o; }
using std::floor;
using std::fmax;
using std::fmin;
// one thread per "warp": votes and warp barriers are the identity
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int pred) { return pred; }
static inline int __all_sync(unsigned, int pred) { return pred; }
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline void __stcs(double2 *p, double2 v) { *p = v; }
#define __maxnreg__(...)
#define __noinline__
