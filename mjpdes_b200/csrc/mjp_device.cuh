// mjp_device.cuh -- device arithmetic of the substream contract
// (include/mjpdes_b200.h, "random numbers").  Every floating-point operation
// that feeds an event time is an explicit round-to-nearest intrinsic, so the
// results are bit-identical to a host implementation of the same formulas
// regardless of -fmad.
#pragma once
#include <cstdint>

namespace mjp {

__device__ __forceinline__ double bits2d(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t d2bits(double x) { return (uint64_t)__double_as_longlong(x); }

// Philox4x32-10 (Salmon et al., SC'11).  counter = {idx, consumer, g_lo, g_hi}, key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// clojure.core/rand stand-in: 53 bits, [0,1)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  const uint64_t k = (((uint64_t)a << 32) | b) >> 11;
  return __dmul_rn(__ull2double_rn(k), 0x1.0p-53);
}
// argument of -ln(): 52 bits, strictly inside (0,1)
__device__ __forceinline__ double u52open(uint32_t a, uint32_t b) {
  const uint64_t k = (((uint64_t)a << 32) | b) >> 12;
  return __dmul_rn(__dadd_rn(__ull2double_rn(k), 0.5), 0x1.0p-52);
}

// ln(x), normal x > 0: x = 2^k m, m in [sqrt(1/2), sqrt 2), s = (m-1)/(m+1),
// ln m = 2s(1 + s^2/3 + ... + s^20/21).
__device__ __forceinline__ double det_log(double x) {
  const uint64_t b = d2bits(x);
  int k = (int)(b >> 52) - 1023;
  const uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
  double m;
  if (mant >= 0x6A09E667F3BCDull) { k += 1; m = bits2d(mant | (1022ull << 52)); }
  else                            {         m = bits2d(mant | (1023ull << 52)); }
  const double f = __dadd_rn(m, -1.0);
  const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
  const double z = __dmul_rn(s, s);
  double q = 1.0 / 21.0;
  q = __fma_rn(q, z, 1.0 / 19.0);
  q = __fma_rn(q, z, 1.0 / 17.0);
  q = __fma_rn(q, z, 1.0 / 15.0);
  q = __fma_rn(q, z, 1.0 / 13.0);
  q = __fma_rn(q, z, 1.0 / 11.0);
  q = __fma_rn(q, z, 1.0 / 9.0);
  q = __fma_rn(q, z, 1.0 / 7.0);
  q = __fma_rn(q, z, 1.0 / 5.0);
  q = __fma_rn(q, z, 1.0 / 3.0);
  const double two_s = __dmul_rn(2.0, s);
  const double dk = (double)k;
  const double lo = __fma_rn(__dmul_rn(two_s, z), q, __dmul_rn(dk, bits2d(0x3DEA39EF35793C76ull)));
  return __dadd_rn(__dmul_rn(dk, bits2d(0x3FE62E42FEE00000ull)), __dadd_rn(two_s, lo));
}

// e^x for x <= 0 (stands in for Math/pow Math/E x, core.clj:165); 0 below -40.
__device__ __forceinline__ double det_exp(double x) {
  if (x < -40.0) return 0.0;
  const double kf = floor(__dadd_rn(__dmul_rn(x, bits2d(0x3FF71547652B82FEull)), 0.5));
  double r = __fma_rn(-kf, bits2d(0x3FE62E42FEE00000ull), x);
  r = __fma_rn(-kf, bits2d(0x3DEA39EF35793C76ull), r);
  double p = 1.0 / 87178291200.0;
  p = __fma_rn(p, r, 1.0 / 6227020800.0);
  p = __fma_rn(p, r, 1.0 / 479001600.0);
  p = __fma_rn(p, r, 1.0 / 39916800.0);
  p = __fma_rn(p, r, 1.0 / 3628800.0);
  p = __fma_rn(p, r, 1.0 / 362880.0);
  p = __fma_rn(p, r, 1.0 / 40320.0);
  p = __fma_rn(p, r, 1.0 / 5040.0);
  p = __fma_rn(p, r, 1.0 / 720.0);
  p = __fma_rn(p, r, 1.0 / 120.0);
  p = __fma_rn(p, r, 1.0 / 24.0);
  p = __fma_rn(p, r, 1.0 / 6.0);
  p = __fma_rn(p, r, 0.5);
  p = __fma_rn(p, r, 1.0);
  p = __fma_rn(p, r, 1.0);
  const int k = (int)kf;
  return __dmul_rn(p, bits2d((uint64_t)(1023 + k) << 52));
}

}  // namespace mjp
