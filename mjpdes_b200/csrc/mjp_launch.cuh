// mjp_launch.cuh -- launchers of the event-loop kernel variants.
//
// launch_flags<MaskT, MT, COLD, LOG> picks the (REPLAY, HASH, SLICE) instantiation of mjp_line_kernel at
// run time.  Each <MaskT, MT, COLD, LOG> family is compiled in its own translation unit (mjp_variant.cu with
// -DMJP_V_*; see the Makefile) so that the ~180 kernels build in parallel; mjp_abi.cu only sees the extern
// template declarations below.
#pragma once
#include <cuda_runtime.h>

#include "mjp_params.h"

namespace mjp {

// occupancy != nullptr: do not launch, report how many CTAs of that instantiation fit on one SM
template <typename MaskT, int MT, bool COLD, bool LOG>
cudaError_t launch_flags(bool replay, bool hash, bool slice, const KParams &kp, int blocks, int threads, size_t smem,
                         cudaStream_t st, int *occupancy);

#define MJP_FAMILIES(X)                                                                                      \
  X(uint32_t, 0, false) X(uint32_t, 2, false) X(uint32_t, 3, false) X(uint32_t, 4, false) X(uint32_t, 5, false) \
  X(uint32_t, 6, false) X(uint32_t, 7, false) X(uint32_t, 8, false) X(uint32_t, 0, true) X(uint64_t, 0, false) \
  X(uint64_t, 0, true)

#ifndef MJP_V_MASK
#define MJP_X(T, MT, COLD)                                                                                   \
  extern template cudaError_t launch_flags<T, MT, COLD, false>(bool, bool, bool, const KParams &, int, int, size_t, cudaStream_t, int *); \
  extern template cudaError_t launch_flags<T, MT, COLD, true>(bool, bool, bool, const KParams &, int, int, size_t, cudaStream_t, int *);
MJP_FAMILIES(MJP_X)
#undef MJP_X
#endif

}  // namespace mjp
