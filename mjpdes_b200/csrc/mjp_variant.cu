// mjp_variant.cu -- one family <MJP_V_MASK, MJP_V_MT, MJP_V_COLD, MJP_V_LOG> of event-loop kernels
// (8 instantiations: REPLAY x HASH x SLICE).  Compiled once per family by the Makefile.
#include "mjp_launch.cuh"
#include "mjp_line_kernel.cuh"

namespace mjp {

template <typename MaskT, int MT, bool REPLAY, bool HASH, bool SLICE, bool LOG, bool COLD>
static cudaError_t launch_variant(const KParams &kp, int blocks, int threads, size_t smem, cudaStream_t st, int *occupancy) {
  auto k = mjp_line_kernel<MaskT, MT, REPLAY, HASH, SLICE, LOG, COLD>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (occupancy) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, k, threads, smem);
  k<<<blocks, threads, smem, st>>>(kp);
  return cudaGetLastError();
}

template <typename MaskT, int MT, bool COLD, bool LOG>
cudaError_t launch_flags(bool replay, bool hash, bool slice, const KParams &kp, int blocks, int threads, size_t smem,
                         cudaStream_t st, int *occupancy) {
  const int v = (replay ? 4 : 0) | (hash ? 2 : 0) | (slice ? 1 : 0);
#define MJP_CASE(n) \
  case n: return launch_variant<MaskT, MT, ((n) & 4) != 0, ((n) & 2) != 0, ((n) & 1) != 0, LOG, COLD>(kp, blocks, threads, smem, st, occupancy);
  switch (v) {
    MJP_CASE(0) MJP_CASE(1) MJP_CASE(2) MJP_CASE(3) MJP_CASE(4) MJP_CASE(5) MJP_CASE(6) default: MJP_CASE(7)
  }
#undef MJP_CASE
}

template cudaError_t launch_flags<MJP_V_MASK, MJP_V_MT, MJP_V_COLD, MJP_V_LOG>(bool, bool, bool, const KParams &, int, int,
                                                                             size_t, cudaStream_t, int *);

}  // namespace mjp
