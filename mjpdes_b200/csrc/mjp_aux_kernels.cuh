// mjp_aux_kernels.cuh -- the small kernels beside the event loop: the aggregate (calc-basics per instance ->
// moment vector) and the end-time probe.  Included by mjp_abi.cu only (plain __global__ definitions).
#pragma once
#include "mjp_line_kernel.cuh"

namespace mjp {

// ---- job-requires (utils.clj:95-101) for sweeps: w/W of every (job type, machine, instance), divided once --------
// The same correctly rounded division the loop would do at every job start (dur_of in mjp_line_kernel.cuh).
__global__ void mjp_dur_plane(const KParams p, double *dur) {
  const size_t n = p.n_inst, total = (size_t)p.K * p.M * n;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const size_t row = q / n, g = q - row * n;
    const int i = (int)(row % p.M);
    const double wv = p.w_i ? p.w_i[q] : p.w[row];
    const double Wv = p.W_i ? p.W_i[(size_t)i * n + g] : p.W[i];
    dur[q] = __ddiv_rn(wv, Wv);
  }
}

// ---- aggregate: calc-basics (core.clj:617-634) per instance, then sum(x), sum(x^2) -------------
// One block-wide tree per metric; partials[block][metric] are summed in a fixed
// order by mjp_aggregate_final, so the result does not depend on scheduling.
__global__ void mjp_aggregate_partial(const KParams p, double *partials, int n_metrics) {
  extern __shared__ double red[];      // [2][blockDim]
  const int tid = threadIdx.x, nt = blockDim.x, M = p.M;
  const uint64_t n = p.n_inst;
  const double run_time = p.run_to - p.warm_up;
  for (int metric = -1; metric < n_metrics; ++metric) {
    double s1 = 0.0, s2 = 0.0;
    for (uint64_t g = (uint64_t)blockIdx.x * nt + tid; g < n; g += (uint64_t)gridDim.x * nt) {
      if (p.status[g] != MJP_INST_NORMAL_END) continue;
      double x;
      if (metric < 0) x = 1.0;
      else if (metric == 0) x = (double)p.njobs[g] / run_time;
      else if (metric == 1) x = p.njobs[g] ? p.residence[g] / (double)p.njobs[g] : 0.0;
      else if (metric < 2 + M) x = p.blocked[(size_t)(metric - 2) * n + g] / run_time;
      else if (metric < 2 + 2 * M) x = p.starved[(size_t)(metric - 2 - M) * n + g] / run_time;
      else {
        const int b = metric - 2 - 2 * M;
        const int lv = p.lvl[b], cap = p.N[b];
        double acc = 0.0;
        for (int k = 0; k <= cap; ++k) acc = acc + (double)k * p.occ[(size_t)(lv + k) * n + g];
        x = acc / run_time;
      }
      s1 += x; s2 += x * x;
    }
    red[tid] = s1; red[nt + tid] = s2;
    __syncthreads();
    for (int off = nt >> 1; off > 0; off >>= 1) {
      if (tid < off) { red[tid] += red[tid + off]; red[nt + tid] += red[nt + tid + off]; }
      __syncthreads();
    }
    if (tid == 0) {
      const int slot = (metric < 0) ? 0 : 1 + 2 * metric;
      double *row = partials + (size_t)blockIdx.x * (1 + 2 * n_metrics);
      row[slot] = red[0];
      if (metric >= 0) row[slot + 1] = red[nt];
    }
    __syncthreads();
  }
}

__global__ void mjp_aggregate_final(const double *partials, int n_blocks, int len, double *out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= len) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += partials[(size_t)b * len + q];
  out[q] = s;
}

// ---- probe: end-time (core.clj:138-151) over a literal :up&down schedule ------------------------
__global__ void mjp_end_time_probe(const uint8_t *kinds, const double *times, int n, double clock, double dur,
                                   double *out) {
  if (blockIdx.x || threadIdx.x) return;
  const double QNAN = __longlong_as_double(0x7FF8000000000000ll);
  int k = 0;
  while (k < n && times[k] < clock) ++k;                     // drop-while core.clj:143
  if (k >= n) { *out = QNAN; return; }
  double v;
  bool pending;
  if (kinds[k] == 1) { pending = end_time_begin(true, times[k], clock, dur, v); ++k; }   // :future is :down
  else pending = end_time_begin(false, 0.0, clock, dur, v);
  while (pending) {                                          // k is at the next :up event
    if (k + 1 >= n) { v = QNAN; break; }
    pending = end_time_resume(times[k], times[k + 1], v, v);
    k += 2;
  }
  *out = v;
}

}  // namespace mjp
