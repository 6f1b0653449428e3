// mjp_plan.h -- host-side compilation of a line descriptor into the kernel's
// constant table and shared-memory layout ("preprocess-model" for the device,
// core.clj:578-615).  Pure host C++, no CUDA calls.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mjpdes_b200.h"
#include "mjp_params.h"

namespace mjp {

// hasheq of the Clojure keyword :<name> (clojure 1.9: Keyword.hasheq = Symbol.hasheq + 0x9e3779b9, Symbol.hasheq =
// hashCombine(Murmur3.hashUnencodedChars(name), 0)); names here are ASCII.  Restated from the published Clojure
// sources; (hash :a) = -2123407586 is the known value it reproduces (tests/test_cljhash.py).
inline uint32_t clj_keyword_hash(const std::string &name) {
  auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
  auto mix_k1 = [&](uint32_t k) { k *= 0xcc9e2d51u; k = rotl(k, 15); return k * 0x1b873593u; };
  uint32_t h = 0;
  const size_t n = name.size();
  for (size_t i = 1; i < n; i += 2) {
    h ^= mix_k1((uint32_t)(unsigned char)name[i - 1] | ((uint32_t)(unsigned char)name[i] << 16));
    h = rotl(h, 13); h = h * 5u + 0xe6546b64u;
  }
  if (n & 1) h ^= mix_k1((uint32_t)(unsigned char)name[n - 1]);
  h ^= (uint32_t)(2 * n); h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  const uint32_t seed = h;                                               // hashCombine(seed, 0)
  const uint32_t sym = seed ^ (0u + 0x9e3779b9u + (seed << 6) + (uint32_t)((int32_t)seed >> 2));
  return sym + 0x9e3779b9u;
}
// PersistentHashMap iterates its 32-way trie by ascending 5-bit chunks of the hash, lowest chunk first
inline uint64_t clj_trie_key(uint32_t h) {
  uint64_t k = 0;
  for (int s = 0; s < 35; s += 5) k = (k << 5) | ((h >> s) & 31u);
  return k;
}
// rank[i] of the default machine names :m1 .. :mM in hash-map order
inline std::vector<int32_t> default_group_rank(int M) {
  std::vector<int> order(M);
  std::vector<uint64_t> key(M);
  for (int i = 0; i < M; ++i) { order[i] = i; key[i] = clj_trie_key(clj_keyword_hash("m" + std::to_string(i + 1))); }
  std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<int32_t> rank(M);
  for (int pos = 0; pos < M; ++pos) rank[order[pos]] = pos;
  return rank;
}

struct LinePlan {
  KParams kp{};                 // everything except device pointers, seed and first_id
  std::vector<double> cd;       // lam[M] mu[M] W[M] w[K*M] dur[K*M] thr[K]
  std::vector<int32_t> ci;      // N[M-1] lvl[M-1]
  std::vector<uint8_t> by_rank; // [M] machines in hash-map order of their names (util/log.clj:44 beyond 8 groups)
  size_t per_thread_bytes = 0;

  // offsets (in doubles) of the sections of cd
  size_t off_lam() const { return 0; }
  size_t off_mu() const { return (size_t)kp.M; }
  size_t off_W() const { return 2 * (size_t)kp.M; }
  size_t off_w() const { return 3 * (size_t)kp.M; }
  size_t off_dur() const { return off_w() + (size_t)kp.K * kp.M; }
  size_t off_thr() const { return off_dur() + (size_t)kp.K * kp.M; }
};

inline size_t plan_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Per-thread shared-memory slots of one kernel variant (must mirror the macros in mjp_line_kernel.cuh):
//   cold f64: FUT_T[M] NXT_T[M] BS[M] SS[M] LASTCLK[M-1] (+ ENDS[M] when M > 32 and the layout is cold)         cold u32: FUT_N[M] STARTED[M] JT[R/4] EIDX[M] if REPLAY
//   hot  f64: TICKCLK EJENT TMINUD | ENDS[M] GE[G] GU[G] unless :ends are in registers (G = ceil(M/4) groups)
//   hot  u32: CNT[M-1] (+ STARTED[M] in a cold layout when kp.started_hot: set by the caller BEFORE plan_layout when the
//             extra words cost no CTA per SM)
// SCADA capture (LOG) stages a tick's messages in a global scratch of stage_cap words per instance, not here
// (staging in shared memory was measured twice and lost both times, profiles/README.md).
// Without `cold` the cold slots precede the hot ones in the same shared arrays (u32: FUT_N STARTED CNT JT EIDX);
// with it they live in a global scratch of cold_f64_slots / cold_u32_slots rows per instance.
inline int cold_f64_slots(const KParams &kp) { return 5 * kp.M - 1; }
// lines of more than 32 machines keep :ends in the cold scratch as well (rows 5M-1 .. 6M-2), see mjp_line_kernel.cuh
inline int cold_ends_slots(const KParams &kp) { return kp.M; }      // (allocated for every cold layout; used when kp.ends_cold)
inline int cold_u32_slots(const KParams &kp, bool replay) {
  return 2 * kp.M + (kp.K > 1 ? (kp.R >> 2) : 0) + (replay ? kp.M : 0);      // no job-type ring for a single type
}
#ifndef MJP_COLD_CONST_GLOBAL
#define MJP_COLD_CONST_GLOBAL 1
#endif
// CTA constant table: lam[M] mu[M] dur[K M] thr[K] N[M-1] lvl[M-1]; cold layouts of more than 32 machines keep only
// thr / N / lvl on chip and read the rest from global memory (mjp_line_kernel.cuh)
inline int plan_const_bytes(int M, int K, bool cold) {
  const size_t f64 = (cold && M > 32 && MJP_COLD_CONST_GLOBAL) ? (size_t)K : (size_t)(2 * M + K * M + K);
  return (int)plan_align_up(f64 * 8 + (size_t)2 * M * 4, 16);
}
inline size_t plan_layout(KParams &kp, bool ends_in_regs, bool replay, bool log, bool cold = false) {
  const int M = kp.M;
  (void)log;
  kp.const_bytes = plan_const_bytes(M, kp.K, cold);
  const int groups = (M + 3) / 4;                                  // GS = 4
  const int ends_hot = (cold && kp.ends_cold) ? 0 : M;
  kp.nf64 = (cold ? 0 : cold_f64_slots(kp)) + 3 + (ends_in_regs ? 0 : ends_hot + 2 * groups);
  kp.nu32 = (cold ? 0 : cold_u32_slots(kp, replay)) + (M - 1) + ((cold && kp.started_hot) ? M : 0);
  return (size_t)kp.nf64 * 8 + (size_t)kp.nu32 * 4;
}

inline LinePlan build_plan(const mjp_line_desc *L, uint64_t n_inst) {
  LinePlan pl;
  KParams &kp = pl.kp;
  const int M = L->M, K = L->K;
  kp.n_inst = n_inst;
  kp.M = M; kp.K = K;
  kp.warm_up = L->warm_up; kp.run_to = L->run_to; kp.run_to_job = L->run_to_job;
  kp.jm2dm = L->jm2dm ? 1 : 0; kp.up_down = L->up_down ? 1 : 0; kp.log_on = L->log ? 1 : 0;
  kp.max_lines = L->max_lines;
  for (int i = 0; i < M; ++i) if (L->discipline[i] == MJP_BBS) kp.bbs_mask |= (1ull << i);
  {
    std::vector<int32_t> rank = L->group_rank ? std::vector<int32_t>(L->group_rank, L->group_rank + M) : default_group_rank(M);
    pl.by_rank.assign(M, 0);
    for (int i = 0; i < M; ++i) pl.by_rank[rank[i]] = (uint8_t)i;           // (validated to be a permutation)
    for (int i = 0; i + 1 < M; ++i) if (rank[i] < rank[i + 1]) kp.hlead_mask |= (1ull << i);
  }
  int sumN = 0, levels = 0;
  pl.ci.assign(L->N, L->N + (M - 1));
  for (int b = 0; b < M - 1; ++b) { pl.ci.push_back(levels); levels += L->N[b] + 1; sumN += L->N[b]; }
  kp.L = levels;
  int R = 4;
  while (R < M + sumN) R <<= 1;        // jobs in the line <= M + sum N
  kp.R = R;
  pl.cd.insert(pl.cd.end(), L->lambda, L->lambda + M);
  pl.cd.insert(pl.cd.end(), L->mu, L->mu + M);
  pl.cd.insert(pl.cd.end(), L->W, L->W + M);
  pl.cd.insert(pl.cd.end(), L->w, L->w + (size_t)K * M);
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < M; ++i) pl.cd.push_back(L->w[k * M + i] / L->W[i]);      // job-requires utils.clj:95-101
  {
    // new-job thresholds, core.clj:332-339: accum starts at p0, then adds the CURRENT portion
    double accum = L->portion[0];
    for (int k = 0; k < K; ++k) { pl.cd.push_back(accum); accum = accum + L->portion[k]; }
  }
  kp.const_bytes = plan_const_bytes(M, K, false);
  // SCADA capture stages one tick of messages per instance.  A machine logs at most nine messages in a tick
  // (complete, start, :bl, :ub, :st, :us, :up, :down and a re-fired :up/:down); beyond the cap the instance ends
  // with MJP_INST_TICK_OVERFLOW.  The stage is a global scratch whose unused rows are never touched.
  kp.stage_cap = L->log ? (9 * M + 4 + 15) / 16 * 16 : 0;      // whole 128-byte lines per instance
  pl.per_thread_bytes = plan_layout(kp, false, true, L->log != 0);   // the largest variant
  return pl;
}

}  // namespace mjp
