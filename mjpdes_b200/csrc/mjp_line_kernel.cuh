// mjp_line_kernel.cuh -- the MJPdes run loop, one line instance per thread.
//
// Replaces core/main-loop-loop (core.clj:729-744) and everything it calls:
// runables and its nine guards (core.clj:343-531), the actions (core.clj:215-326),
// end-time / machine-future (core.clj:138-151,194-198), the up/down process
// (core.clj:156-192), new-job (core.clj:328-340), log / push-log / clean-log-buf /
// add-compute-log! (util/log.clj:30-109,159-217).  Paths are relative to the
// reference's src/gov/nist/MJPdes/.
//
// How the reference's data structures map onto the machine:
//  * the nine guard scans are boolean algebra on per-instance bitmasks
//    (occupied / up / blocked / starved / buffer-full / feed-empty), bit i = machine i;
//  * per-machine numbers (:future time+index, :status :ends, interval starts)
//    live in shared memory as [slot][thread], so lanes that are looking at
//    different machines never bank-conflict;
//  * jobs are FIFO through a serial line, so a job is its id: the job on machine i
//    is #started[i], buffer contents are consecutive ids, and :type / :enters sit in
//    rings indexed by id;
//  * end-time's look-ahead over future up/down events is resolved lazily: a job
//    that does not fit in the current up-interval keeps its remaining work and is
//    resolved when the machine's next :up event fires -- the same f64 operations in
//    the same order as core.clj:145-151, so :ends is bit-identical, and every
//    up/down event is generated exactly once;
//  * random draws are made for the whole warp at once (Philox + ln is ~200 instructions
//    that only ~5 % of the events need): the NEXT up/down event of a machine is generated,
//    out of line, for every lane that lacks one when the first lane cannot wait any longer,
//    and job types are drawn two jobs ahead for every lane that is down to one;
//  * a clock tick's messages are not stored: what clean-log-buf + add-compute-log!
//    would do with them is a function of (machines seen, in which order; parity of
//    :bl/:ub and :st/:us per machine -- a pair cancels; which buffers got a :bj / :sm),
//    all kept as bitmasks and applied when the next tick's first message arrives
//    (with LOG the messages ARE staged, in an L2-resident scratch, and written as records);
//  * accumulators are added straight into the caller-visible output planes with
//    fire-and-forget RED.ADD.F64 (one writer per address: order and rounding are
//    those of the sequential sum).
//
// Execution model: one micro-step (one trip of the reference's loop/recur) per trip of the
// kernel loop, the warp in lock-step; every batch section is a loop voted with __any_sync
// over the lanes that are active in this trip, and no lane leaves the body early -- without
// this the lanes of a warp drift apart for good (ncu: 3.9 of 32 lanes per instruction).
// All instances of a batch take the same time, so the batch must be ONE wave: <= 80
// registers and <= ~19 KB of shared memory per 64-thread CTA (11-12 CTAs per SM) are hard
// constraints, which is why the generators are out of line, the shared layout is sized per
// variant, and slicing / tapes / SCADA capture / wide-line (COLD) layouts are variants.
#pragma once
#include "mjp_device.cuh"
#include "mjp_params.h"
#include "../../include/mjpdes_b200.h"

namespace mjp {

template <typename MaskT> struct MaskOps;
template <> struct MaskOps<uint32_t> {
  static __device__ __forceinline__ int ffs(uint32_t m) { return __ffs((int)m) - 1; }
  static __device__ __forceinline__ int popc(uint32_t m) { return __popc(m); }
};
template <> struct MaskOps<uint64_t> {
  static __device__ __forceinline__ int ffs(uint64_t m) { return __ffsll((long long)m) - 1; }
  static __device__ __forceinline__ int popc(uint64_t m) { return __popcll(m); }
};

__device__ __forceinline__ void hmix(uint64_t &h, uint64_t w) { h = (h ^ w) * 0x100000001b3ull; }

// release / acquire on a device-scope flag, and a short sleep while polling it (rotation, see the kernel)
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
#ifdef __CUDA_ARCH__
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return *p;
#endif
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ void nano_sleep(unsigned ns) {
#ifdef __CUDA_ARCH__
  __nanosleep(ns);
#else
  (void)ns;
#endif
}

#ifndef MJP_COLD32_LOG_MINBLOCKS
#define MJP_COLD32_LOG_MINBLOCKS 8
#endif
#ifndef MJP_MT_LOG_MINBLOCKS
#define MJP_MT_LOG_MINBLOCKS 10
#endif
#ifndef MJP_FLUSH_MERGED
#define MJP_FLUSH_MERGED 1   /* cold layouts: one flush loop that takes an interval of each kind and a buffer per trip */
#endif
#ifndef MJP_EMIT_RUNNING_POINTER
#define MJP_EMIT_RUNNING_POINTER 1
#endif
#ifndef MJP_HOT32_MINBLOCKS
#define MJP_HOT32_MINBLOCKS 6
#endif
// CTAs per SM each kernel family is compiled for, i.e. its register budget (65536 / (64 n), in steps of 8):
//  * lines of up to 8 machines (MT > 0) and the 32-bit cold statistics kernel: 12 -> 80 registers.  The headline batch
//    fits one wave that way, and the 10-machine sweep runs 12 CTAs per SM (measured: compiled for 8 it loses 23 %);
//  * lines of up to 8 machines WITH capture, or replaying the reference's draw loop (REPLAY / TAPE): 10 -> 96
//    registers.  At 80 that code spills; a batch that no longer fits one wave is rotated (measured on example.clj
//    with capture, 104 192 instances: 41.2 -> 38.4 ms);
//  * the 32-bit cold kernel WITH capture: 8 -> 128 registers.  Its shared memory holds 7-8 CTAs of the 20-machine
//    line anyway, and at 80 registers the capture code spills (measured: C4 capture 275.6 -> 208.0 ms);
//  * the 32-bit generic kernel with everything in shared memory (9..32 machines, one wave): shared memory holds at
//    most 5-6 CTAs of it, so 6 -> 168 registers costs no residency;
//  * 64-bit masks: 5 (all state in shared memory) / 7 (cold), as measured in round 1.
__host__ __device__ constexpr int line_kernel_min_ctas(size_t mask_bytes, int mt, bool log, bool cold, bool replay) {
  return mask_bytes == 4 ? (mt > 0 ? ((log || replay) ? MJP_MT_LOG_MINBLOCKS : 12) : (cold ? (log ? MJP_COLD32_LOG_MINBLOCKS : 12) : MJP_HOT32_MINBLOCKS)) : (cold ? 7 : 5);
}

// fire-and-forget add into an output plane
__device__ __forceinline__ void red_add(double *addr, double v) { atomicAdd(addr, v); }

// ---- SCADA records (LOG) -------------------------------------------------------------------------
// class of a message for clean-log-buf (util/log.clj:46-56): 0 = :bl/:ub, 1 = :st/:us, 2 = the rest
__device__ __forceinline__ unsigned log_class(unsigned act) {
  return (act == MJP_ACT_BL || act == MJP_ACT_UB) ? 0u : ((act == MJP_ACT_ST || act == MJP_ACT_US) ? 1u : 2u);
}
struct LogTick {                       // what the records of one tick of one instance have in common
  mjp_log_record *out;
  unsigned long long base, capacity;   // first reserved record, size of the record buffer
  double clk, aj_ends, ej_ent;
  uint32_t line_cnt, inst;
  int k;                               // records reserved for this tick
};
// A staged message is laid out like the record's last two words so that emitting it is two register moves:
// bits 0-31 the job number, bits 32-63 what follows `line` in the record: n | ord << 8 | act << 16 | m << 24
// (ord = the message's index in :log-buf, i.e. its place in the tick's execution order).
__device__ __forceinline__ unsigned long long stage_word(unsigned act, unsigned m, unsigned n, unsigned ord, uint32_t job) {
  return (unsigned long long)job | ((unsigned long long)((n & 0xFFu) | ((ord & 0xFFu) << 8) | (act << 16) | (m << 24)) << 32);
}
__device__ __forceinline__ unsigned stage_act(unsigned long long w) { return ((unsigned)(w >> 32) >> 16) & 0xFFu; }
__device__ __forceinline__ unsigned stage_m(unsigned long long w) { return (unsigned)(w >> 32) >> 24; }
// record number q of the tick from staged word w.
// Returns 0 = written, 1 = buffer full (dropped), 2 = more records than reserved.
__device__ __forceinline__ int put_record(const LogTick &t, int q, unsigned long long w) {
  const unsigned long long at = t.base + (unsigned)q;
  if (q >= t.k) return 2;
  if (at >= t.capacity) return 1;
  const unsigned act = stage_act(w);
  const double aux = act == MJP_ACT_AJ ? t.aj_ends : (act == MJP_ACT_EJ ? t.ej_ent : __longlong_as_double(0x7FF8000000000000ll));
  const unsigned long long w2 = (unsigned long long)(uint32_t)w | ((unsigned long long)t.inst << 32);
  const unsigned long long w3 = (unsigned long long)(t.line_cnt + (uint32_t)q) | (w & 0xFFFFFFFF00000000ull);
  // two 16-byte streaming stores = one full 32-byte sector per record, not kept in L2
  double2 *dst = reinterpret_cast<double2 *>(t.out + at);
  __stcs(dst, make_double2(t.clk, aux));
  __stcs(dst + 1, make_double2(__longlong_as_double((long long)w2), __longlong_as_double((long long)w3)));
  return 0;
}
// clean-log-buf in general (util/log.clj:38-57): machines in order of first appearance (group-by :m), within a
// machine the classes in order of first appearance, within a class the original order; a machine's :bl/:ub
// (:st/:us) are dropped iff exactly two (masks two_b / two_s).  Out of line: ticks that are not already in
// this order are the exception.  Returns records written | dropped << 16 | mismatch << 31.
// by_rank != nullptr: more than 8 machines logged in this tick, so the groups come in hash-map order of the
// machine names (by_rank[0..M) = machines in that order) instead of first appearance.
template <typename MaskT>
__device__ __noinline__ unsigned emit_general(const LogTick t, const unsigned long long *stage, int ns, MaskT two_b,
                                              MaskT two_s, const uint8_t *by_rank, int M) {
  unsigned q = 0, dropped = 0, bad = 0;
  MaskT done_m = 0;                                          // machines whose group has been written
  for (int a0 = 0; a0 < (by_rank ? M : ns); ++a0) {
    int a = a0;
    unsigned m;
    if (by_rank) {
      m = by_rank[a0];
      for (a = 0; a < ns && stage_m(stage[a]) != m; ++a) {}
      if (a == ns) continue;                                 // this machine logged nothing in the tick
    } else m = stage_m(stage[a]);
    const MaskT mbit = (MaskT)1 << m;
    if (done_m & mbit) continue;
    done_m |= mbit;
    uint32_t order = 0;                                      // classes of machine m in order of first appearance
    int no = 0;
    unsigned have = 0;
    for (int b = a; b < ns; ++b) {
      const unsigned long long wb = stage[b];
      if (stage_m(wb) != m) continue;
      const unsigned c = log_class(stage_act(wb));
      if (!((have >> c) & 1u)) { have |= 1u << c; order |= c << (2 * no); ++no; }
    }
    for (int r = 0; r < no; ++r) {
      const unsigned c = (order >> (2 * r)) & 3u;
      if ((c == 0u && (two_b & mbit)) || (c == 1u && (two_s & mbit))) continue;         // util/log.clj:51-53
      for (int b = a; b < ns; ++b) {
        const unsigned long long w = stage[b];
        if (stage_m(w) != m || log_class(stage_act(w)) != c) continue;
        const int rc = put_record(t, (int)q, w);
        dropped += rc == 1; bad |= rc == 2;
        ++q;
      }
    }
  }
  return (q & 0xFFFFu) | ((dropped & 0x7FFFu) << 16) | (bad << 31);
}

// mjp_log_out.first_state: word i = number of the last job started on machine i, word M+i = bit 0 machine i holds a
// job, bits 8-15 (count :holding) of buffer i.  Out of line: called once per instance.
static __device__ __noinline__ void write_line_state(uint32_t *out, const uint32_t *started, uint32_t started_stride,
                                                     const uint32_t *cnt, uint32_t cnt_stride, uint64_t occ, int M) {
  for (int i = 0; i < M; ++i) {
    out[i] = started[(size_t)i * started_stride];
    out[M + i] = (uint32_t)((occ >> i) & 1u) | (i < M - 1 ? (cnt[(size_t)i * cnt_stride] & 0xFFu) << 8 : 0u);
  }
}

// end-time (core.clj:138-151) split at the points where it needs an event that may not exist yet.
// begin: at job start.  `machine_up`: :future is a :down event at fut_t.  Returns true when the job
// does not fit before the next :down; v is then the work still to do, else v is :ends.
__device__ __forceinline__ bool end_time_begin(bool machine_up, double fut_t, double clock, double na, double &v) {
  if (machine_up) {
    const double avail = fut_t - clock;                     // (- etime now)
    if (avail > na) { v = clock + na; return false; }       // core.clj:147-148
    v = na - avail; return true;                            // core.clj:149-150
  }
  v = na; return true;                                      // :up event first: now <- etime, core.clj:145-146
}
// resume: the machine's next :up event (t_up) and the :down after it (t_down) are known.
__device__ __forceinline__ bool end_time_resume(double t_up, double t_down, double rem, double &v) {
  const double avail = t_down - t_up;
  if (avail > rem) { v = t_up + rem; return false; }
  v = rem - avail; return true;
}

// The heavy arithmetic (Philox + ln, ~180 instructions) is needed by ~5 % of the events.  It is kept
// out of line so that its temporaries do not raise the register pressure of the whole micro-step loop
// (inlined, the loop spilled its bitmasks to local memory).
#ifndef MJP_NOINLINE
#define MJP_NOINLINE __noinline__
#endif
// sojourn of COUNTER mode: Erlang-2 from one Philox block {idx, consumer, gid}
static __device__ MJP_NOINLINE double sojourn_counter(uint32_t idx, uint32_t consumer, uint64_t gid, uint64_t seed, double rate) {
  uint32_t o[4];
  philox4x32_10(idx, consumer, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return __ddiv_rn(-det_log(__dmul_rn(u52open(o[0], o[1]), u52open(o[2], o[3]))), rate);
}
// sojourn of REPLAY / TAPE mode: the literal loop of core.clj:185-191; e is the cursor of the exponential
// stream.  tu/te != nullptr: the draws are read from this machine's tapes (NaN once a tape is exhausted).
static __device__ MJP_NOINLINE double sojourn_replay(uint32_t idx, uint32_t consumer, uint64_t gid, uint64_t seed, double rate,
                                             uint32_t *e_io, const double *tu, const double *te, uint32_t len_u,
                                             uint32_t len_e) {
  uint32_t o[4], e = *e_io;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), g0 = (uint32_t)gid, g1 = (uint32_t)(gid >> 32);
  const double QN = __longlong_as_double(0x7FF8000000000000ll);
  double some_num, T;
  if (tu) {
    some_num = idx < len_u ? tu[idx] : QN;                                     // core.clj:185
    T = e < len_e ? te[e] : QN; ++e;                                           // core.clj:187
    while (T == T && some_num == some_num && !(some_num < __dadd_rn(1.0, -det_exp(-__dmul_rn(rate, T))))) {
      T = __dadd_rn(T, e < len_e ? te[e] : QN); ++e;                           // core.clj:190-191
    }
    if (!(some_num == some_num)) T = QN;
  } else {
    philox4x32_10(idx, consumer, g0, g1, k0, k1, o);
    some_num = u53(o[0], o[1]);                                                // core.clj:185
    philox4x32_10(e++, consumer + 1u, g0, g1, k0, k1, o);
    T = __ddiv_rn(-det_log(u52open(o[0], o[1])), rate);                        // core.clj:187
    while (!(some_num < __dadd_rn(1.0, -det_exp(-__dmul_rn(rate, T))))) {      // core.clj:188, 162-165
      philox4x32_10(e++, consumer + 1u, g0, g1, k0, k1, o);
      T = __dadd_rn(T, __ddiv_rn(-det_log(u52open(o[0], o[1])), rate));        // core.clj:190-191
    }
  }
  *e_io = e;
  return T;
}
// the two uniforms of job-type block `blk` (consumer 0)
static __device__ MJP_NOINLINE void jobtype_uniforms(uint32_t blk, uint64_t gid, uint64_t seed, double *u0, double *u1) {
  uint32_t o[4];
  philox4x32_10(blk, 0u, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  *u0 = u53(o[0], o[1]); *u1 = u53(o[2], o[3]);
}

#define MJP_BIT(i) ((MaskT)1 << (i))
#ifndef MJP_COLD_CONST_GLOBAL
#define MJP_COLD_CONST_GLOBAL 1   /* 0: cold layouts copy lam / mu / w/W into shared memory too (A/B builds) */
#endif
#ifndef MJP_FIXED_NT
#define MJP_FIXED_NT 64      /* CTA size of the MT > 0 variants (compile-time [slot][thread] strides) */
#endif

// MT > 0: the line has exactly MT machines and their :ends live in registers (fully unrolled scan);
// MT == 0: any M, :ends in shared memory.
// COLD (wide lines): everything that is not read every micro-step lives in an L2-resident scratch
// [slot][instance] instead of shared memory, so that enough warps fit on an SM.
// SLICE: the launch may stop at a slice boundary and/or resume parked instances (mjp_launch_slice); kept out
// of the plain variants because the park/unpark code keeps every state variable alive (4 % on the headline).
template <typename MaskT, int MT, bool REPLAY, bool HASH, bool SLICE, bool LOG, bool COLD = false>
__global__ void __launch_bounds__(64, line_kernel_min_ctas(sizeof(MaskT), MT, LOG, COLD, REPLAY)) mjp_line_kernel(const KParams p) {
  using MO = MaskOps<MaskT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = (MT > 0) ? MJP_FIXED_NT : (int)blockDim.x, M = (MT > 0) ? MT : p.M, K = p.K;

  // ---- CTA constant table --------------------------------------------------
  // COLD lines of more than 32 machines: lam / mu / w/W are read where they lie in global memory, through the
  // read-only path, by the few lanes that start a job or draw a sojourn -- for 50 machines x 16 job types the table
  // is 7.2 KB of a 35 KB CTA, the difference between 6 and 8 CTAs per SM (mjp_plan.h: plan_const_bytes; measured:
  // C5 quantum 57.2 -> 37.5 ms.  The 32-bit cold kernels gain no CTA from it and lose 1-4 %, so they keep the table)
  constexpr bool CONST_GLOBAL = COLD && sizeof(MaskT) == 8 && MJP_COLD_CONST_GLOBAL;
  double *s_lam = reinterpret_cast<double *>(smem_raw);
  double *s_mu = s_lam + M;
  double *s_dur = s_mu + M;          // [K*M]  w/W  (job-requires, utils.clj:95-101)
  double *c_thr = CONST_GLOBAL ? s_lam : s_dur + K * M;     // [K]    new-job thresholds (core.clj:332-339)
  const double *c_lam = CONST_GLOBAL ? p.lam : s_lam, *c_mu = CONST_GLOBAL ? p.mu : s_mu, *c_dur = CONST_GLOBAL ? p.dur : s_dur;
  int *c_N = reinterpret_cast<int *>(c_thr + K);
  int *c_lvl = c_N + M;
  if (!CONST_GLOBAL) {
    for (int q = tid; q < M; q += nt) { s_lam[q] = p.lam[q]; s_mu[q] = p.mu[q]; }
    for (int q = tid; q < K * M; q += nt) s_dur[q] = p.dur[q];
  }
  for (int q = tid; q < K; q += nt) c_thr[q] = p.thr[q];
  for (int q = tid; q < M - 1; q += nt) { c_N[q] = p.N[q]; c_lvl[q] = p.lvl[q]; }
  __syncthreads();

  // ---- which instances this CTA runs ----
  // Normally CTA b runs instances [b nt, (b+1) nt) from start (or from the park) to the end (or to the slice
  // boundary).  ROTATION (SLICE variants, p.rot_slices > 0; batches a little larger than what fits on the GPU at
  // once): the run is cut into p.rot_slices slices of sim time and the resident CTAs pull (chunk of nt instances,
  // slice) items from a queue -- restore, advance one slice, park -- so that every lane does the same share of the
  // work instead of the last, partly filled wave costing a full pass.  Slice k of a chunk waits for slice k-1 of
  // the same chunk (rot_flags, release/acquire); items are handed out slice-major, so the wait is almost never real.
  __shared__ uint32_t s_rot_item;
  const uint32_t n_inst = (uint32_t)p.n_inst;
  for (;;) {
  const bool rotate = SLICE && p.rot_slices > 0;
  uint32_t chunk = blockIdx.x, rot_k = 0;
  if (rotate) {
    __syncthreads();                                         // the previous item's shared slots are no longer in use
    if (tid == 0) s_rot_item = atomicAdd(p.rot_queue, 1u);
    __syncthreads();
    const uint32_t item = s_rot_item;
    if (item >= p.rot_chunks * (uint32_t)p.rot_slices) break;
    chunk = item % p.rot_chunks; rot_k = item / p.rot_chunks;
    if (rot_k > 0) {
      if (tid == 0) while (ld_acquire_u32(p.rot_flags + chunk) < rot_k) nano_sleep(200);
      __syncthreads();
    }
  }
  const uint32_t g = chunk * (uint32_t)nt + tid;             // instances per call < 2^31 (mjp_upload)
  // Lanes past the end of the batch stay in the warp (the loop below runs in lock-step) but do nothing.
  bool active = g < n_inst;
  // A resumed launch continues exactly the instances that were parked at the previous slice boundary.
  const bool resume = SLICE && (rotate ? rot_k > 0 : p.resume != 0);
  const double slice_until = (rotate && rot_k + 1 < (uint32_t)p.rot_slices) ? p.rot_dt * (double)(rot_k + 1) : p.slice_until;
  if (resume && active && p.status[g] != MJP_INST_PAUSED) active = false;
  const bool fresh_start = active && !resume;
  const bool ran = active;                      // this lane simulates an instance in this launch

  // ---- per-instance state: [slot][thread] in shared memory; the cold part optionally in global ----
  double *sf = reinterpret_cast<double *>(smem_raw + p.const_bytes) + tid;
  uint32_t *su = reinterpret_cast<uint32_t *>(smem_raw + p.const_bytes + (size_t)p.nf64 * nt * 8) + tid;
  double *cfp = COLD ? p.cold_f + g : sf;                   // cold f64 slots and their stride
  uint32_t *cup = COLD ? p.cold_u + g : su;
  const size_t cs = COLD ? (size_t)n_inst : (size_t)nt;
  const int hb = COLD ? 0 : 5 * M - 1;                     // first hot f64 slot in shared memory
  const int hu = COLD ? 0 : 2 * M;                         // first hot u32 slot in shared memory
#define FUT_T(i)   cfp[(i) * cs]                 /* :future etime                      */
#define NXT_T(i)   cfp[(M + (i)) * cs]           /* etime of the event after :future (valid unless `need`) */
#define BS(i)      cfp[(2 * M + (i)) * cs]       /* :bs (NaN = nil), util/log.clj:176  */
#define SS(i)      cfp[(3 * M + (i)) * cs]       /* :ss                                */
#define LASTCLK(b) cfp[(4 * M + (b)) * cs]       /* :lastclk, util/log.clj:151,164     */
#define TICKCLK    sf[(hb) * nt]                /* :clk of the messages in :log-buf   */
#define EJENT      sf[(hb + 1) * nt]            /* :ent of this tick's :ej            */
#define TMINUD     sf[(hb + 2) * nt]            /* earliest pending up/down time (cache) */
  // :ends (MT == 0): with the per-group minima below it is read only when a group is re-scanned, so the widest
  // lines keep it in the cold scratch too when that buys CTAs per SM -- it was over half of their shared memory
  const bool ends_cold = COLD && p.ends_cold != 0;         // decided by the host: when it buys CTAs per SM
  const int hg = hb + 3 + (ends_cold ? 0 : M);             // first group-minimum slot
  double *const ends_base = ends_cold ? cfp + (size_t)(5 * M - 1) * cs : sf + (size_t)(hb + 3) * nt;
  const uint32_t ends_stride = ends_cold ? (uint32_t)cs : (uint32_t)nt;
#define ENDS_S(i)  ends_base[(size_t)(i) * ends_stride]   /* :status :ends, or remaining work (MT == 0 only) */
#define GMINE(gr)  sf[(hg + (gr)) * nt]          /* MT == 0: min :ends over the tracked machines of group gr   */
#define GMINU(gr)  sf[(hg + G + (gr)) * nt]      /* MT == 0: min pending up/down time over group gr            */
#define FUT_N(i)   cup[(i) * cs]                 /* :future index                      */
  // jobs started on machine i: read by every job movement, so a cold layout keeps it in shared memory after the
  // buffer counts whenever the M extra words per instance cost no CTA per SM (p.started_hot, decided by the host)
  const bool started_hot = COLD && p.started_hot != 0;
  uint32_t *const st_base = started_hot ? su + (size_t)(hu + M - 1) * nt : cup + (size_t)M * cs;
  const uint32_t st_stride = started_hot ? (uint32_t)nt : (uint32_t)cs;       // (instances per call < 2^31)
#define STARTED(i) st_base[(size_t)(i) * st_stride]
#define CNT(b)     su[(hu + (b)) * nt]          /* bits 0-7 (count :holding) of buffer b; bits 8-18 this tick's
                                                   occupancy bookkeeping, see buffer_tick_bits; bits 24-31 :N */
  uint32_t *s_jt = cup + (size_t)(COLD ? 2 * M : 3 * M - 1) * cs;  /* job type ring, 4 per word    */
  const int jtw = K > 1 ? (p.R >> 2) : 0;                         /* a single job type needs no ring */
  uint32_t *s_eidx = s_jt + (size_t)jtw * cs;                     /* REPLAY only: exponential cursor */
  const uint32_t Rm = (uint32_t)p.R - 1;

  auto getb = [&](uint32_t *base, size_t stride, int idx) -> uint32_t {
    return (base[(idx >> 2) * stride] >> ((idx & 3) * 8)) & 0xFFu;
  };
  auto setb = [&](uint32_t *base, size_t stride, int idx, uint32_t v) {
    const int sh = (idx & 3) * 8;
    uint32_t wd = base[(idx >> 2) * stride];
    base[(idx >> 2) * stride] = (wd & ~(0xFFu << sh)) | (v << sh);
  };
  // Tick / loop flags share one register (bits 0-7; bits 8.. hold the end status, see finish()):
  //   F_EJ     this tick has an :ej                    F_TICK   TICKCLK == clock
  //   F_OVF    a per-tick structure overflowed         F_PAUSE  stopped at a slice boundary
  //   F_TAPE   a draw tape ran out                     F_DURPL / F_CAPPL  per-instance W|w / N planes are present
  enum { F_EJ = 1, F_TICK = 2, F_OVF = 4, F_PAUSE = 8, F_TAPE = 16, F_DURPL = 32, F_CAPPL = 64 };
  uint32_t flags = ((p.W_i || p.w_i) ? F_DURPL : 0u) | (p.N_i ? F_CAPPL : 0u);
  // line parameters, with the optional per-instance planes
  auto cst = [&](const double *q) -> double { return CONST_GLOBAL ? __ldg(q) : *q; };
  auto lam_of = [&](int i) -> double { return (p.lam_i != nullptr) ? p.lam_i[(size_t)i * n_inst + g] : cst(c_lam + i); };
  auto mu_of = [&](int i) -> double { return (p.mu_i != nullptr) ? p.mu_i[(size_t)i * n_inst + g] : cst(c_mu + i); };
  auto cap_of = [&](int b) -> int { return (flags & F_CAPPL) ? p.N_i[(size_t)b * n_inst + g] : c_N[b]; };
  auto dur_of = [&](int k, int i) -> double {
    if (flags & F_DURPL) {
      if (p.dur_i != nullptr) return p.dur_i[(size_t)(k * M + i) * n_inst + g];   // job-requires, divided once at upload
      const double wv = p.w_i ? p.w_i[(size_t)(k * M + i) * n_inst + g] : p.w[k * M + i];
      const double Wv = p.W_i ? p.W_i[(size_t)i * n_inst + g] : p.W[i];
      return __ddiv_rn(wv, Wv);
    }
    return cst(c_dur + k * M + i);
  };

  double r_end[MT > 0 ? MT : 1] = {};
  auto get_end = [&](int i) -> double {
    if (MT > 0) {
      double v = r_end[0];
#pragma unroll
      for (int j = 1; j < MT; ++j) v = (j == i) ? r_end[j] : v;
      return v;
    }
    return ENDS_S(i);
  };
  auto set_end = [&](int i, double v) {
    if (MT > 0) {
#pragma unroll
      for (int j = 0; j < MT; ++j) if (j == i) r_end[j] = v;
    } else ENDS_S(i) = v;
  };

  // Philox block {idx, consumer, instance id} under key = seed; nothing of it is kept live in registers
  auto px = [&](uint32_t idx, uint32_t consumer, uint32_t o[4]) {
    const uint64_t gid = p.first_id + g;
    philox4x32_10(idx, consumer, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)p.seed, (uint32_t)(p.seed >> 32), o);
  };
  const MaskT ALL = (M >= (int)(8 * sizeof(MaskT))) ? ~(MaskT)0 : (MJP_BIT(M) - 1);
  const MaskT bbs = (MaskT)p.bbs_mask;
  const double INF = __longlong_as_double(0x7FF0000000000000ll);
  const double QNAN = __longlong_as_double(0x7FF8000000000000ll);

  // sojourn after the event (kind_was_up ? :up : :down) numbered n_prev at time t_prev  (core.clj:180-191)
  // `e` is the machine's cursor in its exponential stream (REPLAY only); the caller decides whether to commit it.
  auto next_event_time = [&](int i, uint32_t n_prev, bool prev_is_up, double t_prev, uint32_t &e) -> double {
    const double rate = prev_is_up ? lam_of(i) : mu_of(i);
    const uint64_t gid = p.first_id + g;
    const bool tape = REPLAY && p.tape_u != nullptr;
    const double T = REPLAY ? sojourn_replay(n_prev + 1, 1u + 2u * i, gid, p.seed, rate, &e,
                                             tape ? p.tape_u + ((size_t)g * M + i) * p.tape_len_u : nullptr,
                                             tape ? p.tape_e + ((size_t)g * M + i) * p.tape_len_e : nullptr,
                                             p.tape_len_u, p.tape_len_e)
                            : sojourn_counter(n_prev + 1, 1u + 2u * i, gid, p.seed, rate);
    return __dadd_rn(t_prev, T);                      // NaN when a tape ran out: caught at the use sites
  };

  // ---- preprocess-model: first up/down event of each machine (core.clj:156-160,557-562) ----
  MaskT up = 0;
  for (int i = 0; fresh_start && i < M; ++i) {
    const double lam = lam_of(i), mu = mu_of(i);
    const double p_up = __ddiv_rn(mu, __dadd_rn(lam, mu));
    uint32_t o[4];
    bool is_up;
    double first;                                   // sample-exp 1 :rate (if up? lambda mu), core.clj:160
    if (REPLAY && p.tape_u != nullptr) {
      const double u0 = p.tape_len_u ? p.tape_u[((size_t)g * M + i) * p.tape_len_u] : QNAN;
      first = p.tape_len_e ? p.tape_e[((size_t)g * M + i) * p.tape_len_e] : QNAN;
      if (!(u0 == u0)) first = QNAN;
      is_up = u0 < p_up;
      s_eidx[i * cs] = 1u;
    } else {
      px(0u, 1u + 2u * i, o);
      is_up = u53(o[0], o[1]) < p_up;
      double eu;
      if (REPLAY) {
        uint32_t o2[4];
        px(0u, 2u + 2u * i, o2);
        eu = u52open(o2[0], o2[1]);
        s_eidx[i * cs] = 1u;
      } else {
        eu = u52open(o[2], o[3]);
      }
      first = __ddiv_rn(-det_log(eu), is_up ? lam : mu);
    }
    FUT_T(i) = first;
    FUT_N(i) = 0u;
    STARTED(i) = 0u;
    set_end(i, 0.0);
    BS(i) = QNAN;
    SS(i) = QNAN;
    if (is_up) up |= MJP_BIT(i);
  }
  for (int b = 0; fresh_start && b < M - 1; ++b) LASTCLK(b) = p.warm_up;
  for (int b = 0; fresh_start && b < M - 1; ++b) CNT(b) = (uint32_t)cap_of(b) << 24;   // :N <= 254 (mjp_upload)
  // Wide lines (MT == 0): the two minima a micro-step needs -- the earliest strictly-future :ends among the
  // machines that own an :ends candidate, and the earliest pending up/down time -- are kept per GROUP of GS
  // machines (GMINE / GMINU) and maintained incrementally, so that no pass over all M machines is left on the
  // per-micro-step path: a group is re-read only when its minimum is consumed or a member leaves it.
  // Group size GS = 4 (at most 16 groups): a rescan reads GS values at a few active lanes, the per-micro-step min reads
  // G values at all lanes.  Measured on 10, 20 and 50 machines: groups of 4 beat groups of 8 (by 11 % on 50 machines)
  // and groups of 2 (which cost the 20-machine line a CTA per SM).
  constexpr int GSH = 2, GS = 1 << GSH;
  constexpr uint32_t GSM = (1u << GS) - 1u;
  const int G = (M + GS - 1) >> GSH;
  auto grp_bits = [&](MaskT m, int gr) -> uint32_t { return (uint32_t)(m >> (GS * gr)) & GSM; };
  // min of :ends over the machines `members` (bit j = machine GS gr + j) of group gr
  auto group_min_ends = [&](int gr, uint32_t members) -> double {
    double acc = INF;
#pragma unroll
    for (int j = 0; j < GS; ++j) if ((members >> j) & 1u) { const double e = ENDS_S(GS * gr + j); acc = e < acc ? e : acc; }
    return acc;
  };
  // min over the machines' pending up/down times and who attains it: changes only when one fires (MT > 0)
  MaskT m_udc = 0;
  auto rescan_fut = [&]() {
    double t0 = INF;
    m_udc = 0;
    for (int i = 0; i < M; ++i) {
      const double t = FUT_T(i);
      if (t < t0) { t0 = t; m_udc = 0; }
      if (t == t0) m_udc |= MJP_BIT(i);
    }
    TMINUD = t0;
  };
  if (fresh_start) {
    if (MT > 0) rescan_fut();
    else {
      double t0 = INF;
      for (int gr = 0; gr < G; ++gr) {
        double acc = INF;
        for (int i = GS * gr; i < GS * gr + GS && i < M; ++i) { const double t = FUT_T(i); acc = t < acc ? t : acc; }
        GMINU(gr) = acc; GMINE(gr) = INF;
        if (acc < t0) { t0 = acc; m_udc = 0; }
        if (acc == t0) m_udc |= MJP_BIT(gr);               // MT == 0: m_udc = the GROUPS that attain the minimum
      }
      TMINUD = t0;
    }
  }

  // ---- model state in registers ---------------------------------------------
  double clock = 0.0;
  MaskT occ = 0, B = 0, S = 0, full = 0, pend = 0, dup = 0, fired = 0;
  // MT == 0: `late` (machines whose job is over: occupied, not pending, clock >= :ends) is carried from micro-step
  // to micro-step, and `tracked` is the set of machines whose :ends is accounted for in GMINE (all strictly future)
  MaskT late = 0, tracked = 0;
  MaskT need = ALL;                             // machines whose NXT_T has not been generated yet
  MaskT empty = ALL & ~(MaskT)1;               // feed-buffer-empty? (nil for the entry machine, utils.clj:85)
  const MaskT upj = p.jm2dm ? ALL : (MaskT)0;  // :jm2dm, core.clj:490,495
  uint32_t curjob = 0;
  uint32_t n_ev = 0;                            // run-action calls not yet added to p.n_events
  uint64_t h = 0xcbf29ce484222325ull;
  // tick state (what :log-buf holds, util/log.clj:89-109)
  MaskT seen = 0, lead = 0, touched = 0;       // touched: buffers with a :bj / :sm in this tick
  // :bl/:ub (resp. :st/:us) messages of the tick per machine: parity and "two seen".  A machine cannot
  // toggle three times in one tick while processing times are positive (it would have to finish a job it
  // started in the same tick); a third toggle is reported as MJP_INST_TICK_OVERFLOW.
  MaskT pB = 0, twoB = 0, pS = 0, twoS = 0;
  // The instance ends with this :status.  Kept in bits 8.. of `flags` and stored once after the loop: the
  // rare paths that end an instance are if-converted, and a predicated global store at each of them was
  // ~30 issued instructions per micro-step.
  auto finish = [&](int code) { flags |= (uint32_t)(code + 1) << 8; };   // at most once per instance and launch
  if (REPLAY && fresh_start)
    for (int i = 0; i < M; ++i) if (!(FUT_T(i) == FUT_T(i))) flags |= F_TAPE;   // a tape too short for the first event

  // Job types drawn ahead of the jobs that will use them: bits 0-23 hold up to three types (0xFF = nil) for
  // jobs curjob+1.., bits 24.. their count.  A Philox block yields the types of two consecutive jobs, and the
  // draw runs for every lane of the warp that is down to <= 1 whenever ONE lane has none left (a draw per
  // lane-in-need was 92 % of the warp's Philox executions at 1.7 active lanes).  Which block a job uses
  // depends on its number only, so drawing early changes nothing observable.
  uint32_t tq = 0;
  const bool jt_random = !(K == 1 && c_thr[0] >= 1.0) && !(REPLAY && p.tape_jt != nullptr);

  auto mark_seen = [&](int i) {       // (group-by :m) order, util/log.clj:44 -- branch-free
    // lead bit i: machine i's first message of the tick came before machine i+1's (both masks are zero
    // at tick start and a machine is fresh once per tick, so the bit only ever needs setting)
    const MaskT bit = MJP_BIT(i);
    lead |= bit & ~seen & ~(seen >> 1);
    seen |= bit;
  };
  // buf+ / buf- (util/log.clj:159-171): of the messages a tick puts on a buffer, the FIRST IN CLEANED ORDER
  // gets the elapsed time at its level n, the others add (clk - clk) = 0.  A buffer sees at most one :bj (from
  // machine b) and one :sm (from machine b+1) per tick; cleaned order is by machine -- in order of first
  // appearance while at most 8 machines log in the tick, in hash-map order of the machine names beyond
  // (group-by :m, util/log.clj:44) -- so when both occur the winner is only known when the tick is complete.
  // Kept in the count word: bits 8-15 the level of the FIRST message, bit 16 = it was the :bj, bit 17 = one
  // message so far, bit 18 = two (the second one's level follows: a :sm after a :bj sees one job more, a :bj
  // after a :sm one less).  Stale bits of an older tick are overwritten by the first message of a new one.
  // Lines of up to 8 machines (MT > 0) can never have more than 8 groups: there the winner is decided right away
  // (`precedes`: this message's machine comes first) and bits 8-15 simply hold the winning level.
  auto buffer_tick_bits = [&](uint32_t w, MaskT bbit, int n, bool is_bj, bool precedes) -> uint32_t {
    if (!(touched & bbit)) {
      touched |= bbit;
      return (w & 0xFF000000u) | ((uint32_t)n << 8) | (is_bj ? 0x10000u : 0u) | 0x20000u;
    }
    uint32_t hi = w & 0xFFFFFF00u;
    if ((((hi >> 16) & 1u) != 0) == is_bj || (hi & 0x40000u)) flags |= F_OVF;   // same kind twice, or a third message
    if (MT > 0 && precedes) hi = (hi & 0xFFFF00FFu) | ((uint32_t)n << 8);
    return hi | 0x40000u;
  };
  auto toggle = [&](MaskT &par, MaskT &two, MaskT bit) {
    if (two & bit) flags |= F_OVF;
    two |= par & bit;
    par ^= bit;
  };
  // new-job (core.clj:328-340): job number d+1 uses half (d & 1) of Philox block {d >> 1, consumer 0}
  auto type_from = [&](double choose) -> int {
    for (int k = 0; k < K; ++k) if (choose <= c_thr[k]) return k;
    return -1;
  };
  // the type of the next job, not consumed (-1: none -- the reference then dies in job-requires, core.clj:240-254)
  auto next_type = [&]() -> int {
    if (jt_random) { const int t = (int)(tq & 0xFFu); return t == 0xFF ? -1 : t; }
    if (!(REPLAY && p.tape_jt != nullptr)) return 0;
    if (curjob < p.tape_len_jt) return type_from(p.tape_jt[(size_t)g * p.tape_len_jt + curjob]);   // MJP_RNG_TAPE: rand of new-job, core.clj:331
    flags |= F_TAPE;
    return -1;
  };
  // ---- SCADA capture (LOG): :log-buf of the current tick, staged in an L2-resident scratch ----
  // [instance][slot]: the messages of an instance share cache lines with nobody else, so after the first
  // load of a tick the cleaning loops below run out of L1; rows beyond a tick's few messages are never
  // touched.  The capture costs no shared memory (same residency as a stats-only run).
  // word: see stage_word
  unsigned long long *g_stage = LOG ? p.stage + (size_t)g * (size_t)p.stage_cap : nullptr;
  const int stage_cap = p.stage_cap;
  int n_staged = 0;
  uint32_t line_cnt = 0;                                   // [:report :line-cnt]
  double aj_ends = 0.0;
  auto stage_msg = [&](int act, int m, int n, uint32_t job) {      // log/log util/log.clj:30-36
    // print-now? (util/log.clj:220-226) is constant within a tick: a tick before the warm-up or after
    // :max-lines has been reached is never printed, so there is nothing to stage for it
    if (clock < p.warm_up || (uint64_t)line_cnt >= p.max_lines) return;
    if (n_staged < stage_cap)
      g_stage[n_staged] = stage_word((unsigned)act, (unsigned)m, (unsigned)n, (unsigned)n_staged, job);
    else flags |= F_OVF;
    n_staged++;
  };
  // clean-log-buf (util/log.clj:38-57) + print-lines (util/log.clj:111-137).  The number of messages that
  // survive cleaning follows from the tick masks (a machine's :bl/:ub or :st/:us messages are dropped iff
  // there are exactly two, util/log.clj:51-53), so space is reserved first -- ONE atomicAdd per warp -- and
  // the 32-byte records are written while the cleaned order is being discovered: machines in order of first
  // appearance (group-by :m, util/log.clj:44), within a machine the classes {:bl :ub}, {:st :us}, rest in
  // order of first appearance, within a class the original order.
  auto emit_tick = [&](const bool on, const double tick_clk) {
    constexpr unsigned fm = 0xffffffffu;
    const int lane = tid & 31;
    const double ej_ent = EJENT;
    const bool print_now = on && p.log_on && tick_clk >= p.warm_up && p.max_lines > (uint64_t)line_cnt;   // util/log.clj:220-226
    const int ns = (print_now && !(flags & F_OVF)) ? n_staged : 0;
    const int k = ns ? ns - 2 * (MO::popc(twoB) + MO::popc(twoS)) : 0;
    // warp-aggregated reservation over the lanes that flush in this micro-step: shuffle scan of k
    unsigned prefix = 0, total = 0;
    if (__any_sync(fm, k != 0)) {
      unsigned v = (unsigned)k;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(fm, v, d);
        if (lane >= d) v += t;
      }
      prefix = v - (unsigned)k;
      total = __shfl_sync(fm, v, 31);
    }
    unsigned long long base = 0;
    const int leader = 0;
    if (lane == leader && total) base = atomicAdd(p.log_cursor, (unsigned long long)total);
    base = __shfl_sync(fm, base, leader) + prefix;
    const LogTick lt{reinterpret_cast<mjp_log_record *>(p.log_records), base, p.log_capacity, tick_clk, aj_ends, ej_ent,
                     line_cnt, (uint32_t)g, k};
    unsigned dropped = 0;
    int q = 0;
    // Nearly every tick is grouped already (each machine's messages adjacent, each class within them adjacent):
    // then cleaning keeps the original order and ONE voted pass writes the records.  The pass notices when a
    // tick is not grouped; that tick is then written again by the general, out-of-line routine (same k, same
    // slots).
    MaskT gdone = 0;
    unsigned cdone = 0, pm = 0xFFFFFFFFu, pc = 3u;
    // (a tick in which more than 8 machines logged is regrouped in hash-map order: general routine)
    const bool wide = MT == 0 && MO::popc(seen) > 8;
    bool grouped = !wide;
#if MJP_EMIT_RUNNING_POINTER
    const unsigned long long *sp = g_stage;                  // (a running pointer: the row address is not recomputed per trip)
    for (int x = 0; __any_sync(fm, !wide && x < ns); ++x, ++sp) if (!wide && x < ns) {
      const unsigned long long w = *sp;
#else
    for (int x = 0; __any_sync(fm, !wide && x < ns); ++x) if (!wide && x < ns) {
      const unsigned long long w = g_stage[x];
#endif
      const unsigned act = stage_act(w), m = stage_m(w), c = log_class(act);
      const MaskT mbit = MJP_BIT(m);
      if (m != pm) { if (gdone & mbit) grouped = false; gdone |= mbit; cdone = 0u; pm = m; pc = 3u; }
      if (c != pc) { if ((cdone >> c) & 1u) grouped = false; cdone |= 1u << c; pc = c; }
      if ((c == 0u && (twoB & mbit)) || (c == 1u && (twoS & mbit))) continue;           // util/log.clj:51-53
      const int r = put_record(lt, q, w);
      dropped += r == 1; if (r == 2) flags |= F_OVF;
      ++q;
    }
    if (!grouped && ns) {
      const unsigned r = emit_general<MaskT>(lt, g_stage, ns, twoB, twoS, wide ? p.by_rank : nullptr, M);
      q = (int)(r & 0xFFFFu); dropped = (r >> 16) & 0x7FFFu;
      if (r >> 31) flags |= F_OVF;
    }
    if (ns) {
      if (q != k) flags |= F_OVF;
      if (dropped) atomicAdd(p.log_cursor + 1, (unsigned long long)dropped);
      // (log/detail model) util/log.clj:11-28: nothing before the warm-up reaches the record stream, so the host is
      // told where the jobs were at the END of the tick that produced the instance's first records (the state has not
      // moved since: this is the first logging micro-step after it, before its batch).  Once per instance, out of line.
      if (line_cnt == 0u && k > 0 && p.first_state != nullptr)
        write_line_state(p.first_state + (size_t)g * (size_t)(2 * M), st_base, st_stride, &CNT(0), (uint32_t)nt, (uint64_t)occ, M);
      line_cnt += (uint32_t)k;                               // util/log.clj:137
    }
    if (on) n_staged = 0;
  };
  // push-log for the tick that just ended (util/log.clj:98-108): clean-log-buf + add-compute-log!.
  // Entered by every lane of the warp (`on` = this lane has a tick to flush): its loops are voted over the
  // full warp, lanes with nothing to flush bring empty lists.
  auto flush_tick = [&](const bool on) {
    constexpr unsigned fm = 0xffffffffu;
    const double tick_clk = TICKCLK;
    const bool warm = on && tick_clk >= p.warm_up;                         // util/log.clj:104
    // :bl/:ub (and :st/:us) of a machine: dropped iff exactly two (util/log.clj:51-53); otherwise replayed
    // in order, which nets to: close the interval that was open at tick start (block- / starve-), then open
    // one iff the machine ends the tick blocked (starved).  Both kinds share one voted loop.
    auto interval = [&](int i, bool is_s) {
      double *slot = &cfp[((is_s ? 3 : 2) * M + i) * cs];                       // SS(i) : BS(i)
      const double t0 = *slot;
      // exactly one message of this kind in the tick (two cancel, util/log.clj:51-53): the state flipped
      const bool open1 = (((is_s ? S : B) >> i) & 1) != 0, open0 = !open1;
      if (open0 && t0 == t0) red_add((is_s ? p.starved : p.blocked) + (size_t)i * n_inst + g, tick_clk - t0);
      *slot = open1 ? tick_clk : QNAN;
    };
    if (COLD && MJP_FLUSH_MERGED) {
      // Cold layouts read :bs / :ss / :lastclk from the L2-resident scratch, and with one voted loop per kind every
      // trip waits for its own L2 round trip (ncu: half of the flush's samples are long-scoreboard stalls, the flush
      // is 27-31 % of all samples on 10 and 20 machines).  One loop takes an item of EACH kind per trip: the three
      // loads are issued before the first use, so they are in flight together, and the trip count is the longest
      // list instead of the sum.  Different kinds add into different planes, so the order within each plane -- and
      // with it every rounding -- is unchanged.  (Measured: 20 machines 137.8 -> 130.0 ms, 10 machines 168.6 -> 166.2.)
      const MaskT bj_first_c = (MT == 0 && MO::popc(seen) > 8) ? (MaskT)p.hlead_mask : lead;
      MaskT qb = warm ? pB : (MaskT)0, qs = warm ? pS : (MaskT)0, qt = warm ? touched : (MaskT)0;
      while (__any_sync(fm, (qb | qs | qt) != 0)) if (qb | qs | qt) {
        const bool hb_ = qb != 0, hs_ = qs != 0, ht_ = qt != 0;
        const int ib = hb_ ? MO::ffs(qb) : 0, is_ = hs_ ? MO::ffs(qs) : 0, bt = ht_ ? MO::ffs(qt) : 0;
        qb &= qb - 1; qs &= qs - 1; qt &= qt - 1;             // (an empty list stays empty)
        double t0b = 0.0, t0s = 0.0, lc = 0.0;
        if (hb_) t0b = BS(ib);
        if (hs_) t0s = SS(is_);
        if (ht_) lc = LASTCLK(bt);
        if (hb_) {
          const bool open1 = ((B >> ib) & 1) != 0;
          if (!open1 && t0b == t0b) red_add(p.blocked + (size_t)ib * n_inst + g, tick_clk - t0b);
          BS(ib) = open1 ? tick_clk : QNAN;
        }
        if (hs_) {
          const bool open1 = ((S >> is_) & 1) != 0;
          if (!open1 && t0s == t0s) red_add(p.starved + (size_t)is_ * n_inst + g, tick_clk - t0s);
          SS(is_) = open1 ? tick_clk : QNAN;
        }
        if (ht_) {
          const uint32_t cw = CNT(bt);
          int n = (int)((cw >> 8) & 0xFFu);
          if (MT == 0 && (cw & 0x40000u)) {
            const bool first_bj = (cw & 0x10000u) != 0, bj_wins = ((bj_first_c >> bt) & 1) != 0;
            n += (first_bj == bj_wins) ? 0 : (first_bj ? 1 : -1);
          }
          red_add(p.occ + (size_t)(c_lvl[bt] + n) * n_inst + g, tick_clk - lc);
          LASTCLK(bt) = tick_clk;
        }
      }
    } else {
    if (MT > 0) {                                             // M <= 8: both masks fit one 32-bit word
      uint32_t q = warm ? ((uint32_t)pB | ((uint32_t)pS << 16)) : 0u;
      while (__any_sync(fm, q != 0)) if (q) {
        const int j = __ffs((int)q) - 1; q &= q - 1;
        interval(j & 15, j >= 16);
      }
    } else if (sizeof(MaskT) == 4) {
      uint64_t q = warm ? ((uint64_t)pB | ((uint64_t)pS << 32)) : 0ull;
      while (__any_sync(fm, q != 0)) if (q) {
        const int j = __ffsll((long long)q) - 1; q &= q - 1;
        interval(j & 31, j >= 32);
      }
    } else {
      for (MaskT q = warm ? pB : (MaskT)0; __any_sync(fm, q != 0);) if (q) { const int i = MO::ffs(q); q &= q - 1; interval(i, false); }
      for (MaskT q = warm ? pS : (MaskT)0; __any_sync(fm, q != 0);) if (q) { const int i = MO::ffs(q); q &= q - 1; interval(i, true); }
    }
    // buf+ / buf-: one RED per touched buffer, at the level recorded by buffer_tick_bits.  When the buffer saw
    // both a :bj and a :sm, the one whose machine comes first in cleaned order counts (see buffer_tick_bits).
    const MaskT bj_first = (MT == 0 && MO::popc(seen) > 8) ? (MaskT)p.hlead_mask : lead;
    for (MaskT q = warm ? touched : (MaskT)0; __any_sync(fm, q != 0);) if (q) {
      const int b = MO::ffs(q); q &= q - 1;
      const uint32_t cw = CNT(b);
      int n = (int)((cw >> 8) & 0xFFu);
      if (MT == 0 && (cw & 0x40000u)) {
        const bool first_bj = (cw & 0x10000u) != 0, bj_wins = ((bj_first >> b) & 1) != 0;
        n += (first_bj == bj_wins) ? 0 : (first_bj ? 1 : -1);
      }
      const double lc = LASTCLK(b);
      red_add(p.occ + (size_t)(c_lvl[b] + n) * n_inst + g, tick_clk - lc);
      LASTCLK(b) = tick_clk;
    }
    }
    if (on && (flags & F_EJ)) {
      if (warm) {                                             // end-job util/log.clj:212-217
        red_add(p.residence + g, tick_clk - EJENT);
        atomicAdd(reinterpret_cast<unsigned long long *>(p.njobs) + g, 1ull);
      }
      // n_ev counts the events since it was last folded into p.n_events; it is folded whenever a job has
      // left the line, so 3e9 events without a single exit means the instance is stuck (watchdog above)
      if (n_ev > 0x40000000u) {
        atomicAdd(reinterpret_cast<unsigned long long *>(p.n_events) + g, (unsigned long long)n_ev); n_ev = 0;
      }
    }
    if (LOG) emit_tick(on, tick_clk);
    if (on) {
      seen = lead = touched = 0;
      pB = twoB = pS = twoS = 0;
      flags &= ~(uint32_t)F_EJ;
    }
  };
  // MT == 0: a job whose :ends v has just become known is entered into its group's minimum right away -- nearly
  // always the machine owns an :ends candidate in the next micro-step; when it does not (a :BBS machine behind a
  // full buffer) the membership check at the top of the loop takes it out again and re-reads the group
  auto track = [&](int i, double v) {
    if (v > clock) {
      const int gr = i >> GSH;
      const double cur = GMINE(gr);
      GMINE(gr) = v < cur ? v : cur;
      tracked |= MJP_BIT(i);
    }
  };
  // end-time at job start (core.clj:138-151), lazily: fits in the current up interval, or pending
  auto start_job = [&](int i, int type) {
    const MaskT bit = MJP_BIT(i);
    double v;
    if (end_time_begin((up & bit) != 0, FUT_T(i), clock, dur_of(type, i), v)) pend |= bit; else pend &= ~bit;
    set_end(i, v);
    if (MT == 0) {
      late &= ~bit;                                         // a new job: not over
      if (!(pend & bit)) track(i, v);
    }
  };

  // ---- time slicing: the complete per-instance state, parked in HBM as [word][instance] ----
  // One routine moves every loop-carried variable and every shared-memory slot in a fixed order, so saving
  // and restoring cannot disagree.  (COLD state and the :enters ring already live in global memory.)
  // (Values are passed and returned, never referenced: taking the address of a loop-carried variable would
  // pin it in local memory for the whole kernel.)
#define MJP_PARK_LIST(D, MK, U, Q)                                                                          \
  D(clock)                                                                                                  \
  MK(up) MK(occ) MK(B) MK(S) MK(full) MK(empty) MK(pend) MK(dup) MK(fired) MK(need)                         \
  MK(seen) MK(lead) MK(touched) MK(pB) MK(twoB) MK(pS) MK(twoS) MK(m_udc) MK(late) MK(tracked) \
  U(curjob) U(n_ev) U(flags) U(tq)   /* (flags is parked before the :status code is put into it) */
  // state that exists only in some variants (parking it unconditionally would keep it alive everywhere)
#define MJP_PARK_LOG(D, U) D(aj_ends) U(line_cnt) U(n_staged)
  auto park_save = [&]() {
    uint32_t *ps = p.park + g;
    size_t wd = 0;
    auto w32 = [&](uint32_t v) { ps[wd * n_inst] = v; ++wd; };
    auto w64 = [&](uint64_t v) { w32((uint32_t)v); w32((uint32_t)(v >> 32)); };
#define PD(x) w64(d2bits(x));
#define PM(x) if (sizeof(MaskT) == 8) w64((uint64_t)(x)); else w32((uint32_t)(x));
#define PU(x) w32((uint32_t)(x));
#define PQ(x) w64((uint64_t)(x));
    MJP_PARK_LIST(PD, PM, PU, PQ)
    if (LOG) { MJP_PARK_LOG(PD, PU) }
    if (HASH) { PQ(h) }
#undef PD
#undef PM
#undef PU
#undef PQ
#pragma unroll
    for (int j = 0; j < (MT > 0 ? MT : 0); ++j) w64(d2bits(r_end[j]));
    for (int k = 0; k < p.nf64; ++k) w64(d2bits(sf[(size_t)k * nt]));
    for (int k = 0; k < p.nu32; ++k) w32(su[(size_t)k * nt]);
  };
  if (resume && active) {
    const uint32_t *ps = p.park + g;
    size_t wd = 0;
    auto r32 = [&]() -> uint32_t { const uint32_t v = ps[wd * n_inst]; ++wd; return v; };
    auto r64 = [&]() -> uint64_t { const uint64_t lo = r32(); const uint64_t hi = r32(); return lo | (hi << 32); };
#define GD(x) x = bits2d(r64());
#define GM(x) x = (MaskT)(sizeof(MaskT) == 8 ? r64() : (uint64_t)r32());
#define GU(x) x = (decltype(x))r32();
#define GQ(x) x = r64();
    MJP_PARK_LIST(GD, GM, GU, GQ)
    if (LOG) { MJP_PARK_LOG(GD, GU) }
    if (HASH) { GQ(h) }
#undef GD
#undef GM
#undef GU
#undef GQ
#pragma unroll
    for (int j = 0; j < (MT > 0 ? MT : 0); ++j) r_end[j] = bits2d(r64());
    for (int k = 0; k < p.nf64; ++k) sf[(size_t)k * nt] = bits2d(r64());
    for (int k = 0; k < p.nu32; ++k) su[(size_t)k * nt] = r32();
  }
#undef MJP_PARK_LIST
#undef MJP_PARK_LOG

  // ======================= main-loop-loop, core.clj:729-744 =======================
  // One micro-step per trip with the warp in lock-step.  Nothing leaves the body early: a lane that
  // ends (or errs, or lies past the end of the batch) empties its batch and rides along until the whole
  // warp is done, so every section can be a loop voted over the FULL warp -- __any_sync is the
  // reconvergence point (ncu showed 3.9 of 32 lanes per instruction without it), and with a constant
  // full mask it is one VOTE instruction instead of four (REDUX + R2UR + BRA.DIV + VOTE, +8 %).
  constexpr unsigned am = 0xffffffffu;
  for (;;) {
    if (!__any_sync(am, active)) break;
    bool live = active;
    // end-main-loop? core.clj:722-727
    if (active) {
      if (!((p.run_to_job >= 0 && (long long)curjob <= p.run_to_job) || (clock <= p.run_to))) {
        finish(MJP_INST_NORMAL_END); live = false;
      } else if (n_ev > 0xC0000000u) { finish(MJP_INST_ITER_LIMIT); live = false; }   // watchdog, see flush_tick
      else if (SLICE && clock > slice_until) { flags |= F_PAUSE; live = false; }   // slice boundary, between two micro-steps
    }

    // new-job core.clj:328-340: keep at least one drawn job type per lane (see tq); here, before the guards,
    // few values are live across the out-of-line call
    if (jt_random && __any_sync(am, live && (tq >> 24) == 0u)) {
      const uint32_t cnt = tq >> 24;
      if (live && cnt <= 1u) {
        double u0, u1;
        jobtype_uniforms((curjob + cnt) >> 1, p.first_id + g, p.seed, &u0, &u1);   // curjob + cnt is even
        const uint32_t t0 = (uint32_t)type_from(u0) & 0xFFu, t1 = (uint32_t)type_from(u1) & 0xFFu;
        tq = (tq & ((1u << (8u * cnt)) - 1u)) | (t0 << (8u * cnt)) | (t1 << (8u * cnt + 8u)) | ((cnt + 2u) << 24);
      }
    }

    // ---- runables (core.clj:515-531): guards on the pre-batch state ----
    const MaskT notfull = ~full, nonempty = ~empty;
    const MaskT c_aj = (~occ & ~B & up & (~bbs | notfull)) & (MaskT)1;                       // core.clj:343-354
    const MaskT c_tm = (up | upj) & ~(MaskT)1 & ~B & ~S & ~occ & nonempty & (~bbs | notfull) & ALL;   // core.clj:485-506
    const MaskT c_us = S & nonempty;                                                          // core.clj:462-470
    const MaskT c_ub = B & notfull & ALL;                                                     // core.clj:426-449
    const MaskT c_blb = bbs & ~B & ~occ & nonempty & full;     // BBS, time = clock          core.clj:387-394,406-411
    const MaskT c_bla = ~bbs & ~B & full & occ & ~pend & ALL;  // BAS, time = :ends          core.clj:376-383,401-405
    const MaskT c_tb = occ & ~B & notfull & ~pend & ALL;       // time = max(clock, :ends)   core.clj:472-483

    // Machines whose job is already over (clock >= :ends).  For them max(clock, :ends) = clock, so their
    // advance2buffer candidate is a "now" candidate; a BAS new-blocked candidate that is late has
    // :ends == clock (an occupied, unblocked BAS machine always owns a candidate at :ends, so the clock
    // cannot pass it).  Only the strictly-future :ends need a min.
    // The same pass takes the min over the future candidates: pending up/down events (cached) and the
    // strictly-future :ends of the machines that own an :ends candidate (c_tb | c_bla, all occupied and not pending).
    const MaskT cand = c_tb | c_bla;
    const double tmin_ud = TMINUD;
    double tmin = tmin_ud;
    double te = INF;                                         // MT == 0: earliest tracked :ends
    uint32_t ge_arg = 0;                                     // MT == 0: groups that attain it
    if (MT > 0) {                                            // :ends in registers: the late mask first ...
      late = 0;
#pragma unroll
      for (int i = 0; i < MT; ++i) if (clock >= r_end[i]) late |= MJP_BIT(i);
      late &= occ & ~pend;
    } else {
      // (a) a job that is over on a machine WITHOUT an :ends candidate (a :BBS machine working into a full buffer)
      // passes unnoticed; new-starved? (finished?, utils.clj:68-74) is the one guard that can see it
      for (MaskT q = live ? (occ & ~pend & ~late & ~cand & empty & ~S) : (MaskT)0; __any_sync(am, q != 0);) if (q) {
        const int i = MO::ffs(q); q &= q - 1;
        if (clock >= ENDS_S(i)) late |= MJP_BIT(i);
      }
      // (b) machines that joined or left the candidates since the last micro-step (a job started or was resumed;
      // a buffer filled up or drained behind a :BBS machine)
      const MaskT fut0 = cand & ~late;
      uint32_t gdirty = 0;
      for (MaskT q = live ? (tracked & ~fut0) : (MaskT)0; __any_sync(am, q != 0);) if (q) {
        const int gr = MO::ffs(q) >> GSH;
        q &= ~((MaskT)GSM << (GS * gr));
        gdirty |= 1u << gr;                                  // a member left: the group's minimum is re-read below
      }
      for (MaskT q = live ? (fut0 & ~tracked) : (MaskT)0; __any_sync(am, q != 0);) if (q) {
        const int i = MO::ffs(q); q &= q - 1;
        const double e = ENDS_S(i);
        if (clock >= e) late |= MJP_BIT(i);                  // already over: a "now" candidate, never tracked
        else { const int gr = i >> GSH; const double cur = GMINE(gr); GMINE(gr) = e < cur ? e : cur; }
      }
      if (live) tracked = cand & ~late;
      for (uint32_t q = gdirty; __any_sync(am, q != 0);) if (q) {
        const int gr = __ffs((int)q) - 1; q &= q - 1;
        GMINE(gr) = group_min_ends(gr, grp_bits(tracked, gr));
      }
      // (c) the earliest of them: G group minima instead of M machines
      for (int gr = 0; gr < G; ++gr) {
        const double v = GMINE(gr);
        if (v < te) { te = v; ge_arg = 0u; }
        if (v == te) ge_arg |= 1u << gr;
      }
      tmin = te < tmin ? te : tmin;
    }
    const MaskT c_st = ~S & empty & (~occ | late) & ALL;       // finished? = no job or clock >= :ends   core.clj:451-460, utils.clj:68-74
    const MaskT tb_now = c_tb & late, bl_now = c_bla & late, fut = cand & ~late;
    const bool now_any = (c_aj | c_tm | c_st | c_us | c_ub | c_blb | dup | tb_now | bl_now) != 0;
    if (MT > 0) {                                            // ... then the min (two short passes are cheaper here)
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const double e = (fut & MJP_BIT(i)) ? r_end[i] : INF;
        tmin = e < tmin ? e : tmin;
      }
    }
    // (the paths that end an instance sit behind a vote each: as predicated code their operands cost ~17 issued
    // instructions per micro-step, ncu)
    const bool stuck = live && (now_any ? (tmin < clock) : (tmin == INF));
    if (__any_sync(am, stuck)) {
      if (stuck) {
        finish(now_any ? MJP_INST_CLOCK_BACK : MJP_INST_NO_RUNABLES);                         // core.clj:133,739
        live = false;
      }
    }
    const bool no_type = live && now_any && c_aj != 0 && next_type() < 0;
    if (__any_sync(am, no_type)) if (no_type) {
      // add-job (a "now" candidate, first action of its batch) has no job type to draw: the reference dies in
      // it before logging anything -- the previous tick is never pushed, nothing of this batch is executed
      finish((flags & F_TAPE) ? MJP_INST_TAPE_EXHAUSTED : MJP_INST_JOBTYPE_NIL); live = false;
    }
    const double tstar = now_any ? clock : tmin;
    MaskT m_f = 0;                                            // future :ends equal to t*
    MaskT m_ud = 0;                                           // pending up/down events at t*
    if (MT > 0) {
      if (!now_any) {
#pragma unroll
        for (int i = 0; i < MT; ++i) if ((fut & MJP_BIT(i)) && r_end[i] == tstar) m_f |= MJP_BIT(i);
      }
      m_ud = (tmin_ud == tstar) ? (m_udc & ~dup) : (MaskT)0;
    } else {
      // the groups whose minimum is t*: their members at t* fire (and leave the tracked set: the clock has
      // reached their :ends), the others give the group's new minimum -- one pass does both
      for (uint32_t q = (live && !now_any && te == tstar) ? ge_arg : 0u; __any_sync(am, q != 0);) if (q) {
        const int gr = __ffs((int)q) - 1; q &= q - 1;
        const uint32_t members = grp_bits(tracked, gr);
        uint32_t hit = 0;
        double acc = INF;
#pragma unroll
        for (int j = 0; j < GS; ++j) if ((members >> j) & 1u) {
          const double e = ENDS_S(GS * gr + j);
          if (e == tstar) hit |= 1u << j; else acc = e < acc ? e : acc;
        }
        GMINE(gr) = acc;
        m_f |= (MaskT)hit << (GS * gr);
      }
      tracked &= ~m_f;
      late |= m_f;                                            // (every use of the pre-advance `late` is above)
      // same for the up/down times: members of the groups at t* that fire, new group minimum over the others
      // (the fired machines add their next event time below)
      uint32_t uq = (live && tmin_ud == tstar) ? (uint32_t)m_udc : 0u;
      for (; __any_sync(am, uq != 0);) if (uq) {
        const int gr = __ffs((int)uq) - 1; uq &= uq - 1;
        const uint32_t is_dup = grp_bits(dup, gr);
        uint32_t hit = 0;
        double acc = INF;
#pragma unroll
        for (int j = 0; j < GS; ++j) if (GS * gr + j < M) {
          const double t = FUT_T(GS * gr + j);
          if (t == tstar && !((is_dup >> j) & 1u)) hit |= 1u << j; else acc = t < acc ? t : acc;
        }
        GMINU(gr) = acc;
        m_ud |= (MaskT)hit << (GS * gr);
      }
    }
    const MaskT keep = live ? ~(MaskT)0 : (MaskT)0, keep_now = (live && now_any) ? ~(MaskT)0 : (MaskT)0;
    MaskT b_aj = c_aj & keep_now;
    MaskT b_tm = c_tm & keep_now, b_st = c_st & keep_now, b_us = c_us & keep_now, b_ub = c_ub & keep_now;
    MaskT b_bl = ((m_f & c_bla) & keep) | ((c_blb | bl_now) & keep_now);
    MaskT b_ud = (m_ud & keep) | (dup & keep_now);
    MaskT b_tb = ((m_f & c_tb) & keep) | (tb_now & keep_now);

    // ---- advance-clock core.clj:128-135 ----
    if (live) {
      if (tstar != clock) { fired = 0; flags &= ~(uint32_t)F_TICK; }
      clock = tstar;
    }
    const bool will_log = live && (p.up_down ? true : ((b_aj | b_tb | b_tm | b_st | b_us | b_bl | b_ub) != 0));
    {
      const bool do_flush = will_log && seen != 0 && !(flags & F_TICK);   // push-log of the previous tick
      flush_tick(do_flush);
      if (will_log && !(flags & F_TICK)) { TICKCLK = clock; flags |= F_TICK; }
    }
    if (MT > 0)                                               // M <= 8: two packed words instead of eight counts
      n_ev += __popc((uint32_t)b_st | ((uint32_t)b_us << 8) | ((uint32_t)b_bl << 16) | ((uint32_t)b_ub << 24)) +
              __popc((uint32_t)b_aj | ((uint32_t)b_tb << 8) | ((uint32_t)b_tm << 16) | ((uint32_t)b_ud << 24));
    else
      n_ev += MO::popc(b_aj) + MO::popc(b_ud) + MO::popc(b_tb) + MO::popc(b_tm) + MO::popc(b_st) +
              MO::popc(b_us) + MO::popc(b_bl) + MO::popc(b_ub);

    // ---- run the batch in generator order, core.clj:742 ----
    if (__any_sync(am, b_aj != 0)) {
      if (b_aj) {                                            // add-job core.clj:240-254
        const int type = next_type();                        // >= 0: checked when the batch was formed
        if (jt_random) tq = ((tq >> 8) & 0xFFFFu) | (((tq >> 24) - 1u) << 24);
        const uint32_t id = curjob + 1;
        start_job(0, type);
        if (HASH) { hmix(h, (uint64_t)MJP_ACT_AJ | ((uint64_t)type << 16) | ((uint64_t)id << 32)); hmix(h, d2bits(clock)); hmix(h, 0); }
        mark_seen(0);
        if (LOG) {
          stage_msg(MJP_ACT_AJ, 0, type, id);
          // the :aj record carries :ends (core.clj:250): when the lazy end-time is still pending, run the
          // literal look-ahead of core.clj:141-151 over events that are generated but not committed
          if (!(pend & 1)) aj_ends = get_end(0);
          else {
            double rem = get_end(0), t = FUT_T(0), v;
            uint32_t n = FUT_N(0), ecur = REPLAY ? s_eidx[0] : 0u;
            bool have_nxt = !(need & 1);                     // event n+1 is already realised in NXT_T(0)
            auto after = [&](bool prev_is_up, double t_prev) -> double {
              double r;
              if (have_nxt) { r = NXT_T(0); have_nxt = false; }
              else r = next_event_time(0, n, prev_is_up, t_prev, ecur);
              ++n;
              return r;
            };
            if (up & 1) t = after(false, t);                 // the :up after the pending :down
            for (;;) {
              const double t_dn = after(true, t);
              if (!(t_dn == t_dn)) { v = QNAN; flags |= F_TAPE; break; }       // tape exhausted
              if (!end_time_resume(t, t_dn, rem, v)) break;
              rem = v;
              t = after(false, t_dn);
            }
            aj_ends = v;
          }
        }
        curjob = id; occ |= 1; STARTED(0) = id;
        if (K > 1) setb(s_jt, cs, (int)(id & Rm), (uint32_t)type);
        p.enters[(size_t)g * (Rm + 1u) + (id & Rm)] = clock;
        if (fired & 1) { dup |= 1; up ^= 1; }              // machine-future core.clj:194-198
      }
    }
    {
      // Pre-generate the event AFTER each machine's pending one (core.clj:180-191), lazily: the Philox + ln
      // path is ~200 instructions and only ~5 % of the events need it, so it waits until one lane of the warp
      // cannot wait any longer -- its pending event fires in this micro-step -- and then runs for every lane
      // that has something to generate.
      const MaskT need_l = live ? need : (MaskT)0;           // (a lane that has just parked or ended touches nothing)
      if (__any_sync(am, (b_ud & need_l & ~dup) != 0)) {
        for (MaskT q = need_l; __any_sync(am, q != 0);) if (q) {
          const int i = MO::ffs(q); q &= q - 1;
          const bool pending_is_up = !((up ^ dup) & MJP_BIT(i));
          uint32_t ecur = REPLAY ? s_eidx[i * cs] : 0u;
          NXT_T(i) = next_event_time(i, FUT_N(i), pending_is_up, FUT_T(i), ecur);
          if (REPLAY) s_eidx[i * cs] = ecur;
        }
        if (live) need = 0;
      }
    }
    {                                                        // bring-machine-up, then -down: core.clj:215-238
      // M <= 8: both masks are one 32-bit list (ups in bits 0-7, downs in bits 8-15)
      const MaskT dup_before = dup;
      uint32_t q32 = MT > 0 ? ((uint32_t)(b_ud & ~up) | ((uint32_t)(b_ud & up) << 8)) : 0u;
      MaskT q_up = MT > 0 ? (MaskT)0 : (b_ud & ~up), q_dn = MT > 0 ? (MaskT)0 : (b_ud & up);
      while (__any_sync(am, MT > 0 ? q32 != 0u : (q_up | q_dn) != 0)) if (MT > 0 ? q32 != 0u : (q_up | q_dn) != 0) {
        int i;
        MaskT bit;
        if (MT > 0) {
          const int j = __ffs((int)q32) - 1; q32 &= q32 - 1u;
          i = j & 7; bit = MJP_BIT(i);
        } else {
          const bool from_up = q_up != 0;                    // (no reference-select: it would pin both masks in local memory)
          const MaskT qq = from_up ? q_up : q_dn;
          bit = qq & (~qq + 1);
          i = MO::ffs(qq);
          q_up &= from_up ? ~bit : ~(MaskT)0;
          q_dn &= from_up ? ~(MaskT)0 : ~bit;
        }
        const bool was_up_event = !(up & bit);
        if (HASH) {
          hmix(h, (uint64_t)(was_up_event ? MJP_ACT_UP : MJP_ACT_DOWN) | ((uint64_t)i << 8));
          hmix(h, d2bits(clock)); hmix(h, 0);
        }
        if (p.up_down) { mark_seen(i); if (LOG) stage_msg(was_up_event ? MJP_ACT_UP : MJP_ACT_DOWN, i, 0, 0u); }
        fired |= bit;
        if (dup & bit) { dup &= ~bit; up ^= bit; }           // re-fire of an already realised event
        else {
          const double t_fire = FUT_T(i);
          const uint32_t n = FUT_N(i);
          const double t_next = NXT_T(i);                  // generated above at the latest
          if (REPLAY && !(t_next == t_next)) flags |= F_TAPE;
          need |= bit;
          FUT_T(i) = t_next; FUT_N(i) = n + 1; up ^= bit;
          if (was_up_event && (pend & bit)) {                // end-time resumes at this :up  core.clj:145-151
            double v;
            const bool still = end_time_resume(t_fire, t_next, get_end(i), v);
            if (!still) pend &= ~bit;
            set_end(i, v);
            if (MT == 0 && !still) track(i, v);
          }
          if (MT > 0) rescan_fut();
          else { const int gr = i >> GSH; const double cur = GMINU(gr); GMINU(gr) = t_next < cur ? t_next : cur; }
        }
      }
      if (MT == 0 && __any_sync(am, (b_ud & ~dup_before) != 0)) {
        if (b_ud & ~dup_before) {                            // an event fired: the earliest pending up/down time anew
          double t0 = INF;
          for (int gr = 0; gr < G; ++gr) {
            const double v = GMINU(gr);
            if (v < t0) { t0 = v; m_udc = 0; }
            if (v == t0) m_udc |= MJP_BIT(gr);
          }
          TMINUD = t0;
        }
      }
    }
    for (MaskT q = b_tb; __any_sync(am, q != 0);) if (q) {   // advance2buffer core.clj:285-300
      const int i = MO::ffs(q); q &= q - 1;
      const MaskT bit = MJP_BIT(i);
      const uint32_t id = STARTED(i);
      if (i < M - 1) {
        const uint32_t cw = CNT(i);
        const int n = (int)(cw & 0xFFu);
        if (HASH) { hmix(h, (uint64_t)MJP_ACT_BJ | ((uint64_t)i << 8) | ((uint64_t)n << 16) | ((uint64_t)id << 32)); hmix(h, d2bits(clock)); hmix(h, d2bits(get_end(i))); }
        mark_seen(i);
        if (LOG) stage_msg(MJP_ACT_BJ, i, n, id);
        CNT(i) = (uint32_t)(n + 1) | buffer_tick_bits(cw, bit, n, true, ((lead >> i) & 1) != 0);
        if (n + 1 == (int)(cw >> 24)) full |= bit;
        empty &= ~(bit << 1);
      } else {
        const double ent = p.enters[(size_t)g * (Rm + 1u) + (id & Rm)];
        if (HASH) { hmix(h, (uint64_t)MJP_ACT_EJ | ((uint64_t)i << 8) | ((uint64_t)id << 32)); hmix(h, d2bits(clock)); hmix(h, d2bits(ent)); hmix(h, d2bits(get_end(i))); }
        mark_seen(i);
        if (LOG) stage_msg(MJP_ACT_EJ, i, 0, id);
        if (flags & F_EJ) flags |= F_OVF;
        flags |= F_EJ; EJENT = ent;
      }
      occ &= ~bit;
      if (MT == 0) late &= ~bit;
    }
    for (MaskT q = b_tm; __any_sync(am, q != 0);) if (q) {   // advance2machine core.clj:263-277
      const int i = MO::ffs(q); q &= q - 1;
      const MaskT bit = MJP_BIT(i), bbit = bit >> 1;
      const int b = i - 1;
      const uint32_t cw = CNT(b);
      const int n = (int)(cw & 0xFFu);
      const uint32_t id = STARTED(i) + 1;
      const int type = K > 1 ? (int)getb(s_jt, cs, (int)(id & Rm)) : 0;
      start_job(i, type);
      if (HASH) { hmix(h, (uint64_t)MJP_ACT_SM | ((uint64_t)i << 8) | ((uint64_t)n << 16) | ((uint64_t)id << 32)); hmix(h, d2bits(clock)); hmix(h, 0); }
      mark_seen(i);
      if (LOG) stage_msg(MJP_ACT_SM, i, n, id);
      CNT(b) = (uint32_t)(n - 1) | buffer_tick_bits(cw, bbit, n, false, ((lead >> b) & 1) == 0);
      if (n - 1 == 0) empty |= bit;
      full &= ~bbit;
      occ |= bit; STARTED(i) = id;
      if (fired & bit) { dup |= bit; up ^= bit; }
    }
    if (MT > 0) {                                            // record-actions core.clj:304-326, in the order
      // new-starved, -unstarved, -blocked, -unblocked; M <= 8: the four masks are one 32-bit list
      uint32_t q = (uint32_t)b_st | ((uint32_t)b_us << 8) | ((uint32_t)b_bl << 16) | ((uint32_t)b_ub << 24);
      while (__any_sync(am, q != 0)) if (q) {
        const int j = __ffs((int)q) - 1; q &= q - 1;
        const int which = j >> 3, i = j & 7;
        const MaskT bit = MJP_BIT(i);
        if (HASH) { hmix(h, (uint64_t)(MJP_ACT_ST + which) | ((uint64_t)i << 8)); hmix(h, d2bits(clock)); hmix(h, 0); }
        mark_seen(i);
        if (LOG) stage_msg(MJP_ACT_ST + which, i, 0, 0u);
        if (which < 2) { S ^= bit; toggle(pS, twoS, bit); }
        else { B ^= bit; toggle(pB, twoB, bit); }
      }
    } else {
      MaskT q_st = b_st, q_us = b_us, q_bl = b_bl, q_ub = b_ub;
      while (__any_sync(am, (q_st | q_us | q_bl | q_ub) != 0)) if (q_st | q_us | q_bl | q_ub) {
        const int which = q_st ? 0 : (q_us ? 1 : (q_bl ? 2 : 3));
        const MaskT qq = which == 0 ? q_st : (which == 1 ? q_us : (which == 2 ? q_bl : q_ub));
        const MaskT bit = qq & (~qq + 1);
        const int i = MO::ffs(qq);
        q_st &= which == 0 ? ~bit : ~(MaskT)0;
        q_us &= which == 1 ? ~bit : ~(MaskT)0;
        q_bl &= which == 2 ? ~bit : ~(MaskT)0;
        q_ub &= which == 3 ? ~bit : ~(MaskT)0;
        if (HASH) { hmix(h, (uint64_t)(MJP_ACT_ST + which) | ((uint64_t)i << 8)); hmix(h, d2bits(clock)); hmix(h, 0); }
        mark_seen(i);
        if (LOG) stage_msg(MJP_ACT_ST + which, i, 0, 0u);
        if (which < 2) { S ^= bit; toggle(pS, twoS, bit); }
        else { B ^= bit; toggle(pB, twoB, bit); }
      }
    }
    const bool broken = live && ((flags & (F_TAPE | F_OVF)) != 0 || seen == 0);
    if (__any_sync(am, broken)) if (broken) {                             // rare: one vote on the common path
      finish((flags & F_TAPE) ? MJP_INST_TAPE_EXHAUSTED
                              : ((flags & F_OVF) ? MJP_INST_TICK_OVERFLOW
                                                 : MJP_INST_LOGBUF_EMPTY));   // push-log returned nil, util/log.clj:95
      live = false;
    }
    active = live;
  }
  if (ran) {
    atomicAdd(reinterpret_cast<unsigned long long *>(p.n_events) + g, (unsigned long long)n_ev);
    n_ev = 0;
    if (SLICE && (flags & F_PAUSE)) {               // park the untouched state (the lane idled since the boundary)
      flags &= ~(uint32_t)F_PAUSE;
      park_save();
      finish(MJP_INST_PAUSED);
    }
    if (flags >> 8) p.status[g] = (uint8_t)((flags >> 8) - 1);

    // the reference never flushes the last tick: whatever is still in :log-buf is lost
    if (p.final_clock) p.final_clock[g] = clock;
    if (p.curjob) p.curjob[g] = (long long)curjob;
    if (HASH && p.hash) p.hash[g] = h;
  }
  if (!rotate) break;
  __threadfence();                                           // this slice's park / status / scratch before the flag
  __syncthreads();
  if (tid == 0) st_release_u32(p.rot_flags + chunk, rot_k + 1);
  }   // for (;;) over work items
#undef FUT_T
#undef ENDS_S
#undef GMINE
#undef GMINU
#undef TMINUD
#undef NXT_T
#undef TICKCLK
#undef EJENT
#undef CNT
#undef BS
#undef SS
#undef LASTCLK
#undef FUT_N
#undef STARTED
}

}  // namespace mjp
