/*
 * mjp_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Literal restatement of the MJPdes run loop.  Every function cites the
 * reference lines it follows; paths are relative to
 * /root/reference/src/gov/nist/MJPdes/ ("core" = core.clj, "log" = util/log.clj,
 * "utils" = util/utils.clj).  The data structures deliberately mirror the
 * reference's (a growing realised :up&down seq per machine, a :log-buf vector
 * of message structs, FIFO :holding vectors, candidate-action lists rebuilt
 * every micro-step) so that each step can be read against the Clojure.
 *
 * Parity: pinned on the reference tests' deterministic goldens, unpinned at the
 * random-draw level (see mjp_oracle.h).  Known, documented deviations:
 *   D1  (closed in round 2) util/log.clj:44 `group-by :m` is insertion-ordered only up to
 *       8 keys (array-map); a tick in which MORE than 8 machines log is regrouped in
 *       Clojure hash-map order of the machine keywords.  That order is restated below
 *       (clj_keyword_hash / clj_trie_key, from the published Clojure 1.9 sources: no JVM
 *       here to execute them; (hash :a) = -2123407586 is the pinned known value).
 *       Machine names come from the descriptor's group_rank, default :m1 .. :mM.
 *   D2  a micro-step whose minimum candidate time is ##Inf ends the instance
 *       with MJP_INST_NO_RUNABLES (the reference would set :clock ##Inf, fire
 *       every such action and report :normal-end).  Not reachable on a serial line.
 *   D3  Math/pow(Math/E, x) at core.clj:165 is computed as exp(x) by mjpo_exp.
 */
#include "mjp_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <thread>
#include <vector>

namespace {

const double kInf = std::numeric_limits<double>::infinity();
const double kNaN = std::numeric_limits<double>::quiet_NaN();

inline uint64_t d2b(double x) { uint64_t b; std::memcpy(&b, &x, 8); return b; }
inline double b2d(uint64_t b) { double x; std::memcpy(&x, &b, 8); return x; }

/* ------------------------------------------------------------------------ */
/* Clojure map iteration order (clojure 1.9.0, project.clj:11).              */
/* clojure.lang.Murmur3.hashUnencodedChars / Util.hashCombine / Symbol.hasheq */
/* / Keyword.hasheq; PersistentHashMap iterates by 5-bit hash chunks, lowest  */
/* first.  ASCII names only.                                                  */
/* ------------------------------------------------------------------------ */
inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint32_t clj_murmur3_chars(const char *name, size_t n) {
  uint32_t h1 = 0;
  for (size_t i = 1; i < n; i += 2) {
    uint32_t k1 = (uint32_t)(unsigned char)name[i - 1] | ((uint32_t)(unsigned char)name[i] << 16);
    k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;             /* mixK1 */
    h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5u + 0xe6546b64u;             /* mixH1 */
  }
  if (n & 1) {
    uint32_t k1 = (uint32_t)(unsigned char)name[n - 1];
    k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;
    h1 ^= k1;
  }
  h1 ^= (uint32_t)(2 * n);                                                 /* fmix(h1, 2 * length) */
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}
uint32_t clj_keyword_hash(const char *name, size_t n) {                    /* keyword without namespace */
  uint32_t seed = clj_murmur3_chars(name, n);
  /* Util.hashCombine(seed, Util.hash(null) = 0): seed ^= hash + 0x9e3779b9 + (seed << 6) + (seed >> 2) */
  uint32_t sym = seed ^ (0u + 0x9e3779b9u + (seed << 6) + (uint32_t)((int32_t)seed >> 2));
  return sym + 0x9e3779b9u;                                                /* Keyword.hasheq */
}
uint64_t clj_trie_key(uint32_t h) {
  uint64_t k = 0;
  for (int sft = 0; sft < 35; sft += 5) k = (k << 5) | ((h >> sft) & 31u);
  return k;
}

/* ------------------------------------------------------------------------ */
/* Substream arithmetic (the contract in include/mjpdes_b200.h, "random      */
/* numbers").  Philox4x32-10: Salmon, Moraes, Dror, Shaw, SC'11.            */
/* ------------------------------------------------------------------------ */
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ln(x) for normal x > 0: x = 2^k m, m in [sqrt(1/2), sqrt 2), s = (m-1)/(m+1),
 * ln m = 2s (1 + s^2/3 + s^4/5 + ...), eleven terms (|s| <= 0.1716).  Only
 * + - * / and fma, so the CUDA twin is bit-identical.                        */
double det_log(double x) {
  uint64_t b = d2b(x);
  int k = (int)(b >> 52) - 1023;
  uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
  double m;
  if (mant >= 0x6A09E667F3BCDull) { k += 1; m = b2d(mant | (1022ull << 52)); }
  else                            {          m = b2d(mant | (1023ull << 52)); }
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double q = 1.0 / 21.0;
  q = std::fma(q, z, 1.0 / 19.0);
  q = std::fma(q, z, 1.0 / 17.0);
  q = std::fma(q, z, 1.0 / 15.0);
  q = std::fma(q, z, 1.0 / 13.0);
  q = std::fma(q, z, 1.0 / 11.0);
  q = std::fma(q, z, 1.0 / 9.0);
  q = std::fma(q, z, 1.0 / 7.0);
  q = std::fma(q, z, 1.0 / 5.0);
  q = std::fma(q, z, 1.0 / 3.0);
  double two_s = 2.0 * s;
  double dk = (double)k;
  double lo = std::fma(two_s * z, q, dk * b2d(0x3DEA39EF35793C76ull)); /* k*ln2_lo */
  return dk * b2d(0x3FE62E42FEE00000ull) + (two_s + lo);               /* k*ln2_hi */
}

/* e^x for x <= 0.  Below -40 the result cannot change 1.0 - e^x, return 0.   */
double det_exp(double x) {
  if (x < -40.0) return 0.0;
  double kf = std::floor(x * b2d(0x3FF71547652B82FEull) + 0.5);       /* x/ln2 */
  double r = std::fma(-kf, b2d(0x3FE62E42FEE00000ull), x);
  r = std::fma(-kf, b2d(0x3DEA39EF35793C76ull), r);
  double p = 1.0 / 87178291200.0;              /* 1/14! */
  p = std::fma(p, r, 1.0 / 6227020800.0);      /* 1/13! */
  p = std::fma(p, r, 1.0 / 479001600.0);
  p = std::fma(p, r, 1.0 / 39916800.0);
  p = std::fma(p, r, 1.0 / 3628800.0);
  p = std::fma(p, r, 1.0 / 362880.0);
  p = std::fma(p, r, 1.0 / 40320.0);
  p = std::fma(p, r, 1.0 / 5040.0);
  p = std::fma(p, r, 1.0 / 720.0);
  p = std::fma(p, r, 1.0 / 120.0);
  p = std::fma(p, r, 1.0 / 24.0);
  p = std::fma(p, r, 1.0 / 6.0);
  p = std::fma(p, r, 0.5);
  p = std::fma(p, r, 1.0);
  p = std::fma(p, r, 1.0);
  int k = (int)kf;
  return p * b2d((uint64_t)(1023 + k) << 52);
}

struct Draws {
  uint64_t seed, inst;
  void block(uint32_t consumer, uint32_t idx, uint32_t out[4]) const {
    uint32_t ctr[4] = {idx, consumer, (uint32_t)inst, (uint32_t)(inst >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key, out);
  }
  /* clojure.core/rand: [0,1) */
  static double u53(uint32_t a, uint32_t b) {
    uint64_t k = (((uint64_t)a << 32) | b) >> 11;
    return (double)k * 0x1.0p-53;
  }
  /* argument of -ln(): strictly inside (0,1) */
  static double u52open(uint32_t a, uint32_t b) {
    uint64_t k = (((uint64_t)a << 32) | b) >> 12;
    return ((double)k + 0.5) * 0x1.0p-52;
  }
  double uniform(uint32_t consumer, uint32_t idx) const {
    uint32_t o[4]; block(consumer, idx, o); return u53(o[0], o[1]);
  }
  double open_uniform(uint32_t consumer, uint32_t idx) const {
    uint32_t o[4]; block(consumer, idx, o); return u52open(o[0], o[1]);
  }
};

/* ------------------------------------------------------------------------ */
/* The :up&down lazy seq of one machine (core:156-192).                      */
/* ------------------------------------------------------------------------ */
enum { EV_UP = 0, EV_DOWN = 1 };
struct Ev { uint32_t n; uint8_t kind; double t; };

struct UpDown {
  double lambda = 0, mu = 1;
  int mach = 0, mode = MJP_RNG_COUNTER;
  const Draws *d = nullptr;
  bool finite = false;          /* literal seq, or '([0 :down ##Inf]) for lambda == 0 */
  std::vector<Ev> seq;          /* realised prefix; the head is retained, as in the reference */
  uint32_t u_idx = 0, e_idx = 0;
  /* MJP_RNG_TAPE: recorded draws of this machine (what rand / sample-exp returned in a JVM run) */
  const double *tape_u = nullptr, *tape_e = nullptr;
  uint32_t len_u = 0, len_e = 0;
  bool exhausted = false;
  /* recording (what a JVM user does with with-redefs on rand / sample-exp): every draw as returned */
  double *rec_u = nullptr, *rec_e = nullptr; uint32_t rec_len_u = 0, rec_len_e = 0, rec_nu = 0, rec_ne = 0;
  double note_u(double v) { if (rec_u && rec_nu < rec_len_u) rec_u[rec_nu] = v; ++rec_nu; return v; }
  double note_e(double v) { if (rec_e && rec_ne < rec_len_e) rec_e[rec_ne] = v; ++rec_ne; return v; }
  double tape_uniform() { if (u_idx < len_u) return tape_u[u_idx++]; exhausted = true; ++u_idx; return kNaN; }
  double tape_exp() { if (e_idx < len_e) return tape_e[e_idx++]; exhausted = true; ++e_idx; return kNaN; }

  uint32_t cu() const { return 1u + 2u * (uint32_t)mach; }
  uint32_t ce() const { return 2u + 2u * (uint32_t)mach; }

  /* sample-exp 1 :rate r  ==  -ln(U)/r  (incanter -> Colt Exponential) */
  double exp_draw(double rate) { return note_e(-det_log(d->open_uniform(ce(), e_idx++)) / rate); }

  /* expo-up&down-init-event core:156-160, then expo-up&down core:174-179 */
  void init() {
    double u, e;
    double p_up = mu / (lambda + mu);
    if (mode == MJP_RNG_TAPE) {
      u = tape_uniform();
      bool up = u < p_up;
      e = tape_exp();
      if (exhausted) e = kNaN;
      seq.push_back({0, (uint8_t)(up ? EV_DOWN : EV_UP), e});
    } else if (mode == MJP_RNG_REPLAY) {
      u = note_u(d->uniform(cu(), u_idx++));
      bool up = u < p_up;
      e = exp_draw(up ? lambda : mu);
      seq.push_back({0, (uint8_t)(up ? EV_DOWN : EV_UP), e});
    } else {
      uint32_t o[4]; d->block(cu(), 0, o);
      u = Draws::u53(o[0], o[1]);
      bool up = u < p_up;
      e = -det_log(Draws::u52open(o[2], o[3])) / (up ? lambda : mu);
      seq.push_back({0, (uint8_t)(up ? EV_DOWN : EV_UP), e});
    }
    finite = (lambda == 0.0);   /* core:178-179: the whole seq is the first event */
  }

  /* one step of the `iterate` fn, core:180-191 */
  void realise_next() {
    Ev prev = seq.back();
    bool up_now = (prev.kind == EV_UP);
    double rate = up_now ? lambda : mu;
    double T;
    if (mode == MJP_RNG_TAPE) {
      double some_num = tape_uniform();                            /* core:185 */
      T = tape_exp();                                              /* core:187 */
      while (!exhausted && !(some_num < 1.0 - det_exp(-(rate * T))))
        T = T + tape_exp();                                        /* core:190-191 */
      if (exhausted) T = kNaN;
    } else if (mode == MJP_RNG_REPLAY) {
      double some_num = note_u(d->uniform(cu(), u_idx++));         /* core:185 */
      T = exp_draw(rate);                                          /* core:187 */
      while (!(some_num < 1.0 - det_exp(-(rate * T))))             /* core:188, p-state-change core:162-165 */
        T = T + exp_draw(rate);                                    /* core:190-191 */
    } else {
      uint32_t o[4]; d->block(cu(), prev.n + 1, o);
      double u1 = Draws::u52open(o[0], o[1]), u2 = Draws::u52open(o[2], o[3]);
      T = -det_log(u1 * u2) / rate;                                /* Erlang-2, SURVEY a7 */
    }
    seq.push_back({prev.n + 1, (uint8_t)(up_now ? EV_DOWN : EV_UP), prev.t + T});
  }

  /* (nth up&down k), nullptr past the end of a finite seq */
  const Ev *at(size_t k) {
    while (seq.size() <= k) {
      if (finite || exhausted) return nullptr;     /* a dry tape ends the seq; the loop reports it */
      realise_next();
    }
    return &seq[k];
  }
};

/* ------------------------------------------------------------------------ */
/* Model state (core:38-73, 578-615)                                         */
/* ------------------------------------------------------------------------ */
struct Job { int type = -1; int64_t id = 0; double enters = 0, starts = 0, ends = 0; };

struct Machine {
  double W = 1.0;
  int bbs = 0;
  bool occ = false;      /* :status non-nil */
  Job job;
  bool has_future = true;
  Ev future{0, EV_DOWN, 0};
  UpDown ud;
  size_t ge = 0;         /* memo of the drop-while at core:143,198 (clock is monotone) */
};

struct Msg {             /* one log message map */
  uint8_t act = 0; int m = 0; int64_t j = 0; int n = 0; double aux = 0.0; double clk = 0.0;
  Msg() = default;
  Msg(uint8_t act_, int m_, int64_t j_, int n_, double aux_, double clk_) : act(act_), m(m_), j(j_), n(n_), aux(aux_), clk(clk_) {}
  int ord = 0;           /* its index in :log-buf among the messages of its tick (execution order) */
  /* :dets (log/detail, log:11-28), flattened: the M values of :run (-1 = nil), then per buffer (count, ids...) */
  std::shared_ptr<std::vector<int64_t>> dets;
};

struct Steady {          /* log:144-152 */
  double residence_sum = 0.0; int64_t njobs = 0;
  std::vector<double> blocked, starved, bs, ss;    /* bs/ss: NaN = nil */
  std::vector<std::vector<double>> occ;            /* occ[b][n] */
  std::vector<double> lastclk;
};

enum { FN_AJ, FN_UP, FN_DOWN, FN_TOBUF, FN_TOMACH, FN_ST, FN_US, FN_BL, FN_UB };
struct Act { double time; int fn; int m; Job job; };

struct Sim {
  int M = 0, K = 0;
  std::vector<Machine> mach;
  std::vector<int> N;
  std::vector<std::deque<Job>> holding;
  std::vector<double> portion, w;      /* w[k*M+i] */
  std::vector<uint8_t> blocked, starved;
  double clock = 0.0, warm_up = 0, run_to = 0;
  int64_t run_to_job = -1, current_job = 0;
  bool jm2dm = false, log_on = false, up_down = false;
  uint64_t max_lines = 0, line_cnt = 0;
  std::vector<Msg> log_buf;
  Steady steady;
  Draws draws;
  uint32_t jt_idx = 0;
  const double *tape_jt = nullptr; uint32_t len_jt = 0; bool tape_out = false;
  double *rec_jt = nullptr; uint32_t rec_len_jt = 0;
  int status = MJP_INST_NOT_RUN;
  /* instrumentation */
  uint64_t n_events = 0, hash = 0xcbf29ce484222325ull, iterations = 0, ticks = 0;
  uint64_t act_counts[10] = {0};
  uint32_t max_tick_msgs = 0, max_lookahead = 0;
  uint64_t wide_ticks = 0;            /* pushed post-warm-up ticks in which more than 8 machines logged */
  uint64_t wide_ticks_reordered = 0;  /* ... whose cleaned batch differs from the first-appearance grouping */
  std::vector<int> group_rank;        /* hash-map rank of machine i's name (log:44 beyond 8 groups) */
  /* SCADA sink */
  std::vector<mjp_log_record> *sink = nullptr;
  uint32_t inst_index = 0;
  /* [:report :job-detail?] (log:14): messages carry :dets; the printed ones are appended to dets_sink */
  bool job_detail = false;
  std::vector<int64_t> *dets_sink = nullptr;
  /* where the jobs were after the tick that is printed first (mjp_log_out.first_state) */
  std::vector<int64_t> started;       /* number of the last job started on machine i */
  uint32_t *first_state = nullptr;    /* ... = the state at the END of the tick whose records are printed first */
  std::vector<uint32_t> tick_end_state;
  bool have_first = false;

  /* ---- utils.clj ---- */
  bool up(int i) const { return mach[i].has_future && mach[i].future.kind == EV_DOWN; }   /* utils:60-62 */
  bool down(int i) const { return mach[i].has_future && mach[i].future.kind == EV_UP; }   /* utils:64-66 */
  bool occupied(int i) const { return mach[i].occ; }                                       /* utils:76-80 */
  bool finished(int i) const { return !mach[i].occ || clock >= mach[i].job.ends; }         /* utils:68-74 */
  bool feed_buffer_empty(int i) const { return i != 0 && holding[i - 1].empty(); }         /* utils:82-87 */
  bool buffer_full(int i) const {                                                          /* utils:89-93 */
    return i < M - 1 && (int)holding[i].size() == N[i];
  }
  double job_requires(const Job &j, int i) const { return w[(size_t)j.type * M + i] / mach[i].W; } /* utils:95-101 */

  /* ---- event hash: every run-action call, in execution order ---- */
  /* FNV-1a style over 64-bit words: {act|m<<8|n<<16|j<<32, bits(clk), bits(aux)}; aux is the
   * completed job's :ends for :bj, :ent for :ej (which also mixes :ends), 0 otherwise, so every
   * end-time result is checked at the moment the job leaves its machine. */
  void count_event(int act, int m, int64_t j, int n, double clk, double aux, double aux2 = 0.0) {
    n_events++; act_counts[act]++;
    uint64_t w0 = (uint64_t)act | ((uint64_t)(uint8_t)m << 8) | ((uint64_t)(uint16_t)n << 16) |
                  ((uint64_t)(uint32_t)j << 32);
    hash = (hash ^ w0) * 0x100000001b3ull;
    hash = (hash ^ d2b(clk)) * 0x100000001b3ull;
    hash = (hash ^ (act == MJP_ACT_BJ || act == MJP_ACT_EJ ? d2b(aux) : 0ull)) * 0x100000001b3ull;
    if (act == MJP_ACT_EJ) hash = (hash ^ d2b(aux2)) * 0x100000001b3ull;
  }

  /* ---- log/detail, log:11-28: which job is on which machine / in which buffer, BEFORE the action is applied ---- */
  std::shared_ptr<std::vector<int64_t>> detail() const {
    if (!job_detail) return nullptr;                          /* log:14 */
    auto d = std::make_shared<std::vector<int64_t>>();
    for (int i = 0; i < M; ++i) d->push_back(mach[i].occ ? mach[i].job.id : -1);   /* (-> ... :status :id), log:19-22 */
    for (int b = 0; b < M - 1; ++b) {                                              /* (mapv :id :holding), log:25-28 */
      d->push_back((int64_t)holding[b].size());
      for (const Job &j : holding[b]) d->push_back(j.id);
    }
    return d;
  }
  /* ---- log/log, log:30-36 ---- */
  void log(Msg msg) {
    bool keep = up_down || !(msg.act == MJP_ACT_UP || msg.act == MJP_ACT_DOWN);
    if (!keep) return;
    msg.dets = detail();
    if (first_state && !have_first && (log_buf.empty() || msg.clk != log_buf.back().clk)) {
      /* first message of a new tick, logged before its action is applied: the jobs are where the previous tick
       * left them.  Kept until push-log prints that tick (below). */
      tick_end_state.assign(2 * (size_t)M, 0u);
      for (int i = 0; i < M; ++i) {
        tick_end_state[i] = (uint32_t)started[i];
        tick_end_state[M + i] = (mach[i].occ ? 1u : 0u) | (i < M - 1 ? (uint32_t)holding[i].size() << 8 : 0u);
      }
    }
    log_buf.push_back(msg);
  }

  /* first event with etime >= clock: the drop-while of core:143 and core:198 */
  size_t first_ge(int i) {
    Machine &mc = mach[i];
    for (;;) {
      const Ev *e = mc.ud.at(mc.ge);
      if (!e || !(e->t < clock)) return mc.ge;
      mc.ge++;
    }
  }

  /* end-time, core:138-151 */
  double end_time(double dur, int i) {
    double not_achieved = dur, now = clock;
    size_t k = first_ge(i);
    uint32_t depth = 0;
    for (;;) {
      const Ev *e = mach[i].ud.at(k);
      if (!e) return kNaN;                       /* reference: NPE on an exhausted literal seq */
      if (k > mach[i].future.n) { if (++depth > max_lookahead) max_lookahead = depth; }
      if (e->kind == EV_UP) { now = e->t; k++; }                          /* core:145-146 */
      else if ((e->t - now) > not_achieved) return now + not_achieved;    /* core:147-148 */
      else { not_achieved = not_achieved - (e->t - now); now = e->t; k++; } /* core:149-151 */
    }
  }

  /* machine-future, core:194-198 */
  void set_machine_future(int i) {
    const Ev *e = mach[i].ud.at(first_ge(i));
    if (e) { mach[i].future = *e; mach[i].has_future = true; } else mach[i].has_future = false;
  }

  /* ---- actions, core:215-326 ---- */
  void bring_machine(int i, int act) {                   /* core:215-238 */
    Ev f = mach[i].future;
    count_event(act, i, 0, 0, f.t, 0.0);
    log({(uint8_t)act, i, 0, 0, kNaN, f.t});
    const Ev *nx = mach[i].ud.at((size_t)f.n + 1);       /* first with index > n */
    if (nx) mach[i].future = *nx; else mach[i].has_future = false;
  }
  void add_job(Job j, int i) {                           /* core:240-254 */
    double ends = end_time(job_requires(j, i), i);
    j.enters = clock; j.starts = clock; j.ends = ends;
    count_event(MJP_ACT_AJ, i, j.id, j.type, clock, ends);
    log({MJP_ACT_AJ, i, j.id, j.type, ends, clock});
    current_job++;
    mach[i].occ = true; mach[i].job = j; started[i] = j.id;
    set_machine_future(i);
  }
  void advance2machine(int i) {                          /* core:263-277 */
    std::deque<Job> &h = holding[i - 1];
    Job j = h.front();
    j.starts = clock;
    j.ends = end_time(job_requires(j, i), i);
    count_event(MJP_ACT_SM, i, j.id, (int)h.size(), clock, j.ends);
    log({MJP_ACT_SM, i, j.id, (int)h.size(), kNaN, clock});
    mach[i].occ = true; mach[i].job = j; started[i] = j.id;
    set_machine_future(i);
    h.pop_front();
  }
  void advance2buffer(int i) {                           /* core:285-300 */
    Job j = mach[i].job;
    if (i < M - 1) {
      count_event(MJP_ACT_BJ, i, j.id, (int)holding[i].size(), clock, j.ends);
      log({MJP_ACT_BJ, i, j.id, (int)holding[i].size(), kNaN, clock});
    } else {
      count_event(MJP_ACT_EJ, i, j.id, 0, clock, j.enters, j.ends);
      log({MJP_ACT_EJ, i, j.id, 0, j.enters, clock});
    }
    mach[i].occ = false;
    if (i < M - 1) holding[i].push_back(j);
  }
  void record_action(int act, int i) {                   /* core:304-326 */
    count_event(act, i, 0, 0, clock, 0.0);
    log({(uint8_t)act, i, 0, 0, kNaN, clock});
    switch (act) {
      case MJP_ACT_BL: blocked[i] = 1; break;
      case MJP_ACT_UB: blocked[i] = 0; break;
      case MJP_ACT_ST: starved[i] = 1; break;
      case MJP_ACT_US: starved[i] = 0; break;
    }
  }

  /* new-job, core:328-340 (thresholds p0, 2p0, 2p0+p1, ... : SURVEY Q4) */
  Job new_job() {
    /* draw #d of consumer 0 is half (d & 1) of Philox block {d >> 1, 0} (include/mjpdes_b200.h) */
    uint32_t d = jt_idx++, o[4];
    double choose;
    if (tape_jt) {
      if (d < len_jt) choose = tape_jt[d]; else { choose = kNaN; tape_out = true; }
    } else {
      draws.block(0, d >> 1, o);
      choose = (d & 1) ? Draws::u53(o[2], o[3]) : Draws::u53(o[0], o[1]);
      if (rec_jt && d < rec_len_jt) rec_jt[d] = choose;
    }
    int type = -1;
    double accum = portion[0];
    for (int k = 0; k < K; ++k) {
      if (choose <= accum) { type = k; break; }
      accum = accum + portion[k];
    }
    Job j; j.type = type; j.id = current_job + 1;
    return j;
  }

  /* blocked-check-time, core:396-411 */
  double blocked_check_time(int i) const {
    if (!mach[i].bbs) return (buffer_full(i) && occupied(i)) ? mach[i].job.ends : kInf;
    return (!occupied(i) && !feed_buffer_empty(i) && buffer_full(i)) ? clock : kInf;
  }

  /* runables, core:515-531; the nine generators of core:343-506 in concat order */
  void runables(std::vector<Act> &all, std::vector<Act> &out) {
    all.clear(); out.clear();
    /* add-job? core:343-354 */
    if (!occupied(0) && !blocked[0] && up(0) && (!mach[0].bbs || !buffer_full(0)))
      all.push_back({clock, FN_AJ, 0, new_job()});
    /* bring-machine-up? core:356-363 */
    for (int i = 0; i < M; ++i) if (down(i)) all.push_back({mach[i].future.t, FN_UP, i, Job()});
    /* bring-machine-down? core:365-372 */
    for (int i = 0; i < M; ++i) if (up(i)) all.push_back({mach[i].future.t, FN_DOWN, i, Job()});
    /* advance2buffer? core:472-483 */
    for (int i = 0; i < M; ++i)
      if (occupied(i) && !blocked[i] && !buffer_full(i))
        all.push_back({std::fmax(clock, mach[i].job.ends), FN_TOBUF, i, Job()});
    /* advance2machine? core:485-506 */
    for (int i = 0; i < M; ++i)
      if ((up(i) || jm2dm) && i != 0 && !blocked[i] && !starved[i] && !occupied(i) &&
          !feed_buffer_empty(i) && (!mach[i].bbs || !buffer_full(i)))
        all.push_back({clock, FN_TOMACH, i, Job()});
    /* new-starved? core:451-460 */
    for (int i = 0; i < M; ++i)
      if (!starved[i] && finished(i) && feed_buffer_empty(i)) all.push_back({clock, FN_ST, i, Job()});
    /* new-unstarved? core:462-470 */
    for (int i = 0; i < M; ++i)
      if (starved[i] && !feed_buffer_empty(i)) all.push_back({clock, FN_US, i, Job()});
    /* new-blocked? core:413-424 with new-blocked-bas?/bbs? core:376-394; time via get-time core:508-513 */
    for (int i = 0; i < M; ++i) {
      bool nb = !mach[i].bbs ? (!blocked[i] && buffer_full(i) && occupied(i))
                             : (!blocked[i] && !occupied(i) && !feed_buffer_empty(i));
      if (nb) all.push_back({blocked_check_time(i), FN_BL, i, Job()});
    }
    /* new-unblocked? core:426-449 */
    for (int i = 0; i < M; ++i)
      if (blocked[i] && !buffer_full(i)) all.push_back({clock, FN_UB, i, Job()});
    if (all.empty()) return;
    double min_time = kInf;                                    /* core:530 */
    for (const Act &a : all) min_time = std::fmin(min_time, a.time);
    for (const Act &a : all) if (a.time == min_time) out.push_back(a);   /* core:531 */
  }

  void run_action(const Act &a) {                              /* core:533-536 */
    switch (a.fn) {
      case FN_AJ: add_job(a.job, a.m); break;
      case FN_UP: bring_machine(a.m, MJP_ACT_UP); break;
      case FN_DOWN: bring_machine(a.m, MJP_ACT_DOWN); break;
      case FN_TOBUF: advance2buffer(a.m); break;
      case FN_TOMACH: advance2machine(a.m); break;
      case FN_ST: record_action(MJP_ACT_ST, a.m); break;
      case FN_US: record_action(MJP_ACT_US, a.m); break;
      case FN_BL: record_action(MJP_ACT_BL, a.m); break;
      case FN_UB: record_action(MJP_ACT_UB, a.m); break;
    }
  }

  /* ---- clean-log-buf, log:38-57 ---- */
  static int msg_class(const Msg &m) {
    if (m.act == MJP_ACT_BL || m.act == MJP_ACT_UB) return 0;   /* :bl-ub */
    if (m.act == MJP_ACT_ST || m.act == MJP_ACT_US) return 1;   /* :st-us */
    return 2;                                                   /* :other */
  }
  /* by_hash: order the machine groups as a PersistentHashMap of the machine keywords iterates */
  void clean_log_buf(const std::vector<Msg> &acts, std::vector<Msg> &out, bool by_hash) const {
    out.clear();
    std::vector<int> m_order;                                   /* (group-by :m): first appearance ... */
    for (const Msg &a : acts) {
      bool seen = false;
      for (int m : m_order) if (m == a.m) { seen = true; break; }
      if (!seen) m_order.push_back(a.m);
    }
    if (by_hash)                                                /* ... or, from the 9th key on, hash-map order */
      std::sort(m_order.begin(), m_order.end(), [&](int a, int b) { return group_rank[a] < group_rank[b]; });
    for (int m : m_order) {
      int c_order[3], nc = 0;
      std::vector<Msg> grp[3];
      for (const Msg &a : acts) {
        if (a.m != m) continue;
        int c = msg_class(a);
        if (grp[c].empty()) c_order[nc++] = c;
        grp[c].push_back(a);
      }
      for (int q = 0; q < nc; ++q) {
        int c = c_order[q];
        if (grp[c].size() == 2 && (c == 0 || c == 1)) continue;  /* log:51-53 */
        for (const Msg &a : grp[c]) out.push_back(a);
      }
    }
  }
  static int distinct_machines(const std::vector<Msg> &acts) {
    uint64_t seen = 0;
    for (const Msg &a : acts) seen |= 1ull << a.m;
    return __builtin_popcountll(seen);
  }

  /* ---- add-compute-log!, log:71-86 with buf+ ... end-job log:159-217 ---- */
  void add_compute_log(const Msg &o) {
    Steady &r = steady;
    switch (o.act) {
      case MJP_ACT_BJ: {                                        /* buf+ log:159-164 */
        int b = o.m;
        r.occ[b][o.n] = r.occ[b][o.n] + (o.clk - r.lastclk[b]); r.lastclk[b] = o.clk; break; }
      case MJP_ACT_SM: {                                        /* buf- log:166-171 */
        int b = o.m - 1;
        r.occ[b][o.n] = r.occ[b][o.n] + (o.clk - r.lastclk[b]); r.lastclk[b] = o.clk; break; }
      case MJP_ACT_EJ:                                          /* end-job log:212-217 */
        r.residence_sum = r.residence_sum + (o.clk - o.aux); r.njobs++; break;
      case MJP_ACT_BL: r.bs[o.m] = o.clk; break;                /* block+ log:173-176 */
      case MJP_ACT_UB:                                          /* block- log:185-192 */
        if (!std::isnan(r.bs[o.m])) { r.blocked[o.m] = r.blocked[o.m] + (o.clk - r.bs[o.m]); r.bs[o.m] = kNaN; }
        break;
      case MJP_ACT_ST: r.ss[o.m] = o.clk; break;                /* starve+ log:194-197 */
      case MJP_ACT_US:                                          /* starve- log:203-210 */
        if (!std::isnan(r.ss[o.m])) { r.starved[o.m] = r.starved[o.m] + (o.clk - r.ss[o.m]); r.ss[o.m] = kNaN; }
        break;
      default: break;                                           /* :aj :up :down log:84-86 */
    }
  }

  /* print-lines log:111-137: hand the cleaned batch to the record sink */
  void print_lines(const std::vector<Msg> &clean) {
    if (sink) {
      uint64_t line = line_cnt;
      for (const Msg &m : clean) {
        mjp_log_record r;
        r.clk = m.clk; r.aux = m.aux; r.job = (uint32_t)m.j; r.instance = inst_index;
        r.line = (uint32_t)line++; r.n = (uint8_t)m.n; r.ord = (uint8_t)m.ord; r.act = m.act; r.machine = (uint8_t)m.m;
        sink->push_back(r);
        if (dets_sink && m.dets) dets_sink->insert(dets_sink->end(), m.dets->begin(), m.dets->end());
      }
    }
    line_cnt += clean.size();                                   /* log:137 */
  }

  /* push-log, log:89-109.  false = the reference returned nil */
  bool push_log() {
    if (log_buf.empty()) return false;                          /* log:95 */
    double min_time = kInf, max_time = 0;
    for (const Msg &m : log_buf) { min_time = std::fmin(min_time, m.clk); max_time = std::fmax(max_time, m.clk); }
    if ((uint32_t)log_buf.size() > max_tick_msgs) max_tick_msgs = (uint32_t)log_buf.size();
    if (max_time > min_time) {                                  /* log:98 */
      std::vector<Msg> now, later, clean;
      for (const Msg &m : log_buf) (m.clk == min_time ? now : later).push_back(m);
      for (size_t q = 0; q < now.size(); ++q) now[q].ord = (int)q;
      const bool by_hash = distinct_machines(now) > 8;          /* array-map -> hash-map promotion (D1) */
      clean_log_buf(now, clean, by_hash);                       /* log:100-102 */
      if (by_hash && min_time >= warm_up) {
        std::vector<Msg> first_app;
        clean_log_buf(now, first_app, false);
        wide_ticks++;
        bool same = first_app.size() == clean.size();
        for (size_t q = 0; same && q < clean.size(); ++q)
          same = clean[q].m == first_app[q].m && clean[q].act == first_app[q].act;
        if (!same) wide_ticks_reordered++;
      }
      if (min_time >= warm_up) for (const Msg &m : clean) add_compute_log(m);   /* log:104-105 */
      bool print_now = log_on && !clean.empty() && max_lines > line_cnt && min_time >= warm_up; /* log:220-226 */
      if (print_now && first_state && !have_first) {            /* mjp_log_out.first_state */
        have_first = true;
        for (size_t q = 0; q < tick_end_state.size(); ++q) first_state[q] = tick_end_state[q];
      }
      if (print_now) print_lines(clean);
      log_buf.swap(later);                                      /* log:108 */
    }
    return true;
  }

  /* end-main-loop? core:722-727 */
  bool end_main_loop() const {
    return !((run_to_job >= 0 && current_job <= run_to_job) || (clock <= run_to));
  }

  /* main-loop-loop core:729-744 */
  void main_loop_loop() {
    std::vector<Act> all, acts;
    double last_clock = -1.0;
    for (;;) {
      if (end_main_loop()) { status = MJP_INST_NORMAL_END; return; }
      runables(all, acts);
      if (acts.empty()) { status = MJP_INST_NO_RUNABLES; return; }
      double next_clk = acts[0].time;
      if (next_clk == kInf) { status = MJP_INST_NO_RUNABLES; return; }      /* D2 */
      if (clock > next_clk) { status = MJP_INST_CLOCK_BACK; return; }       /* advance-clock core:128-135 */
      for (const Act &a : acts) if (a.fn == FN_AJ && a.job.type < 0) {
        status = tape_out ? MJP_INST_TAPE_EXHAUSTED : MJP_INST_JOBTYPE_NIL; return;
      }
      clock = next_clk;
      iterations++;
      if (clock != last_clock) { ticks++; last_clock = clock; }
      for (const Act &a : acts) run_action(a);                              /* core:742 */
      for (const Machine &mc : mach) if (mc.ud.exhausted) { status = MJP_INST_TAPE_EXHAUSTED; return; }
      if (!push_log()) { status = MJP_INST_LOGBUF_EMPTY; return; }          /* core:743 */
    }
  }
};

bool valid_desc(const mjp_line_desc *L, uint64_t n_inst) {
  if (!L || L->M < 1 || L->M > MJP_MAX_MACHINES || L->K < 1 || L->K > MJP_MAX_JOBTYPES) return false;
  if (!L->lambda || !L->mu || !L->W || !L->discipline || !L->portion || !L->w) return false;
  if (L->M > 1 && !L->N) return false;
  for (int i = 0; i < L->M; ++i) {
    if (!(L->lambda[i] >= 0) || !(L->mu[i] > 0) || !(L->W[i] > 0) || L->discipline[i] > 1) return false;
    if (i < L->M - 1 && L->N[i] < 1) return false;
  }
  for (int k = 0; k < L->K; ++k) if (!(L->portion[k] > 0)) return false;
  for (int q = 0; q < L->K * L->M; ++q) if (!(L->w[q] > 0)) return false;
  if (!(L->warm_up >= 0) || !(L->run_to > 0)) return false;
  if (L->N_inst) for (int b = 0; b < L->M - 1; ++b) for (uint64_t g = 0; g < n_inst; ++g) {
    int v = L->N_inst[(size_t)b * n_inst + g];
    if (v < 1 || v > L->N[b]) return false;
  }
  return true;
}

std::vector<int> default_group_rank(int M) {
  std::vector<std::pair<uint64_t, int>> keyed;
  for (int i = 0; i < M; ++i) {
    std::string nm = "m" + std::to_string(i + 1);
    keyed.push_back({clj_trie_key(clj_keyword_hash(nm.data(), nm.size())), i});
  }
  std::sort(keyed.begin(), keyed.end());
  std::vector<int> rank(M);
  for (int pos = 0; pos < M; ++pos) rank[keyed[pos].second] = pos;
  return rank;
}

/* preprocess-model core:578-615 restricted to what the loop reads */
const mjp_draw_tapes *g_tapes = nullptr;      /* set by mjpo_set_tapes for MJP_RNG_TAPE runs */

void build_sim(Sim &s, const mjp_line_desc *L, uint64_t n_inst, uint64_t g, const mjp_rng_spec *rng) {
  s.M = L->M; s.K = L->K;
  s.draws.seed = rng->seed; s.draws.inst = rng->first_instance_id + g;
  s.mach.resize(s.M); s.N.resize(s.M > 1 ? s.M - 1 : 0); s.holding.resize(s.N.size());
  s.blocked.assign(s.M, 0); s.starved.assign(s.M, 0); s.started.assign(s.M, 0);
  s.portion.assign(L->portion, L->portion + s.K);
  s.w.resize((size_t)s.K * s.M);
  for (int q = 0; q < s.K * s.M; ++q) s.w[q] = L->w_inst ? L->w_inst[(size_t)q * n_inst + g] : L->w[q];
  for (int b = 0; b < s.M - 1; ++b) s.N[b] = L->N_inst ? L->N_inst[(size_t)b * n_inst + g] : L->N[b];
  if (rng->mode == MJP_RNG_TAPE && g_tapes) {
    s.tape_jt = g_tapes->jobtype + (size_t)g * g_tapes->len_jobtype; s.len_jt = g_tapes->len_jobtype;
  }
  /* hash-map rank of the machine keywords: given, or that of the names :m1 .. :mM the reference's models use */
  if (L->group_rank) s.group_rank.assign(L->group_rank, L->group_rank + s.M);
  else s.group_rank = default_group_rank(s.M);
  s.clock = 0.0; s.warm_up = L->warm_up; s.run_to = L->run_to; s.run_to_job = L->run_to_job;
  s.jm2dm = L->jm2dm; s.log_on = L->log; s.up_down = L->up_down; s.max_lines = L->max_lines;
  for (int i = 0; i < s.M; ++i) {
    Machine &m = s.mach[i];
    m.W = L->W_inst ? L->W_inst[(size_t)i * n_inst + g] : L->W[i];
    m.bbs = L->discipline[i];
    m.ud.lambda = L->lambda_inst ? L->lambda_inst[(size_t)i * n_inst + g] : L->lambda[i];
    m.ud.mu = L->mu_inst ? L->mu_inst[(size_t)i * n_inst + g] : L->mu[i];
    m.ud.mach = i; m.ud.mode = rng->mode; m.ud.d = &s.draws;
    if (rng->mode == MJP_RNG_TAPE && g_tapes) {
      m.ud.tape_u = g_tapes->uniform + ((size_t)g * s.M + i) * g_tapes->len_uniform; m.ud.len_u = g_tapes->len_uniform;
      m.ud.tape_e = g_tapes->exp + ((size_t)g * s.M + i) * g_tapes->len_exp; m.ud.len_e = g_tapes->len_exp;
    }
    m.ud.init();                              /* preprocess-equip core:557-562 */
    m.future = m.ud.seq[0];
  }
  /* steady-form log:144-152 */
  Steady &r = s.steady;
  r.blocked.assign(s.M, 0.0); r.starved.assign(s.M, 0.0);
  r.bs.assign(s.M, kNaN); r.ss.assign(s.M, kNaN);
  r.occ.resize(s.N.size()); r.lastclk.assign(s.N.size(), s.warm_up);
  for (size_t b = 0; b < s.N.size(); ++b) r.occ[b].assign(s.N[b] + 1, 0.0);
}

void store_stats(const Sim &s, const mjp_line_desc *L, uint64_t n_inst, uint64_t g, mjp_stats_out *o) {
  if (!o) return;
  if (o->njobs) o->njobs[g] = s.steady.njobs;
  if (o->residence_sum) o->residence_sum[g] = s.steady.residence_sum;
  for (int i = 0; i < s.M; ++i) {
    if (o->blocked_time) o->blocked_time[(size_t)i * n_inst + g] = s.steady.blocked[i];
    if (o->starved_time) o->starved_time[(size_t)i * n_inst + g] = s.steady.starved[i];
  }
  if (o->occ_time) {
    size_t row = 0;
    for (int b = 0; b < s.M - 1; ++b) {
      for (int n = 0; n <= L->N[b]; ++n)
        o->occ_time[(row + n) * n_inst + g] = (n <= s.N[b]) ? s.steady.occ[b][n] : 0.0;
      row += (size_t)L->N[b] + 1;
    }
  }
  if (o->final_clock) o->final_clock[g] = s.clock;
  if (o->current_job) o->current_job[g] = s.current_job;
  if (o->n_events) o->n_events[g] = s.n_events;
  if (o->event_hash) o->event_hash[g] = s.hash;
  if (o->status) o->status[g] = (uint8_t)s.status;
}

}  // namespace

/* ======================================================================== */
extern "C" {

int mjp_aggregate_len_oracle(int32_t M) { return 1 + 2 * (3 * M + 1); }

void mjpo_set_tapes(const mjp_draw_tapes *t) { g_tapes = t; }

/* Run in MJP_RNG_REPLAY and RECORD every draw per consumer into caller buffers shaped like mjp_draw_tapes
 * (the buffers behind the const pointers are written).  used[g*(2M+1) + ...] = draws consumed:
 * [0] job type, [1+i] uniforms of machine i, [1+M+i] exponentials of machine i.                        */
int mjpo_record_tapes(const mjp_line_desc *L, uint64_t n_inst, const mjp_rng_spec *rng, const mjp_draw_tapes *out,
                      uint32_t *used) {
  if (!rng || rng->mode != MJP_RNG_REPLAY || !valid_desc(L, n_inst) || L->M < 2 || !out) return MJP_ERR_INVALID;
  const int M = L->M;
  for (uint64_t g = 0; g < n_inst; ++g) {
    Sim s;
    build_sim(s, L, n_inst, g, rng);
    /* build_sim already drew the first event of every machine: replay those two draws into the tapes */
    s.rec_jt = const_cast<double *>(out->jobtype) + (size_t)g * out->len_jobtype; s.rec_len_jt = out->len_jobtype;
    for (int i = 0; i < M; ++i) {
      UpDown &u = s.mach[i].ud;
      u.rec_u = const_cast<double *>(out->uniform) + ((size_t)g * M + i) * out->len_uniform; u.rec_len_u = out->len_uniform;
      u.rec_e = const_cast<double *>(out->exp) + ((size_t)g * M + i) * out->len_exp; u.rec_len_e = out->len_exp;
      u.rec_nu = u.rec_ne = 0;
      u.note_u(s.draws.uniform(u.cu(), 0));
      u.note_e(u.seq[0].t);
    }
    s.main_loop_loop();
    if (used) {
      uint32_t *row = used + (size_t)g * (2 * M + 1);
      row[0] = s.jt_idx;
      for (int i = 0; i < M; ++i) { row[1 + i] = s.mach[i].ud.rec_nu; row[1 + M + i] = s.mach[i].ud.rec_ne; }
    }
  }
  return MJP_OK;
}

int mjpo_run(const mjp_line_desc *L, uint64_t n_inst, const mjp_rng_spec *rng,
             mjp_stats_out *stats, mjp_log_out *log, mjpo_extra *extra, int n_threads) {
  if (!rng || !valid_desc(L, n_inst) || L->M < 2) return MJP_ERR_INVALID;
  if (rng->mode == MJP_RNG_TAPE && !g_tapes) return MJP_ERR_INVALID;
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > n_inst) n_threads = (int)(n_inst ? n_inst : 1);
  std::vector<std::vector<mjp_log_record>> sinks(n_inst && log && log->records ? n_inst : 0);
  const bool want_dets = extra && extra->dets && !sinks.empty();
  std::vector<std::vector<int64_t>> dets_sinks(want_dets ? n_inst : 0);
  std::atomic<uint64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      uint64_t g = next.fetch_add(1);
      if (g >= n_inst) return;
      Sim s;
      build_sim(s, L, n_inst, g, rng);
      s.inst_index = (uint32_t)g;
      if (!sinks.empty()) s.sink = &sinks[g];
      if (want_dets) { s.job_detail = true; s.dets_sink = &dets_sinks[g]; }
      if (log && log->first_state) {
        s.first_state = log->first_state + (size_t)g * 2 * (size_t)L->M;
        for (int q = 0; q < 2 * L->M; ++q) s.first_state[q] = 0xFFFFFFFFu;
      }
      s.main_loop_loop();
      store_stats(s, L, n_inst, g, stats);
      if (extra) {
        if (extra->act_counts) for (int a = 0; a < 10; ++a) extra->act_counts[(size_t)a * n_inst + g] = s.act_counts[a];
        if (extra->iterations) extra->iterations[g] = s.iterations;
        if (extra->ticks) extra->ticks[g] = s.ticks;
        if (extra->max_tick_msgs) extra->max_tick_msgs[g] = s.max_tick_msgs;
        if (extra->wide_ticks) extra->wide_ticks[g] = s.wide_ticks;
        if (extra->wide_ticks_reordered) extra->wide_ticks_reordered[g] = s.wide_ticks_reordered;
        if (extra->max_lookahead) extra->max_lookahead[g] = s.max_lookahead;
      }
    }
  };
  if (n_threads == 1) worker();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
  int rc = MJP_OK;
  if (log) {
    log->count = 0; log->dropped = 0;
    for (auto &v : sinks) for (const mjp_log_record &r : v) {
      if (log->count < log->capacity) log->records[log->count++] = r; else log->dropped++;
    }
    if (log->dropped) rc = MJP_LOG_TRUNCATED;
  }
  if (want_dets) {                                              /* same order as the records; truncated like them */
    uint64_t len = 0;
    for (auto &v : dets_sinks) for (int64_t x : v) { if (len < extra->dets_capacity) extra->dets[len] = x; ++len; }
    if (extra->dets_len) *extra->dets_len = len;
  }
  if (stats && stats->aggregate && stats->njobs && stats->residence_sum && stats->blocked_time &&
      stats->starved_time && stats->occ_time && stats->status) {
    int len = mjp_aggregate_len_oracle(L->M);
    for (int q = 0; q < len; ++q) stats->aggregate[q] = 0.0;
    for (uint64_t g = 0; g < n_inst; ++g) {
      if (stats->status[g] != MJP_INST_NORMAL_END) continue;
      mjpo_basics b;
      mjpo_calc_basics(L, stats, n_inst, g, &b);
      double *a = stats->aggregate; int q = 1;
      a[0] += 1.0;
      auto add = [&](double x) { a[q] += x; a[q + 1] += x * x; q += 2; };
      add((double)b.njobs / b.runtime);
      add(b.njobs ? b.residence : 0.0);
      for (int i = 0; i < L->M; ++i) add(b.blocked[i]);
      for (int i = 0; i < L->M; ++i) add(b.starved[i]);
      for (int i = 0; i < L->M - 1; ++i) add(b.occupancy[i]);
    }
  }
  return rc;
}

double mjpo_end_time(const uint8_t *kinds, const double *times, int32_t n_events, double clock, double dur) {
  Sim s; s.M = 1; s.mach.resize(1); s.clock = clock;
  Machine &m = s.mach[0];
  m.ud.finite = true;
  for (int j = 0; j < n_events; ++j) m.ud.seq.push_back({(uint32_t)j, kinds[j], times[j]});
  m.future = m.ud.seq[0];
  return s.end_time(dur, 0);
}

int mjpo_updown_events(double lambda, double mu, const mjp_rng_spec *rng, uint64_t inst, int32_t machine,
                       int32_t n, uint8_t *kinds, double *times, uint32_t *draws_u, uint32_t *draws_e) {
  Draws d{rng->seed, rng->first_instance_id + inst};
  UpDown ud; ud.lambda = lambda; ud.mu = mu; ud.mach = machine; ud.mode = rng->mode; ud.d = &d;
  ud.init();
  int got = 0;
  for (; got < n; ++got) {
    const Ev *e = ud.at((size_t)got);
    if (!e) break;
    kinds[got] = e->kind; times[got] = e->t;
  }
  if (draws_u) *draws_u = ud.u_idx;
  if (draws_e) *draws_e = ud.e_idx;
  return got;
}

int mjpo_log_script(int32_t M, const int32_t *N, double warm_up, uint8_t up_down, const uint8_t *ops,
                    const mjp_log_record *msgs, int32_t n_ops, mjp_stats_out *stats, mjp_log_out *log) {
  Sim s; s.M = M; s.warm_up = warm_up; s.up_down = up_down;
  s.group_rank = default_group_rank(M);
  s.N.assign(N, N + (M > 1 ? M - 1 : 0));
  s.log_on = (log != nullptr); s.max_lines = UINT64_MAX;
  std::vector<mjp_log_record> sink; if (log) s.sink = &sink;
  Steady &r = s.steady;
  r.blocked.assign(M, 0.0); r.starved.assign(M, 0.0); r.bs.assign(M, kNaN); r.ss.assign(M, kNaN);
  r.occ.resize(s.N.size()); r.lastclk.assign(s.N.size(), warm_up);
  for (size_t b = 0; b < s.N.size(); ++b) r.occ[b].assign(s.N[b] + 1, 0.0);
  for (int j = 0; j < n_ops; ++j) {
    if (ops[j] == 0) s.log({msgs[j].act, msgs[j].machine, msgs[j].job, msgs[j].n, msgs[j].aux, msgs[j].clk});
    else s.push_log();
  }
  std::vector<int32_t> Ncap(s.N.begin(), s.N.end());
  mjp_line_desc L; std::memset(&L, 0, sizeof L); L.M = M; L.N = Ncap.data();
  s.status = MJP_INST_NORMAL_END;
  store_stats(s, &L, 1, 0, stats);
  if (log) {
    log->count = 0; log->dropped = 0;
    for (const mjp_log_record &q : sink) { if (log->count < log->capacity) log->records[log->count++] = q; else log->dropped++; }
  }
  return MJP_OK;
}

/* calc-basics core:617-634 */
int mjpo_calc_basics(const mjp_line_desc *L, const mjp_stats_out *o, uint64_t n_inst, uint64_t g, mjpo_basics *out) {
  double run_time = L->run_to - L->warm_up;
  int64_t njobs = o->njobs[g];
  out->runtime = run_time;
  out->TP = (float)((double)njobs / run_time);
  out->residence = njobs == 0 ? kNaN : o->residence_sum[g] / (double)njobs;
  out->njobs = njobs;
  for (int i = 0; i < L->M; ++i) {
    out->blocked[i] = o->blocked_time[(size_t)i * n_inst + g] / run_time;
    out->starved[i] = o->starved_time[(size_t)i * n_inst + g] / run_time;
  }
  size_t row = 0;
  for (int b = 0; b < L->M - 1; ++b) {
    double acc = 0.0;
    for (int n = 0; n <= L->N[b]; ++n) acc = acc + (double)n * o->occ_time[(row + n) * n_inst + g];
    out->occupancy[b] = acc / run_time;
    row += (size_t)L->N[b] + 1;
  }
  return MJP_OK;
}

/* calc-bneck core:636-661; bl/st are 1-indexed in the reference */
int mjpo_calc_bneck(int32_t M, const double *bl, const double *st) {
  int n_cand = 0, cand = -1;
  for (int i = 0; i < M - 1; ++i) if (bl[i] < st[i + 1]) { if (!n_cand) cand = i; n_cand++; }   /* core:643-644 */
  if (n_cand == 1) return cand;                                                                 /* core:646-647 */
  int best = 0; double best_v = -1.0;
  for (int i = 0; i < M; ++i) {                                                                 /* core:651-657 */
    double sev;
    if (i == 0) sev = std::fabs(bl[0] - st[1]);
    else if (i == M - 1) sev = std::fabs(bl[M - 2] - st[M - 1]);
    else sev = std::fabs(bl[i - 1] - st[i]) + std::fabs(bl[i] - st[i + 1]);
    if (sev > best_v) { best_v = sev; best = i; }          /* first with the max value, core:659-661 */
  }
  return best;
}

uint32_t mjpo_clj_keyword_hash(const char *name) { return clj_keyword_hash(name, std::strlen(name)); }
uint64_t mjpo_clj_trie_key(uint32_t hash) { return clj_trie_key(hash); }
void mjpo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
double mjpo_log(double x) { return det_log(x); }
double mjpo_exp(double x) { return det_exp(x); }
double mjpo_uniform(uint64_t seed, uint64_t inst, uint32_t c, uint32_t idx) { Draws d{seed, inst}; return d.uniform(c, idx); }
double mjpo_open_uniform(uint64_t seed, uint64_t inst, uint32_t c, uint32_t idx) { Draws d{seed, inst}; return d.open_uniform(c, idx); }

}  // extern "C"
