/*
 * mjp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal single-threaded restatement of the MJPdes run loop
 * (reference: src/gov/nist/MJPdes/core.clj, util/log.clj, util/utils.clj),
 * used ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and the timed CPU baseline.  Nothing
 * under mjpdes_b200/ may link, import or call it.
 *
 * Parity status: PINNED on the reference's deterministic known-answer tests
 * (G1-G9 of SURVEY.md 8c: core_test.clj:26-93,102-175,199-247,251-287 and
 * util/log_test.clj:17-77); UNPINNED at the level of individual random draws,
 * because the reference draws from unseeded JVM-global generators and no JVM
 * exists in this image (see DESIGN.md "Oracle").
 *
 * It shares only the interface structs of include/mjpdes_b200.h with the
 * product; all arithmetic (Philox, log, exp) is implemented here a second
 * time, independently of mjpdes_b200/csrc.
 */
#ifndef MJP_ORACLE_H
#define MJP_ORACLE_H

#include "../include/mjpdes_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Extended per-instance output of the oracle (tests only). */
typedef struct mjpo_extra {
  uint64_t *act_counts;     /* [10][n] executed actions by MJP_ACT_*          */
  uint64_t *iterations;     /* [n] loop iterations (micro-steps)              */
  uint64_t *ticks;          /* [n] distinct clock values                      */
  uint32_t *max_tick_msgs;  /* [n] largest same-clock message batch           */
  uint32_t *max_lookahead;  /* [n] deepest end-time look-ahead (events past :future) */
  uint64_t *wide_ticks;     /* [n] post-warm-up ticks in which more than 8 machines logged (hash-map regrouping) */
  uint64_t *wide_ticks_reordered; /* [n] ... of which the cleaned batch differs from first-appearance grouping */
  /* [:report :job-detail?] (util/log.clj:11-28): non-NULL turns the :dets snapshots on; those of the PRINTED messages
   * are written here in record order, each flattened as the M values of :run (-1 = nil) followed, per buffer, by
   * (count :holding) and the job ids.  dets_len receives the words produced (may exceed dets_capacity).  */
  int64_t  *dets;
  uint64_t  dets_capacity;
  uint64_t *dets_len;
} mjpo_extra;

/* Position of `name` (a keyword without colon or namespace, e.g. "m7") in Clojure hash-map iteration order is
 * decided by this key: PersistentHashMap walks its trie by ascending 5-bit chunks of the key's hasheq.        */
uint32_t mjpo_clj_keyword_hash(const char *name);
uint64_t mjpo_clj_trie_key(uint32_t hash);

/* main-loop-loop (core.clj:729-744) for instances [0, n_inst) on n_threads
 * host threads.  Same descriptor / rng / output structs as mjp_run().       */
int mjpo_run(const mjp_line_desc *line, uint64_t n_inst, const mjp_rng_spec *rng,
             mjp_stats_out *stats, mjp_log_out *log, mjpo_extra *extra, int n_threads);

/* tapes for subsequent MJP_RNG_TAPE runs (pointer kept, not copied; NULL clears) */
void mjpo_set_tapes(const mjp_draw_tapes *tapes);

/* record the draws of a MJP_RNG_REPLAY run per consumer (stand-in for hooking rand / sample-exp in a JVM) */
int mjpo_record_tapes(const mjp_line_desc *line, uint64_t n_inst, const mjp_rng_spec *rng,
                      const mjp_draw_tapes *out, uint32_t *used);

/* end-time (core.clj:138-151) over a literal :up&down seq (core_test.clj:26-45). */
double mjpo_end_time(const uint8_t *kinds, const double *times, int32_t n_events,
                     double clock, double dur);

/* The up/down generator alone (core.clj:156-192): first n events of machine
 * `machine` of instance `inst`.  kinds: 0 = :up, 1 = :down.                   */
int mjpo_updown_events(double lambda, double mu, const mjp_rng_spec *rng,
                       uint64_t inst, int32_t machine, int32_t n,
                       uint8_t *kinds, double *times, uint32_t *draws_u, uint32_t *draws_e);

/* log / push-log / add-compute-log! driven by hand (core_test.clj:102-167):
 * ops[j] = 0: (log/log msg j)  1: (log/push-log).  msgs use mjp_log_record
 * (act, machine, clk, n, aux, job).  Fills the same stats arrays (n_inst=1). */
int mjpo_log_script(int32_t M, const int32_t *N, double warm_up, uint8_t up_down,
                    const uint8_t *ops, const mjp_log_record *msgs, int32_t n_ops,
                    mjp_stats_out *stats, mjp_log_out *log);

/* calc-basics (core.clj:617-634) on one instance's accumulators. */
typedef struct mjpo_basics {
  double runtime; float TP; double residence; /* NaN = :na */ int64_t njobs;
  double blocked[MJP_MAX_MACHINES], starved[MJP_MAX_MACHINES], occupancy[MJP_MAX_MACHINES];
} mjpo_basics;
int mjpo_calc_basics(const mjp_line_desc *line, const mjp_stats_out *stats,
                     uint64_t n_inst, uint64_t inst, mjpo_basics *out);

/* calc-bneck (core.clj:636-661), topology-indexed (exact for M <= 8). */
int mjpo_calc_bneck(int32_t M, const double *blocked, const double *starved);

/* The arithmetic of the substream contract, exposed for cross-checks. */
void   mjpo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double mjpo_log(double x);
double mjpo_exp(double x);
double mjpo_uniform(uint64_t seed, uint64_t inst, uint32_t consumer, uint32_t idx);      /* [0,1) */
double mjpo_open_uniform(uint64_t seed, uint64_t inst, uint32_t consumer, uint32_t idx); /* (0,1) */

#ifdef __cplusplus
}
#endif
#endif
