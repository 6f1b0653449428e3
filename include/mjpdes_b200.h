/*
 * mjpdes_b200.h -- C ABI of the B200-native batched MJPdes run loop.
 *
 * The reference (usnistgov/MJPdes, pure Clojure) has no FFI; the seam this
 * library replaces is ONE function and its side effects:
 *
 *     (core/main-loop-loop model job-end time-end)      src/gov/nist/MJPdes/core.clj:729-744
 *       in : the preprocessed Model map (core.clj:578-615)
 *       out: [:params :status]                          core.clj:735,739
 *            the log-steady ref (accumulators)          util/log.clj:139-152, 71-86
 *            SCADA messages handed to print-lines       util/log.clj:89-137
 *
 * called once per replication by main-loop-once (core.clj:783) and once per
 * future by main-loop-multi (core.clj:764).  The batch dimension n_inst of
 * mjp_run() is main-loop-multi's `(range :number-of-simulations)`.
 *
 * Rules of the ABI: plain C, little-endian, caller owns every host buffer,
 * the library owns every device buffer, every call returns an int status and
 * never aborts or throws; a handle is not thread-safe, distinct handles are.
 * Paths below are relative to the reference's src/gov/nist/MJPdes/.
 */
#ifndef MJPDES_B200_H
#define MJPDES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MJP_ABI_VERSION 3

/* ---- call status ------------------------------------------------------- */
enum {
  MJP_OK              = 0,
  MJP_ERR_INVALID     = 1,  /* descriptor fails the checks of core.clj:75-121 (spec) */
  MJP_ERR_CUDA        = 2,  /* CUDA runtime error, text in mjp_last_error()          */
  MJP_ERR_NOMEM       = 3,
  MJP_ERR_UNSUPPORTED = 4,  /* valid model, outside what the kernels cover           */
  MJP_LOG_TRUNCATED   = 5   /* run completed; SCADA record buffer was too small      */
};

/* ---- per-instance status byte: mirrors [:params :status] ---------------- */
enum {
  MJP_INST_NORMAL_END   = 0,  /* :normal-end   core.clj:735                          */
  MJP_INST_NO_RUNABLES  = 1,  /* :no-runables  core.clj:739                          */
  MJP_INST_CLOCK_BACK   = 2,  /* ex-info "clock running backwards" core.clj:133      */
  MJP_INST_JOBTYPE_NIL  = 3,  /* new-job fell off the shifted thresholds (core.clj:332-339)
                                 -> the reference dies in job-requires (utils.clj:95)  */
  MJP_INST_LOGBUF_EMPTY = 4,  /* push-log returned nil (util/log.clj:95): reference dies */
  MJP_INST_TICK_OVERFLOW= 5,  /* device-only: more same-clock messages than the staging
                                 area holds; instance stopped, statistics invalid     */
  MJP_INST_ITER_LIMIT   = 6,  /* device-only: watchdog on loop iterations tripped        */
  MJP_INST_PAUSED       = 7,  /* mjp_launch_slice(): parked at the slice boundary, resumable  */
  MJP_INST_TAPE_EXHAUSTED = 8, /* MJP_RNG_TAPE: a draw tape ran out                              */
  MJP_INST_NOT_RUN      = 255
};

/* ---- actions ("events"): one per run-action call, core.clj:533-536 ------ */
enum {
  MJP_ACT_AJ = 0,   /* add-job          core.clj:240  -> "<m>-start-job"           */
  MJP_ACT_UP = 1,   /* bring-machine-up core.clj:215  -> "<m>-up"                  */
  MJP_ACT_DOWN = 2, /* bring-machine-down core.clj:230 -> "<m>-down"               */
  MJP_ACT_BJ = 3,   /* advance2buffer   core.clj:285  -> "<m>-complete-job" (:bj)  */
  MJP_ACT_EJ = 4,   /* advance2buffer on the last machine (:ej)                    */
  MJP_ACT_SM = 5,   /* advance2machine  core.clj:263  -> "<m>-start-job"           */
  MJP_ACT_ST = 6,   /* new-starved      core.clj:316                               */
  MJP_ACT_US = 7,   /* new-unstarved    core.clj:322                               */
  MJP_ACT_BL = 8,   /* new-blocked      core.clj:304                               */
  MJP_ACT_UB = 9    /* new-unblocked    core.clj:310                               */
};

#define MJP_BAS 0   /* blocked-after-service  (default, core.clj:556) */
#define MJP_BBS 1   /* blocked-before-service */

#define MJP_MAX_MACHINES 64
#define MJP_MAX_JOBTYPES 255

/* ---- line descriptor: the flat form of the preprocessed Model ----------- */
/* Machines 0..M-1 in :topology order, buffer b between machine b and b+1;
 * :entry-point must be machine 0.  `portion` and the rows of `w` are in
 * :jobmix ITERATION order (insertion order for <= 8 job types), because
 * new-job walks the map (core.clj:332).                                      */
typedef struct mjp_line_desc {
  int32_t M;                 /* machines, 2..MJP_MAX_MACHINES                  */
  int32_t K;                 /* job types, 1..MJP_MAX_JOBTYPES                 */
  const double  *lambda;     /* [M]  failure rate, >= 0  (0: never fails, core.clj:178) */
  const double  *mu;         /* [M]  repair rate,  > 0                         */
  const double  *W;          /* [M]  work capacity, > 0                        */
  const uint8_t *discipline; /* [M]  MJP_BAS / MJP_BBS                         */
  const int32_t *N;          /* [M-1] buffer capacity, >= 1                    */
  const double  *portion;    /* [K]                                            */
  const double  *w;          /* [K*M] w[k*M+i] = work of job type k on machine i */
  double  warm_up;           /* [:params :warm-up-time]                        */
  double  run_to;            /* [:params :run-to-time]                         */
  int64_t run_to_job;        /* [:params :run-to-job], -1 = unset (core.clj:724) */
  uint8_t jm2dm;             /* :jm2dm (jobs move to down machines), core.clj:582 */
  uint8_t log;               /* [:report :log?]                                */
  uint8_t up_down;           /* [:report :up&down?]  (util/log.clj:33)         */
  uint8_t reserved0;
  uint32_t reserved1;
  uint64_t max_lines;        /* [:report :max-lines]; UINT64_MAX = ##Inf       */
  /* Optional per-instance override planes for sweeps, instance-fastest
   * (plane[j*n_inst + inst]); NULL = use the shared vector above.  When
   * N_inst is given, N[b] must be >= every N_inst value of buffer b (it sizes
   * the occupancy-level layout of the output).                               */
  const double  *lambda_inst;  /* [M][n_inst]   */
  const double  *mu_inst;      /* [M][n_inst]   */
  const double  *W_inst;       /* [M][n_inst]   */
  const int32_t *N_inst;       /* [M-1][n_inst] */
  const double  *w_inst;       /* [K*M][n_inst] */
  /* Optional: group_rank[i] = position of machine i's NAME among the machine names in Clojure hash-map
   * iteration order.  clean-log-buf regroups a tick's messages with (group-by :m ...) (util/log.clj:44), whose
   * result iterates in insertion order up to 8 machines and in keyword-hash order from the 9th on, so the
   * order is only needed for ticks in which MORE than 8 machines log.  Must be a permutation of 0..M-1.
   * NULL = the names :m1 .. :mM of the reference's models (resources/example.clj etc.), ranked by the library.            */
  const int32_t *group_rank;   /* [M] */
} mjp_line_desc;

/* ---- random numbers ------------------------------------------------------ */
/* The reference draws from two unseeded JVM-global generators (clojure.core/rand
 * at core.clj:159,185,331 and incanter sample-exp at core.clj:160,187,191), so
 * "the reference stream" is only definable per CONSUMER.  Consumers of one
 * instance: c=0 job type (one uniform per new-job); c=1+2i uniforms of machine i
 * (init up?/down? draw, then one `some-num` per sojourn); c=2+2i exponentials of
 * machine i.  Draw #idx of consumer c of instance g is words (0,1) of
 *     Philox4x32-10(key = seed, counter = {idx, c, g_lo, g_hi}),
 * except for the job-type stream (c=0), which uses both halves of a block:
 * draw #d is words (0,1) of block {d >> 1, 0, g} for even d, words (2,3) for odd d.
 *   MJP_RNG_COUNTER : sojourn n of a machine is Erlang-2, T = -ln(u1*u2)/rate,
 *                     from the four words of ONE Philox block {n, 1+2i, g};
 *                     distribution-identical to the loop at core.clj:185-191.
 *   MJP_RNG_REPLAY  : the literal loop of core.clj:185-191, one uniform from
 *                     stream 1+2i, exponentials -ln(u)/rate from stream 2+2i until
 *                     u < 1 - e^(-rate*T); draw counts match the reference's.
 *   MJP_RNG_TAPE    : the literal loop again, but every draw is READ from caller-supplied tapes
 *                     (mjp_set_tapes): the way to replay the reference's OWN stream -- record what
 *                     clojure.core/rand and incanter's sample-exp returned in a JVM run (with-redefs on
 *                     the call sites above, one tape per consumer) and feed it here.                 */
enum { MJP_RNG_COUNTER = 0, MJP_RNG_REPLAY = 1, MJP_RNG_TAPE = 2 };

/* Draw tapes for MJP_RNG_TAPE, instance-major.  `exp` holds the VARIATES sample-exp returned (already
 * divided by the rate).  An instance that runs off a tape ends with MJP_INST_TAPE_EXHAUSTED.  The device
 * draws lazily: it never needs more of a tape than the reference's own run consumed (end-time realises
 * up/down events ahead of time, core.clj:138-151), so a tape recorded from a complete reference run is
 * always sufficient; on a tape cut short it may get further than a look-ahead implementation would.   */
typedef struct mjp_draw_tapes {
  uint32_t len_jobtype;     /* draws per instance                                                  */
  uint32_t len_uniform;     /* draws per machine per instance                                      */
  uint32_t len_exp;         /* draws per machine per instance                                      */
  uint32_t reserved;
  const double *jobtype;    /* [n_inst][len_jobtype]     rand in new-job, core.clj:331              */
  const double *uniform;    /* [n_inst][M][len_uniform]  rand at core.clj:159 (first), then :185    */
  const double *exp;        /* [n_inst][M][len_exp]      sample-exp at core.clj:160 (first), :187, :191 */
} mjp_draw_tapes;

typedef struct mjp_rng_spec {
  int32_t  mode;
  int32_t  reserved;
  uint64_t seed;
  uint64_t first_instance_id;  /* instance j of this call uses id first_instance_id + j */
} mjp_rng_spec;

/* ---- per-instance results: the content of the log-steady ref ------------------- */
/* Caller-allocated, instance-fastest SoA.  L = sum_b (N[b]+1) occupancy levels;
 * level n of buffer b is row  (sum_{c<b} (N[c]+1)) + n.  Any pointer may be
 * NULL (plane not wanted).                                                   */
typedef struct mjp_stats_out {
  int64_t  *njobs;          /* [n]     :njobs          util/log.clj:217        */
  double   *residence_sum;  /* [n]     :residence-sum  util/log.clj:216        */
  double   *blocked_time;   /* [M][n]  (:blocked (mK)) util/log.clj:190        */
  double   *starved_time;   /* [M][n]  (:starved (mK)) util/log.clj:208        */
  double   *occ_time;       /* [L][n]  (bK n)          util/log.clj:163,170    */
  double   *final_clock;    /* [n]     :clock when the loop ended              */
  int64_t  *current_job;    /* [n]     [:params :current-job]                  */
  uint64_t *n_events;       /* [n]     run-action calls (the benchmark's "event") */
  uint64_t *event_hash;     /* [n]     hash of every executed action, see mjp_event_hash.. */
  uint8_t  *status;         /* [n]     MJP_INST_*                              */
  /* Optional aggregate over the instances of THIS call (for the NCCL
   * reduction across ranks): see mjp_aggregate_len().                        */
  double   *aggregate;
} mjp_stats_out;

/* aggregate vector: [count, then for each metric x in
 *   TP, residence, blocked[0..M), starved[0..M), occupancy[0..M-1): sum(x), sum(x^2)]
 * with the per-instance metrics formed exactly as calc-basics does (core.clj:617-634),
 * over instances whose status is MJP_INST_NORMAL_END.  Length = 1 + 2*(3*M+1).  */
int mjp_aggregate_len(int32_t M);

/* ---- SCADA records -------------------------------------------------------- */
/* What push-log hands to print-lines (util/log.clj:107,111): the cleaned
 * messages of one clock tick, in clean-log-buf order, `line` numbered per
 * instance from 0 (util/log.clj:116,126).  Text rendering is host-side.       */
typedef struct mjp_log_record {
  double   clk;       /* :clk                                                  */
  double   aux;       /* :ends for aj, :ent for ej, NaN otherwise              */
  uint32_t job;       /* :j (aj, bj, ej, sm), else 0                           */
  uint32_t instance;  /* index within this call                                */
  uint32_t line;      /* :line                                                 */
  uint8_t  n;         /* :n for bj/sm (count BEFORE the move); job-type index for aj */
  uint8_t  ord;       /* position of the message in :log-buf, i.e. in the tick's EXECUTION order
                         (mod 256): clean-log-buf regroups a tick by machine, and the :dets
                         snapshots (util/log.clj:11-28) depend on the order the moves were made in */
  uint8_t  act;       /* MJP_ACT_*                                             */
  uint8_t  machine;   /* :m (topology index); :bf is machine-1 for sm, machine for bj */
} mjp_log_record;

typedef struct mjp_log_out {
  mjp_log_record *records;   /* caller buffer                                   */
  uint64_t capacity;         /* in records                                      */
  uint64_t count;            /* OUT: records written (<= capacity)              */
  uint64_t dropped;          /* OUT: records that did not fit                   */
  /* Optional (may be NULL): [n_inst][2 M], where the jobs were at the END of the clock tick whose records are
   * the instance's first -- what (log/detail model) (util/log.clj:11-28) needs to start from, since nothing before
   * the warm-up is in the record stream (undo that tick's moves in reverse `ord` order to get the state before it).
   * Word i: number of the last job started on machine i (0 = none yet); word M+i: bit 0 = machine i holds a job,
   * bits 8-15 = (count :holding) of buffer i.  Jobs pass through a serial line in order, so machine i holds job
   * started[i] and buffer i holds started[i+1]+1 .. started[i+1]+count.  0xFFFFFFFF everywhere: nothing printed yet. */
  uint32_t *first_state;
} mjp_log_out;

/* ---- handle ---------------------------------------------------------------- */
typedef struct mjp_config {
  int32_t device;           /* CUDA device ordinal                              */
  int32_t flags;            /* MJP_CFG_*                                        */
  uint64_t max_instances;   /* hint: device buffers are grown on demand         */
} mjp_config;

#define MJP_CFG_EVENT_HASH 1   /* compute stats.event_hash (parity runs; slower) */

typedef struct mjp_handle mjp_handle;

int  mjp_abi_version(void);
int  mjp_device_count(void);
int  mjp_create(const mjp_config *cfg, mjp_handle **out);
void mjp_destroy(mjp_handle *h);
const char *mjp_last_error(const mjp_handle *h);   /* h may be NULL: last create error */

/* Blocking, like the reference call.  Replaces main-loop-loop (core.clj:729) for
 * n_inst independent instances.  `log` may be NULL.                            */
int mjp_run(mjp_handle *h, const mjp_line_desc *line, uint64_t n_inst,
            const mjp_rng_spec *rng, mjp_stats_out *stats, mjp_log_out *log);

/* Device-resident variant used for measurement: upload once, run many times,
 * download when wanted.  mjp_run() == upload + launch + sync + download.       */
int mjp_upload(mjp_handle *h, const mjp_line_desc *line, uint64_t n_inst);
/* draw tapes for MJP_RNG_TAPE launches of the uploaded line (copied to the device; call after mjp_upload) */
int mjp_set_tapes(mjp_handle *h, const mjp_draw_tapes *tapes);
/* device buffer for SCADA records (needed before a launch with want_log); mjp_run() does it itself */
int mjp_reserve_log(mjp_handle *h, uint64_t capacity);
int mjp_launch(mjp_handle *h, const mjp_rng_spec *rng, int want_log);  /* async on the handle's stream */
int mjp_sync(mjp_handle *h);
int mjp_download(mjp_handle *h, mjp_stats_out *stats, mjp_log_out *log);
/* Time-sliced run (long horizons, streaming): advance every instance until its clock passes `until` (or the
 * run ends), park its complete state in HBM and return; call again with resume = 1 and a later `until` to
 * continue.  Slicing is invisible in the results: statistics, event hash and SCADA records are those of one
 * uninterrupted run.  Instances parked report MJP_INST_PAUSED.  SCADA records of each slice start at the
 * beginning of the record buffer (download them between slices).                                          */
int mjp_launch_slice(mjp_handle *h, const mjp_rng_spec *rng, int want_log, double until, int resume);
/* milliseconds between the CUDA events bracketing the last mjp_launch()'s kernels */
int mjp_last_kernel_ms(mjp_handle *h, float *ms);
/* number of kernels launched by the last mjp_launch() */
int mjp_last_launch_count(mjp_handle *h, int *n);

/* Device address of the aggregate vector of the last launch (mjp_aggregate_len doubles, valid after mjp_sync and
 * until the next mjp_upload): lets a multi-process host reduce it across ranks with NCCL without a host round trip. */
int mjp_device_aggregate(mjp_handle *h, double **device_ptr, int32_t *len);

/* ---- all GPUs of one box behind one call --------------------------------------- */
/* main-loop-multi (core.clj:746-766) fans the replications out over JVM futures; mjp_multi_run() fans a batch out
 * over the devices of one box: contiguous ranges of instance ids, ceil(n/G) per device (an instance's random
 * substreams depend on its id only, so results do not depend on G), one host thread and one stream per device, no
 * exchange while simulating.  The only collective is the sum of the per-device aggregate vectors:
 *   MJP_REDUCE_PEER  device 0 gathers them with peer copies over NVLink and adds them in device order;
 *   MJP_REDUCE_NCCL  ncclAllReduce(sum, f64); libnccl.so.2 is bound at run time (MJP_ERR_UNSUPPORTED if absent).
 * `stats` planes are [rows][n_inst] over the WHOLE batch (each device fills its columns); stats->aggregate
 * receives the reduced vector.  SCADA capture stays per device (use one mjp_handle per device for it).         */
enum { MJP_REDUCE_PEER = 0, MJP_REDUCE_NCCL = 1 };
typedef struct mjp_multi_config {
  int32_t n_devices;        /* 0 = every visible device                         */
  const int32_t *devices;   /* [n_devices] CUDA ordinals; NULL = 0..n_devices-1 */
  int32_t flags;            /* MJP_CFG_* for every per-device handle            */
  int32_t reduce;           /* MJP_REDUCE_*                                     */
} mjp_multi_config;
typedef struct mjp_multi mjp_multi;
int  mjp_multi_create(const mjp_multi_config *cfg, mjp_multi **out);
void mjp_multi_destroy(mjp_multi *m);
const char *mjp_multi_last_error(const mjp_multi *m);   /* m may be NULL: last create error */
int  mjp_multi_device_count(const mjp_multi *m);
int  mjp_multi_run(mjp_multi *m, const mjp_line_desc *line, uint64_t n_inst, const mjp_rng_spec *rng,
                   mjp_stats_out *stats);
/* device time of the last mjp_multi_run: max over devices of the kernel time, and the reduction */
int  mjp_multi_last_ms(mjp_multi *m, float *kernel_ms_max, float *reduce_ms);

/* Probe of one device function: end-time (core.clj:138-151) over a literal
 * :up&down schedule, the seam the reference's own tests use (core_test.clj:26-93).
 * kinds[j]: 0 = :up, 1 = :down.                                               */
int mjp_probe_end_time(mjp_handle *h, const uint8_t *kinds, const double *times,
                       int32_t n_events, double clock, double dur, double *out);

#ifdef __cplusplus
}
#endif
#endif /* MJPDES_B200_H */
