// mjp_params.h -- kernel parameter block shared by host and device code.
#pragma once
#include <cstdint>

namespace mjp {

struct KParams {
  // batch
  uint64_t n_inst;        // instances in this launch
  uint64_t first_id;      // RNG instance id of instance 0
  uint64_t seed;
  // line (flat form of the preprocessed Model, core.clj:578-615)
  int32_t M, K;
  int32_t R;              // job ring size: power of two >= M + sum N (jobs in the line)
  int32_t L;              // occupancy levels, sum (N[b]+1)
  double warm_up, run_to;
  int64_t run_to_job;     // -1 unset
  uint64_t bbs_mask;      // bit i: machine i is :BBS
  uint64_t hlead_mask;    // bit i: machine i's name precedes machine i+1's in Clojure hash-map order (util/log.clj:44, > 8 groups)
  int32_t jm2dm, up_down, log_on, reserved;
  uint64_t max_lines;
  // shared constants (device pointers)
  const double *lam, *mu, *W, *w, *dur, *thr;   // [M] [M] [M] [K*M] [K*M]=w/W [K]=new-job thresholds
  const int32_t *N, *lvl;                       // [M-1] capacity, [M-1] first level row of buffer b
  const uint8_t *by_rank;                       // [M] machines in Clojure hash-map order of their names
  // per-instance override planes, instance-fastest (nullptr = absent)
  const double *lam_i, *mu_i, *W_i, *w_i;
  const double *dur_i;    // [K*M][n_inst] w/W per instance, precomputed at upload when W_i or w_i is present (nullptr: divide in the loop)
  const int32_t *N_i;
  // outputs, instance-fastest (zero-initialised by the host before launch)
  int64_t *njobs; double *residence; double *blocked, *starved, *occ; double *final_clock;
  int64_t *curjob; uint64_t *n_events; uint64_t *hash; uint8_t *status;
  // scratch
  double *enters;         // [n_inst][R] :enters of the jobs in the line: a ring per instance, so that the 4 consecutive jobs
                          // of one 32-byte sector are written and read back while the sector is still in the L2
  double *cold_f;         // COLD variants: [5M-1][n_inst]  FUT_T NXT_T BS SS LASTCLK
  uint32_t *cold_u;       // COLD variants: [2M + R/4 + M][n_inst]  FUT_N STARTED JT EIDX
  // time slicing
  double slice_until;     // park every instance whose clock has passed this (+inf: run to the end)
  int32_t resume;         // 1: continue the instances parked by the previous launch
  int32_t park_words;     // rows of `park`
  uint32_t *park;         // [park_words][n_inst] parked per-instance state
  // rotation (see mjp_line_kernel.cuh): resident CTAs pull (chunk, slice) items; 0 slices = off
  int32_t rot_slices;     // slices of sim time the run is cut into
  uint32_t rot_chunks;    // chunks of blockDim instances
  double rot_dt;          // slice k ends at (k+1) rot_dt; the last one at slice_until
  uint32_t *rot_queue;    // [1] next item
  uint32_t *rot_flags;    // [rot_chunks] slices completed per chunk
  // MJP_RNG_TAPE: recorded draws, instance-major (nullptr: draw from the Philox substreams)
  const double *tape_jt, *tape_u, *tape_e;
  uint32_t tape_len_jt, tape_len_u, tape_len_e, tape_pad;
  // SCADA records
  void *log_records;      // mjp_log_record[log_capacity]
  uint64_t log_capacity;
  unsigned long long *log_cursor;   // [0] next free record, [1] records dropped
  uint32_t *first_state;            // LOG mode, optional: [n_inst][2M] see mjp_log_out.first_state (nullptr: not wanted)
  unsigned long long *stage;        // LOG mode: [n_inst][stage_cap] messages of the current tick (global scratch)
  // shared-memory layout
  int32_t started_hot;    // cold layouts: STARTED[M] lives in shared memory after CNT (1) or in the cold scratch (0)
  int32_t ends_cold;      // cold layouts: ENDS[M] lives in the cold scratch (rows 5M-1 .. 6M-2) instead of shared memory
  int32_t nf64, nu32;     // per-thread slots
  int32_t const_bytes;    // CTA constant table size (multiple of 16)
  int32_t stage_cap;      // LOG mode: messages staged per tick per instance
};

}  // namespace mjp
