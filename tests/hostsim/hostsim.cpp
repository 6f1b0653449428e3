// hostsim.cpp -- TEST-ONLY host execution of the kernel source (see cuda_shim.h).
#include "cuda_shim.h"
#define MJP_FIXED_NT 1        /* the shim runs one thread per block */
#include "../../mjpdes_b200/csrc/mjp_aux_kernels.cuh"
#include "../../mjpdes_b200/csrc/mjp_plan.h"

#include <limits>
#include <vector>

// Rotation (mjp_line_kernel.cuh): > 0 makes the next runs pull (chunk, slice) items from the work queue.  With one
// thread per block a single "CTA" drains the whole queue in order, which exercises restore / advance / park per item.
static int g_rot_slices = 0;
extern "C" void hostsim_set_rotation(int slices) { g_rot_slices = slices; }

// cold layouts: keep STARTED[M] in "shared memory" (what the host does when it costs no CTA per SM)
static int g_started_hot = 0;
extern "C" void hostsim_set_started_hot(int on) { g_started_hot = on; }

// One launch per slice boundary in `slices` (then one to the end), resuming the parked state each time.
extern "C" int hostsim_run_sliced(const mjp_line_desc *L, uint64_t n_inst, const mjp_rng_spec *rng, mjp_stats_out *s,
                                  int want_hash, mjp_log_out *log, int force_cold, const double *slices, int n_slices,
                                  const mjp_draw_tapes *tapes) {
  mjp::LinePlan pl = mjp::build_plan(L, n_inst);
  mjp::KParams kp = pl.kp;
  const int M = kp.M;
  kp.seed = rng->seed; kp.first_id = rng->first_instance_id;
  if (rng->mode == MJP_RNG_TAPE && tapes) {
    kp.tape_jt = tapes->jobtype; kp.tape_u = tapes->uniform; kp.tape_e = tapes->exp;
    kp.tape_len_jt = tapes->len_jobtype; kp.tape_len_u = tapes->len_uniform; kp.tape_len_e = tapes->len_exp;
  }
  const double *dc = pl.cd.data();
  kp.lam = dc + pl.off_lam(); kp.mu = dc + pl.off_mu(); kp.W = dc + pl.off_W(); kp.w = dc + pl.off_w();
  kp.dur = dc + pl.off_dur(); kp.thr = dc + pl.off_thr();
  kp.N = pl.ci.data(); kp.lvl = pl.ci.data() + (M - 1);
  kp.by_rank = pl.by_rank.data();
  kp.lam_i = L->lambda_inst; kp.mu_i = L->mu_inst; kp.W_i = L->W_inst; kp.w_i = L->w_inst; kp.N_i = L->N_inst;
  const bool want_log_early = log && L->log;
  const bool rot = g_rot_slices > 0 && n_slices == 0;
  const bool sweep = n_slices > 0 || rot;     // template slot 5 is SLICE; override planes are a run-time check
  std::vector<uint32_t> rotbuf(n_inst + 1, 0u);
  if (rot) {
    kp.rot_slices = g_rot_slices; kp.rot_chunks = (uint32_t)n_inst; kp.rot_dt = kp.run_to / g_rot_slices;
    kp.rot_queue = rotbuf.data(); kp.rot_flags = rotbuf.data() + 1;
  }
  std::vector<double> enters((size_t)kp.R * n_inst);
  kp.enters = enters.data();
  for (size_t q = 0; q < n_inst; ++q) s->status[q] = MJP_INST_NOT_RUN;
  for (size_t q = 0; q < (size_t)M * n_inst; ++q) { s->blocked_time[q] = 0; s->starved_time[q] = 0; }
  for (size_t q = 0; q < (size_t)kp.L * n_inst; ++q) s->occ_time[q] = 0;
  for (size_t q = 0; q < n_inst; ++q) { s->njobs[q] = 0; s->residence_sum[q] = 0; s->n_events[q] = 0; }
  const bool cold = force_cold != 0;
  const bool mt_ok = !cold && M >= 2 && M <= 8;
  kp.started_hot = (cold && g_started_hot) ? 1 : 0;
  kp.ends_cold = (cold && (M > 32 || g_started_hot)) ? 1 : 0;           // both placements get exercised
  mjp::plan_layout(kp, mt_ok, rng->mode != MJP_RNG_COUNTER, want_log_early, cold);
  std::vector<double> coldf((size_t)(mjp::cold_f64_slots(kp) + mjp::cold_ends_slots(kp)) * n_inst);
  std::vector<uint32_t> coldu((size_t)mjp::cold_u32_slots(kp, true) * n_inst);
  kp.cold_f = coldf.data(); kp.cold_u = coldu.data();
  kp.njobs = s->njobs; kp.residence = s->residence_sum; kp.blocked = s->blocked_time; kp.starved = s->starved_time;
  kp.occ = s->occ_time; kp.final_clock = s->final_clock; kp.curjob = s->current_job; kp.n_events = s->n_events;
  kp.hash = s->event_hash; kp.status = s->status;
  if (pl.kp.const_bytes + pl.per_thread_bytes > sizeof(mjp::smem_raw)) return MJP_ERR_UNSUPPORTED;
  const bool replay = rng->mode != MJP_RNG_COUNTER, hash = want_hash != 0;
  unsigned long long cursor[2] = {0, 0};
  const bool want_log = log && L->log;
  std::vector<unsigned long long> stage(want_log ? (size_t)kp.stage_cap * n_inst : 0);
  if (want_log) { kp.log_records = log->records; kp.log_capacity = log->capacity; kp.log_cursor = cursor; kp.stage = stage.data(); }
  if (want_log && log->first_state) {
    kp.first_state = log->first_state;
    for (size_t q = 0; q < (size_t)2 * M * n_inst; ++q) kp.first_state[q] = 0xFFFFFFFFu;
  }
  kp.park_words = 72 + 2 * kp.nf64 + kp.nu32;
  std::vector<uint32_t> park((size_t)kp.park_words * n_inst);
  kp.park = park.data();
  gridDim.x = (unsigned)n_inst;
  for (int sl = 0; sl <= n_slices; ++sl) {
  kp.slice_until = sl < n_slices ? slices[sl] : std::numeric_limits<double>::infinity();
  kp.resume = sl > 0;
  for (uint64_t g = 0; g < (rot ? 1 : n_inst); ++g) {
    blockIdx.x = (unsigned)g;
#ifdef MJP_HOSTSIM_MIN
    // the sanitizer build (asan_main.cpp) runs with the event hash on and without slicing: instantiating only those
    // kernels cuts its compile time by an order of magnitude
    if (!hash || sweep) return MJP_ERR_UNSUPPORTED;
#define RUN3(MK, MT, LG) \
    if (replay) mjp::mjp_line_kernel<MK, MT, true, true, false, LG>(kp); \
    else mjp::mjp_line_kernel<MK, MT, false, true, false, LG>(kp);
#define RUNC(MK, LG) \
    if (replay) mjp::mjp_line_kernel<MK, 0, true, true, false, LG, true>(kp); \
    else mjp::mjp_line_kernel<MK, 0, false, true, false, LG, true>(kp);
#else
#define RUN3(MK, MT, LG) \
    if (replay && hash && sweep) mjp::mjp_line_kernel<MK, MT, true, true, true, LG>(kp); \
    else if (replay && hash) mjp::mjp_line_kernel<MK, MT, true, true, false, LG>(kp); \
    else if (replay && sweep) mjp::mjp_line_kernel<MK, MT, true, false, true, LG>(kp); \
    else if (replay) mjp::mjp_line_kernel<MK, MT, true, false, false, LG>(kp); \
    else if (hash && sweep) mjp::mjp_line_kernel<MK, MT, false, true, true, LG>(kp); \
    else if (hash) mjp::mjp_line_kernel<MK, MT, false, true, false, LG>(kp); \
    else if (sweep) mjp::mjp_line_kernel<MK, MT, false, false, true, LG>(kp); \
    else mjp::mjp_line_kernel<MK, MT, false, false, false, LG>(kp);
    // the same dispatch as launch_line_kernel in mjp_abi.cu
#define RUNC(MK, LG) \
    if (replay && hash && sweep) mjp::mjp_line_kernel<MK, 0, true, true, true, LG, true>(kp); \
    else if (replay && hash) mjp::mjp_line_kernel<MK, 0, true, true, false, LG, true>(kp); \
    else if (replay && sweep) mjp::mjp_line_kernel<MK, 0, true, false, true, LG, true>(kp); \
    else if (replay) mjp::mjp_line_kernel<MK, 0, true, false, false, LG, true>(kp); \
    else if (hash && sweep) mjp::mjp_line_kernel<MK, 0, false, true, true, LG, true>(kp); \
    else if (hash) mjp::mjp_line_kernel<MK, 0, false, true, false, LG, true>(kp); \
    else if (sweep) mjp::mjp_line_kernel<MK, 0, false, false, true, LG, true>(kp); \
    else mjp::mjp_line_kernel<MK, 0, false, false, false, LG, true>(kp);
#endif
    if (cold) {
      if (M > 32) { if (want_log) { RUNC(uint64_t, true) } else { RUNC(uint64_t, false) } }
      else { if (want_log) { RUNC(uint32_t, true) } else { RUNC(uint32_t, false) } }
    }
    else if (M > 32) { if (want_log) { RUN3(uint64_t, 0, true) } else { RUN3(uint64_t, 0, false) } }
    else switch (M) {
#define RUNM(MT) case MT: if (want_log) { RUN3(uint32_t, MT, true) } else { RUN3(uint32_t, MT, false) } break;
#ifdef MJP_HOSTSIM_MIN
      RUNM(2) RUNM(5) RUNM(8)
#else
      RUNM(2) RUNM(3) RUNM(4) RUNM(5) RUNM(6) RUNM(7) RUNM(8)
#endif
#undef RUNM
      default: if (want_log) { RUN3(uint32_t, 0, true) } else { RUN3(uint32_t, 0, false) } break;
    }
  }
  }
  if (log) {
    log->count = cursor[0] < log->capacity ? cursor[0] : log->capacity;
    log->dropped = cursor[1];
  }
  if (s->aggregate) {
    const int n_metrics = 3 * M + 1, len = 1 + 2 * n_metrics;
    std::vector<double> partials(len);
    blockIdx.x = 0; gridDim.x = 1;
    mjp::mjp_aggregate_partial(kp, partials.data(), n_metrics);
    for (int q = 0; q < len; ++q) s->aggregate[q] = partials[q];
  }
  return MJP_OK;
}

extern "C" int hostsim_run(const mjp_line_desc *L, uint64_t n_inst, const mjp_rng_spec *rng, mjp_stats_out *s,
                           int want_hash, mjp_log_out *log, int force_cold) {
  return hostsim_run_sliced(L, n_inst, rng, s, want_hash, log, force_cold, nullptr, 0, nullptr);
}

extern "C" double hostsim_end_time(const uint8_t *kinds, const double *times, int n, double clock, double dur) {
  double out = 0;
  blockIdx.x = 0;
  mjp::mjp_end_time_probe(kinds, times, n, clock, dur, &out);
  return out;
}
