// mjp_abi.cu -- host side of the C ABI declared in include/mjpdes_b200.h.
//
// The handle owns every device buffer, a stream and two timing events; the
// caller owns every host buffer.  No call throws or aborts: errors come back
// as MJP_ERR_* with text in mjp_last_error().
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/mjpdes_b200.h"
#include "mjp_aux_kernels.cuh"
#include "mjp_launch.cuh"
#include "mjp_plan.h"

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&ptr, bytes ? bytes : 16);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

}  // namespace

struct mjp_handle {
  int device = 0;
  int flags = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t l2_persist_max = 0, l2_window_max = 0;   // persisting-L2 set-aside and access-policy window limits
  bool l2_persist_used = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  // uploaded line
  bool uploaded = false, launched = false, sweep = false;
  mjp::KParams kp{};
  // device buffers
  DevBuf d_const, d_planes, d_out, d_scratch, d_partials, d_agg, d_log, d_logcur, d_cold, d_park, d_tapes, d_stage, d_dur, d_rot, d_first;
  // out-plane offsets inside d_out (bytes)
  size_t o_njobs = 0, o_res = 0, o_blk = 0, o_stv = 0, o_occ = 0, o_clk = 0, o_cur = 0, o_nev = 0, o_hash = 0,
         o_status = 0, out_bytes = 0, zero_bytes = 0;
  int agg_blocks = 0;
  bool want_log = false;
  float last_ms = 0.f;
  int last_launches = 0;
  int block_threads = 0;
  size_t smem_bytes = 0;
  // what the launch that parked the current state looked like (a resume must match it)
  struct ParkKey { int mode = -1, log = 0, cold = 0, threads = 0; uint64_t seed = 0, first_id = 0; } park_key;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                \
      return e__ == cudaErrorMemoryAllocation ? MJP_ERR_NOMEM : MJP_ERR_CUDA;                      \
    }                                                                                              \
  } while (0)

static int fail(mjp_handle *h, int code, const char *msg) {
  if (h) h->err = msg;
  return code;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

/* the checks of core.clj:75-121 (clojure.spec) that concern the loop, plus the kernel's own limits */
static int validate(mjp_handle *h, const mjp_line_desc *L, uint64_t n_inst, uint64_t lo, uint64_t n_total) {
  if (!L) return fail(h, MJP_ERR_INVALID, "line descriptor is NULL");
  if (n_inst == 0) return fail(h, MJP_ERR_INVALID, "n_inst must be >= 1");
  if (n_inst >= (1ull << 31)) return fail(h, MJP_ERR_UNSUPPORTED, "2^31 or more instances per call");
  if (L->M < 2) return fail(h, MJP_ERR_INVALID, ":line needs at least two machines and a buffer (core.clj:120)");
  if (L->M > MJP_MAX_MACHINES) return fail(h, MJP_ERR_UNSUPPORTED, "more than 64 machines");
  if (L->K < 1) return fail(h, MJP_ERR_INVALID, ":jobmix is empty (core.clj:117)");
  if (L->K > MJP_MAX_JOBTYPES) return fail(h, MJP_ERR_UNSUPPORTED, "more than 255 job types");
  if (!L->lambda || !L->mu || !L->W || !L->discipline || !L->N || !L->portion || !L->w)
    return fail(h, MJP_ERR_INVALID, "a required array of the line descriptor is NULL");
  for (int i = 0; i < L->M; ++i) {
    if (!(L->lambda[i] >= 0.0) || !std::isfinite(L->lambda[i])) return fail(h, MJP_ERR_INVALID, ":lambda must be non-negative and finite (core.clj:89)");
    if (!(L->mu[i] > 0.0) || !std::isfinite(L->mu[i])) return fail(h, MJP_ERR_INVALID, ":mu must be positive and finite (core.clj:88)");
    if (!(L->W[i] > 0.0) || !std::isfinite(L->W[i])) return fail(h, MJP_ERR_INVALID, ":W must be positive and finite (core.clj:87)");
    if (L->discipline[i] > 1) return fail(h, MJP_ERR_INVALID, ":discipline must be :BAS or :BBS (core.clj:91)");
  }
  for (int b = 0; b < L->M - 1; ++b) {
    if (L->N[b] < 1) return fail(h, MJP_ERR_INVALID, ":N must be a positive integer (core.clj:95)");
    if (L->N[b] > 254) return fail(h, MJP_ERR_UNSUPPORTED, "buffer capacity above 254");
  }
  for (int k = 0; k < L->K; ++k)
    if (!(L->portion[k] > 0.0) || !std::isfinite(L->portion[k])) return fail(h, MJP_ERR_INVALID, ":portion must be positive and finite (core.clj:109)");
  for (int q = 0; q < L->K * L->M; ++q)
    if (!(L->w[q] > 0.0) || !std::isfinite(L->w[q])) return fail(h, MJP_ERR_INVALID, ":w values must be positive and finite (core.clj:108)");
  if (!(L->warm_up >= 0.0) || !std::isfinite(L->warm_up)) return fail(h, MJP_ERR_INVALID, ":warm-up-time must be non-negative and finite (core.clj:104)");
  // +inf is allowed here (main-loop-continuous never ends, core.clj:801-843); mjp_launch_slice refuses to run it unsliced
  if (!(L->run_to > 0.0)) return fail(h, MJP_ERR_INVALID, ":run-to-time must be positive (core.clj:105)");
  if (L->group_rank) {
    uint64_t seen = 0;
    for (int i = 0; i < L->M; ++i) {
      const int r = L->group_rank[i];
      if (r < 0 || r >= L->M || ((seen >> r) & 1)) return fail(h, MJP_ERR_INVALID, "group_rank must be a permutation of 0..M-1");
      seen |= 1ull << r;
    }
  }
  // per-instance planes are [rows][n_total]; this call covers columns lo .. lo + n_inst
  auto plane_ok = [&](const double *pl, int rows, bool allow_zero) {
    for (int r = 0; r < rows; ++r)
      for (uint64_t g = 0; g < n_inst; ++g) {
        const double v = pl[(size_t)r * n_total + lo + g];
        if (!(allow_zero ? v >= 0.0 : v > 0.0) || !std::isfinite(v)) return false;
      }
    return true;
  };
  if (L->N_inst)
    for (int b = 0; b < L->M - 1; ++b)
      for (uint64_t g = 0; g < n_inst; ++g) {
        const int v = L->N_inst[(size_t)b * n_total + lo + g];
        if (v < 1 || v > L->N[b]) return fail(h, MJP_ERR_INVALID, "N_inst value outside 1..N[b]");
      }
  if (L->lambda_inst && !plane_ok(L->lambda_inst, L->M, true)) return fail(h, MJP_ERR_INVALID, "lambda_inst must be non-negative and finite");
  if (L->mu_inst && !plane_ok(L->mu_inst, L->M, false)) return fail(h, MJP_ERR_INVALID, "mu_inst must be positive and finite");
  if (L->W_inst && !plane_ok(L->W_inst, L->M, false)) return fail(h, MJP_ERR_INVALID, "W_inst must be positive and finite");
  if (L->w_inst && !plane_ok(L->w_inst, L->K * L->M, false)) return fail(h, MJP_ERR_INVALID, "w_inst must be positive and finite");
  return MJP_OK;
}

static int upload_impl(mjp_handle *h, const mjp_line_desc *L, uint64_t n_inst, uint64_t lo, uint64_t n_total);
static int download_impl(mjp_handle *h, mjp_stats_out *s, mjp_log_out *log, uint64_t lo, uint64_t n_total);

static bool use_register_ends(const mjp::KParams &kp, bool log, int threads) {
  static const bool force_generic = std::getenv("MJP_FORCE_GENERIC") != nullptr;   // A/B measurements only
  (void)log;
  return !force_generic && threads == MJP_FIXED_NT && kp.M >= 2 && kp.M <= 8;
}

static cudaError_t launch_line_kernel(bool replay, bool hash, bool slice, bool log, bool cold, const mjp::KParams &kp,
                                      int blocks, int threads, size_t smem, cudaStream_t st, int *occupancy = nullptr) {
#define MJP_GO(T, MT, COLD)                                                                          \
  return log ? mjp::launch_flags<T, MT, COLD, true>(replay, hash, slice, kp, blocks, threads, smem, st, occupancy) \
             : mjp::launch_flags<T, MT, COLD, false>(replay, hash, slice, kp, blocks, threads, smem, st, occupancy)
  if (cold) { if (kp.M > 32) MJP_GO(uint64_t, 0, true); MJP_GO(uint32_t, 0, true); }
  if (kp.M > 32) MJP_GO(uint64_t, 0, false);
  if (use_register_ends(kp, log, threads)) {
    switch (kp.M) {
      case 2: MJP_GO(uint32_t, 2, false);
      case 3: MJP_GO(uint32_t, 3, false);
      case 4: MJP_GO(uint32_t, 4, false);
      case 5: MJP_GO(uint32_t, 5, false);
      case 6: MJP_GO(uint32_t, 6, false);
      case 7: MJP_GO(uint32_t, 7, false);
      case 8: MJP_GO(uint32_t, 8, false);
      default: break;
    }
  }
  MJP_GO(uint32_t, 0, false);
#undef MJP_GO
}

extern "C" {

int mjp_abi_version(void) { return MJP_ABI_VERSION; }

int mjp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int mjp_aggregate_len(int32_t M) { return 1 + 2 * (3 * M + 1); }

int mjp_create(const mjp_config *cfg, mjp_handle **out) {
  if (!out) return MJP_ERR_INVALID;
  *out = nullptr;
  mjp_handle *h = new (std::nothrow) mjp_handle();
  if (!h) return MJP_ERR_NOMEM;
  h->device = cfg ? cfg->device : 0;
  h->flags = cfg ? cfg->flags : 0;
  cudaError_t e = cudaSetDevice(h->device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, h->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e != cudaSuccess) {
    g_create_error = std::string("mjp_create: ") + cudaGetErrorString(e);
    delete h;
    return MJP_ERR_CUDA;
  }
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
  h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
  *out = h;
  return MJP_OK;
}

void mjp_destroy(mjp_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (DevBuf *b : {&h->d_const, &h->d_planes, &h->d_out, &h->d_scratch, &h->d_partials, &h->d_agg, &h->d_log,
                    &h->d_logcur, &h->d_cold, &h->d_park, &h->d_tapes, &h->d_stage, &h->d_dur, &h->d_rot, &h->d_first})
    b->release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *mjp_last_error(const mjp_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mjp_upload(mjp_handle *h, const mjp_line_desc *L, uint64_t n_inst) { return upload_impl(h, L, n_inst, 0, n_inst); }

}  // extern "C"

// Upload instances [lo, lo + n_inst) of a batch of n_total: the per-instance override planes of the descriptor are
// [rows][n_total] and only this window of columns is copied (2-D copies); mjp_upload is the window [0, n).
static int upload_impl(mjp_handle *h, const mjp_line_desc *L, uint64_t n_inst, uint64_t lo, uint64_t n_total) {
  if (!h) return MJP_ERR_INVALID;
  h->uploaded = false; h->launched = false;
  int rc = validate(h, L, n_inst, lo, n_total);
  if (rc != MJP_OK) return rc;
  CK(cudaSetDevice(h->device));
  const int M = L->M, K = L->K;
  mjp::LinePlan pl = mjp::build_plan(L, n_inst);
  mjp::KParams &kp = h->kp;
  kp = pl.kp;
  const int R = kp.R, levels = kp.L;

  // ---- constants: lam mu W w dur thr | N lvl ----
  const size_t dbytes = pl.cd.size() * sizeof(double), ibytes = pl.ci.size() * sizeof(int32_t);
  CK(h->d_const.reserve(dbytes + ibytes + pl.by_rank.size()));
  CK(cudaMemcpyAsync(h->d_const.ptr, pl.cd.data(), dbytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync((char *)h->d_const.ptr + dbytes, pl.ci.data(), ibytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync((char *)h->d_const.ptr + dbytes + ibytes, pl.by_rank.data(), pl.by_rank.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // pl is stack-owned
  const double *dc = h->d_const.as<double>();
  kp.lam = dc + pl.off_lam(); kp.mu = dc + pl.off_mu(); kp.W = dc + pl.off_W(); kp.w = dc + pl.off_w();
  kp.dur = dc + pl.off_dur(); kp.thr = dc + pl.off_thr();
  kp.N = reinterpret_cast<const int32_t *>((const char *)h->d_const.ptr + dbytes);
  kp.lvl = kp.N + (M - 1);
  kp.by_rank = reinterpret_cast<const uint8_t *>((const char *)h->d_const.ptr + dbytes + ibytes);

  // ---- per-instance planes ----
  const size_t pM = align_up(n_inst * M * sizeof(double), 256), pKM = align_up(n_inst * (size_t)K * M * sizeof(double), 256);
  const size_t pN = align_up(n_inst * (size_t)(M - 1) * sizeof(int32_t), 256);
  size_t need = (L->lambda_inst ? pM : 0) + (L->mu_inst ? pM : 0) + (L->W_inst ? pM : 0) + (L->w_inst ? pKM : 0) +
                (L->N_inst ? pN : 0);
  h->sweep = need != 0;
  if (need) {
    CK(h->d_planes.reserve(need));
    char *base = (char *)h->d_planes.ptr;
    size_t off = 0;
    auto put = [&](const void *src, size_t elem, size_t rows, size_t padded) -> const void * {
      if (!src) return nullptr;
      void *dst = base + off;
      cudaMemcpy2DAsync(dst, n_inst * elem, (const char *)src + lo * elem, n_total * elem, n_inst * elem, rows,
                        cudaMemcpyHostToDevice, h->stream);
      off += padded;
      return dst;
    };
    kp.lam_i = (const double *)put(L->lambda_inst, sizeof(double), M, pM);
    kp.mu_i = (const double *)put(L->mu_inst, sizeof(double), M, pM);
    kp.W_i = (const double *)put(L->W_inst, sizeof(double), M, pM);
    kp.w_i = (const double *)put(L->w_inst, sizeof(double), (size_t)K * M, pKM);
    kp.N_i = (const int32_t *)put(L->N_inst, sizeof(int32_t), (size_t)(M - 1), pN);
    CK(cudaGetLastError());
  }
  kp.dur_i = nullptr;
  if (kp.W_i || kp.w_i) {
    // sweeps over W / w: divide w by W once per (job type, machine, instance) instead of at every job start;
    // skipped (the loop then divides) when the plane would be unreasonably large
    const size_t bytes = (size_t)K * M * n_inst * sizeof(double);
    if (bytes <= ((size_t)4 << 30)) {
      CK(h->d_dur.reserve(bytes));
      const int blocks = (int)std::min<size_t>(((size_t)K * M * n_inst + 255) / 256, (size_t)h->sm_count * 16);
      mjp::mjp_dur_plane<<<blocks, 256, 0, h->stream>>>(kp, h->d_dur.as<double>());
      CK(cudaGetLastError());
      kp.dur_i = h->d_dur.as<double>();
    }
  }

  // ---- outputs: [blocked | starved | occ] are accumulated into and must be zeroed per launch ----
  const size_t n = n_inst;
  size_t off = 0;
  auto plane = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  h->o_blk = plane(n * M * 8); h->o_stv = plane(n * M * 8); h->o_occ = plane(n * (size_t)levels * 8);
  h->o_njobs = plane(n * 8); h->o_res = plane(n * 8); h->o_nev = plane(n * 8);
  h->zero_bytes = off;
  h->o_clk = plane(n * 8); h->o_cur = plane(n * 8);
  h->o_hash = plane(n * 8); h->o_status = plane(n);
  h->out_bytes = off;
  CK(h->d_out.reserve(off));
  char *ob = (char *)h->d_out.ptr;
  kp.blocked = (double *)(ob + h->o_blk); kp.starved = (double *)(ob + h->o_stv); kp.occ = (double *)(ob + h->o_occ);
  kp.njobs = (int64_t *)(ob + h->o_njobs); kp.residence = (double *)(ob + h->o_res);
  kp.final_clock = (double *)(ob + h->o_clk); kp.curjob = (int64_t *)(ob + h->o_cur);
  kp.n_events = (uint64_t *)(ob + h->o_nev); kp.hash = (uint64_t *)(ob + h->o_hash);
  kp.status = (uint8_t *)(ob + h->o_status);
  CK(h->d_scratch.reserve(n * (size_t)R * 8));
  kp.enters = h->d_scratch.as<double>();

  // the launch shape (layout, CTA size) depends on the kernel variant and is chosen in mjp_launch_slice
  // aggregate scratch
  h->agg_blocks = h->sm_count > 0 ? h->sm_count : 148;
  const int len = mjp_aggregate_len(M);
  CK(h->d_partials.reserve((size_t)h->agg_blocks * len * 8));
  CK(h->d_agg.reserve((size_t)len * 8));
  CK(h->d_logcur.reserve(16));
  kp.log_cursor = h->d_logcur.as<unsigned long long>();
  CK(cudaStreamSynchronize(h->stream));
  h->uploaded = true;
  return MJP_OK;
}

extern "C" {

int mjp_reserve_log(mjp_handle *h, uint64_t capacity) {
  if (!h) return MJP_ERR_INVALID;
  if (!h->uploaded) return fail(h, MJP_ERR_INVALID, "mjp_reserve_log before mjp_upload");
  CK(cudaSetDevice(h->device));
  CK(h->d_log.reserve((capacity ? capacity : 1) * sizeof(mjp_log_record)));
  h->kp.log_records = h->d_log.ptr;
  h->kp.log_capacity = capacity;
  h->kp.log_cursor = h->d_logcur.as<unsigned long long>();
  return MJP_OK;
}

int mjp_set_tapes(mjp_handle *h, const mjp_draw_tapes *t) {
  if (!h) return MJP_ERR_INVALID;
  if (!h->uploaded) return fail(h, MJP_ERR_INVALID, "mjp_set_tapes before mjp_upload");
  if (!t || !t->jobtype || !t->uniform || !t->exp || !t->len_jobtype || !t->len_uniform || !t->len_exp)
    return fail(h, MJP_ERR_INVALID, "mjp_draw_tapes: all three tapes are required and must not be empty");
  CK(cudaSetDevice(h->device));
  mjp::KParams &kp = h->kp;
  const size_t n = kp.n_inst, M = kp.M;
  const size_t bj = align_up(n * t->len_jobtype * 8, 256), bu = align_up(n * M * t->len_uniform * 8, 256),
               be = n * M * t->len_exp * 8;
  CK(h->d_tapes.reserve(bj + bu + be));
  char *base = (char *)h->d_tapes.ptr;
  CK(cudaMemcpyAsync(base, t->jobtype, n * t->len_jobtype * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(base + bj, t->uniform, n * M * t->len_uniform * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(base + bj + bu, t->exp, be, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  kp.tape_jt = (const double *)base; kp.tape_u = (const double *)(base + bj); kp.tape_e = (const double *)(base + bj + bu);
  kp.tape_len_jt = t->len_jobtype; kp.tape_len_u = t->len_uniform; kp.tape_len_e = t->len_exp;
  return MJP_OK;
}

int mjp_launch(mjp_handle *h, const mjp_rng_spec *rng, int want_log) {
  return mjp_launch_slice(h, rng, want_log, std::numeric_limits<double>::infinity(), 0);
}

int mjp_launch_slice(mjp_handle *h, const mjp_rng_spec *rng, int want_log, double until, int resume) {
  if (!h) return MJP_ERR_INVALID;
  if (!h->uploaded) return fail(h, MJP_ERR_INVALID, "mjp_launch before mjp_upload");
  if (resume && !h->launched) return fail(h, MJP_ERR_INVALID, "resume without a previous sliced launch");
  if (!(until == until)) return fail(h, MJP_ERR_INVALID, "slice boundary is NaN");
  if (!rng || rng->mode < MJP_RNG_COUNTER || rng->mode > MJP_RNG_TAPE)
    return fail(h, MJP_ERR_INVALID, "bad mjp_rng_spec");
  if (rng->mode == MJP_RNG_TAPE && !h->kp.tape_u)
    return fail(h, MJP_ERR_INVALID, "MJP_RNG_TAPE: call mjp_set_tapes after mjp_upload first");
  // the kernel's only exits are the horizon, :run-to-job and the slice boundary: without any of them it would
  // spin for ever on a healthy line (main-loop-continuous, core.clj:801-843, must be driven in slices)
  if (until == std::numeric_limits<double>::infinity() && !std::isfinite(h->kp.run_to) && h->kp.run_to_job < 0)
    return fail(h, MJP_ERR_INVALID, "endless run: :run-to-time is infinite, no :run-to-job, no slice boundary");
  CK(cudaSetDevice(h->device));
  mjp::KParams &kp = h->kp;
  kp.seed = rng->seed;
  kp.first_id = rng->first_instance_id;
  h->want_log = want_log && kp.log_on;
  int launches = 0;
  kp.slice_until = until;
  kp.resume = resume ? 1 : 0;
  if (!resume) {                                  // a resumed slice keeps accumulating into the same planes
    CK(cudaMemsetAsync(h->d_out.ptr, 0, h->zero_bytes, h->stream));
    CK(cudaMemsetAsync((char *)h->d_out.ptr + h->o_status, MJP_INST_NOT_RUN, kp.n_inst, h->stream));
  }
  CK(cudaMemsetAsync(h->d_logcur.ptr, 0, 16, h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  const bool replay = rng->mode != MJP_RNG_COUNTER, hash = (h->flags & MJP_CFG_EVENT_HASH) != 0;
  if (h->want_log && !kp.log_records) return fail(h, MJP_ERR_INVALID, "SCADA capture wanted: call mjp_reserve_log first");
  {
    // Launch shape.  64-thread CTAs spread a batch of equally long instances evenly over the SMs (the step
    // time is set by the fullest SM).  Wide lines: when shared memory would hold fewer than 6 such CTAs per SM
    // AND the batch needs more than one wave, only the per-micro-step state stays on chip and the rest goes
    // to an L2-resident scratch (COLD variants); otherwise the extra latency is not worth it.  If nothing
    // else fits, COLD and/or 32-thread CTAs are the fallback.
    static const char *cold_env = std::getenv("MJP_COLD");            // A/B measurements: 0 = never, 1 = always
    const size_t sms = (size_t)(h->sm_count > 0 ? h->sm_count : 148);
    kp.started_hot = 0;
    static const char *ec_env = std::getenv("MJP_ENDS_COLD");             // A/B measurements: 0 / 1 force it
    kp.ends_cold = ec_env ? (ec_env[0] == '1') : (kp.M > 32);
    auto cta_bytes = [&](bool c, int t) {
      mjp::KParams tmp = kp;
      const bool mt_ = !c && use_register_ends(kp, h->want_log, t);
      const size_t per_thread_ = mjp::plan_layout(tmp, mt_, replay, h->want_log, c);       // (sets tmp.const_bytes too)
      return (size_t)tmp.const_bytes + per_thread_ * t;
    };
    auto fits = [&](bool c, int t) { return cta_bytes(c, t) <= h->smem_optin; };
    int threads = 0;
    bool cold = false;
    if (fits(false, 64)) {
      threads = 64;
      const size_t per_sm = h->smem_optin / (cta_bytes(false, 64) + 1024);
      const bool register_ends = use_register_ends(kp, h->want_log, 64);
      cold = !register_ends && per_sm < 6 && (kp.n_inst + 63) / 64 > sms * per_sm && fits(true, 64);
      if (cold_env) cold = !register_ends && cold_env[0] == '1' && fits(true, 64);
    } else if (fits(true, 64)) { threads = 64; cold = true; }
    else if (fits(false, 32)) { threads = 32; }
    else if (fits(true, 32)) { threads = 32; cold = true; }
    else return fail(h, MJP_ERR_UNSUPPORTED, "line state does not fit in shared memory (M, sum N too large)");
    h->block_threads = threads;
    if (cold) {
      // STARTED[M] (read by every job movement) stays in shared memory when that costs no CTA per SM: the kernels run
      // at most 12 (32-bit masks, statistics only) / 8 (64-bit masks, or with capture) CTAs per SM for their registers
      const size_t reg_limit = (kp.M > 32 || h->want_log) ? 8 : 12;      // line_kernel_min_ctas
      const size_t without = std::min(reg_limit, h->smem_optin / (cta_bytes(true, threads) + 1024));
      mjp::KParams tmp = kp;
      tmp.started_hot = 1;
      const size_t with_thread = mjp::plan_layout(tmp, false, replay, h->want_log, true);
      const size_t with_bytes = (size_t)tmp.const_bytes + with_thread * threads;
      const size_t with = std::min(reg_limit, h->smem_optin / (with_bytes + 1024));
      kp.started_hot = (with >= without && with_bytes <= h->smem_optin) ? 1 : 0;
    }
    {
      // the park layout follows from the variant: a resume must be the same kind of launch on the same substreams
      mjp_handle::ParkKey key;
      key.mode = rng->mode; key.log = h->want_log; key.cold = cold; key.threads = threads;
      key.seed = rng->seed; key.first_id = rng->first_instance_id;
      const mjp_handle::ParkKey &k0 = h->park_key;
      if (resume && (k0.mode != key.mode || k0.log != key.log || k0.cold != key.cold || k0.threads != key.threads ||
                     k0.seed != key.seed || k0.first_id != key.first_id))
        return fail(h, MJP_ERR_INVALID, "resume does not match the launch that parked the state (rng mode, seed, "
                                        "first_instance_id and want_log must be the same)");
      h->park_key = key;
    }
    const bool mt = !cold && use_register_ends(kp, h->want_log, threads);
    const size_t per_thread = mjp::plan_layout(kp, mt, replay, h->want_log, cold);
    if (cold) {
      const size_t fb = mjp::plan_align_up((size_t)(mjp::cold_f64_slots(kp) + mjp::cold_ends_slots(kp)) * kp.n_inst * 8, 256);
      CK(h->d_cold.reserve(fb + (size_t)mjp::cold_u32_slots(kp, replay) * kp.n_inst * 4));
      kp.cold_f = h->d_cold.as<double>();
      kp.cold_u = reinterpret_cast<uint32_t *>((char *)h->d_cold.ptr + fb);
    }
    if (h->want_log) {                              // one tick of staged messages per instance; persists across slices
      CK(h->d_stage.reserve((size_t)kp.stage_cap * kp.n_inst * sizeof(unsigned long long)));
      kp.stage = h->d_stage.as<unsigned long long>();
      // where the jobs were at each instance's first printed message (mjp_log_out.first_state); persists across slices
      const size_t first_bytes = (size_t)2 * kp.M * kp.n_inst * sizeof(uint32_t);
      CK(h->d_first.reserve(first_bytes));
      kp.first_state = h->d_first.as<uint32_t>();
      if (!resume) CK(cudaMemsetAsync(h->d_first.ptr, 0xFF, first_bytes, h->stream));
    }
    h->smem_bytes = (size_t)kp.const_bytes + per_thread * threads;
    int blocks = (int)((kp.n_inst + threads - 1) / threads);
    // Rotation: a complete run of a batch that needs between one and four waves of CTAs is cut into slices of
    // sim time, and the resident CTAs pull (chunk, slice) items from a queue (mjp_line_kernel.cuh): all instances
    // of a batch take about the same time, so otherwise the last, partly filled wave would cost a full pass.
    kp.rot_slices = 0;
    static const char *rot_env = std::getenv("MJP_ROTATE");               // A/B measurements: 0 = never
    static const char *rot_slices_env = std::getenv("MJP_ROTATE_SLICES");
    if (!resume && until == std::numeric_limits<double>::infinity() && std::isfinite(kp.run_to) &&
        !(rot_env && rot_env[0] == '0')) {
      int occ = 0;
      cudaError_t oe = launch_line_kernel(replay, hash, true, h->want_log, cold, kp, 0, threads, h->smem_bytes, h->stream, &occ);
      const long long capacity = (oe == cudaSuccess) ? (long long)occ * (long long)sms : 0;
      if (capacity > 0 && blocks > capacity && blocks < 4 * capacity) {
        kp.rot_slices = rot_slices_env ? std::max(2, std::atoi(rot_slices_env)) : 16;
        kp.rot_chunks = (uint32_t)blocks;
        kp.rot_dt = kp.run_to / (double)kp.rot_slices;
        CK(h->d_rot.reserve(4 * ((size_t)blocks + 1)));
        CK(cudaMemsetAsync(h->d_rot.ptr, 0, 4 * ((size_t)blocks + 1), h->stream));
        kp.rot_queue = h->d_rot.as<uint32_t>();
        kp.rot_flags = h->d_rot.as<uint32_t>() + 1;
        blocks = (int)capacity;
      }
      cudaGetLastError();
    }
    const bool rotate = kp.rot_slices > 0;
    if (rotate || resume || until < std::numeric_limits<double>::infinity()) {   // == slice
      kp.park_words = 72 + 2 * kp.nf64 + kp.nu32;     // loop-carried registers (<= 72 words) + every shared slot
      CK(h->d_park.reserve((size_t)kp.park_words * kp.n_inst * 4));
      kp.park = h->d_park.as<uint32_t>();
    }
    const bool slice = rotate || resume || until < std::numeric_limits<double>::infinity();
    mjp::KParams kl = kp;                         // tapes are read by MJP_RNG_TAPE launches only
    if (rng->mode != MJP_RNG_TAPE) kl.tape_jt = kl.tape_u = kl.tape_e = nullptr;
    // COLD layouts: the cold per-instance state is re-read all through the run while statistics, job-entry times
    // and SCADA records stream past it; pin it in the L2 (persisting access-policy window on the handle's stream)
    static const char *persist_env = std::getenv("MJP_L2_PERSIST");       // A/B measurements: 1 = on
    // (measured on C3 / C4 / C5 and both SCADA shapes: no gain, 0-3 % loss -- kept as an opt-in, MJP_L2_PERSIST=1)
    const bool persist = cold && h->l2_persist_max && h->l2_window_max && persist_env && persist_env[0] == '1';
    if (persist) {
      const size_t cold_bytes = std::min(h->d_cold.cap, h->l2_window_max);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_max);
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.base_ptr = h->d_cold.ptr;
      av.accessPolicyWindow.num_bytes = cold_bytes;
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)h->l2_persist_max / (double)cold_bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
      cudaGetLastError();                                                  // best effort: a refusal is not an error
    }
    cudaError_t e = launch_line_kernel(replay, hash, slice, h->want_log, cold, kl, blocks, threads, h->smem_bytes, h->stream);
    if (persist) {
      h->l2_persist_used = true;
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.num_bytes = 0;
      cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
      cudaGetLastError();
    }
    CK(e);
    launches++;
  }
  {
    const int n_metrics = 3 * kp.M + 1;
    mjp::mjp_aggregate_partial<<<h->agg_blocks, 256, 2 * 256 * sizeof(double), h->stream>>>(
        kp, h->d_partials.as<double>(), n_metrics);
    CK(cudaGetLastError());
    const int len = 1 + 2 * n_metrics;
    mjp::mjp_aggregate_final<<<(len + 127) / 128, 128, 0, h->stream>>>(h->d_partials.as<double>(), h->agg_blocks, len,
                                                                      h->d_agg.as<double>());
    CK(cudaGetLastError());
    launches += 2;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->last_launches = launches;
  h->launched = true;
  return MJP_OK;
}

int mjp_sync(mjp_handle *h) {
  if (!h) return MJP_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->l2_persist_used) { cudaCtxResetPersistingL2Cache(); h->l2_persist_used = false; }   // hand the set-aside L2 back
  if (h->launched) CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  return MJP_OK;
}

int mjp_download(mjp_handle *h, mjp_stats_out *s, mjp_log_out *log) {
  return download_impl(h, s, log, 0, h ? h->kp.n_inst : 0);
}

int mjp_device_aggregate(mjp_handle *h, double **dptr, int32_t *len) {
  if (!h || !dptr) return MJP_ERR_INVALID;
  if (!h->launched) return fail(h, MJP_ERR_INVALID, "mjp_device_aggregate before mjp_launch");
  *dptr = h->d_agg.as<double>();
  if (len) *len = mjp_aggregate_len(h->kp.M);
  return MJP_OK;
}

}  // extern "C"

// Copy the results of this handle's instances into columns lo .. lo + n of caller planes that are n_total wide
// (mjp_download: lo = 0, n_total = n).
static int download_impl(mjp_handle *h, mjp_stats_out *s, mjp_log_out *log, uint64_t lo, uint64_t n_total) {
  if (!h) return MJP_ERR_INVALID;
  if (!h->launched) return fail(h, MJP_ERR_INVALID, "mjp_download before mjp_launch");
  CK(cudaSetDevice(h->device));
  const mjp::KParams &kp = h->kp;
  const size_t n = kp.n_inst;
  const char *ob = (const char *)h->d_out.ptr;
  auto get = [&](void *dst, size_t o, size_t elem, size_t rows) {
    if (!dst) return;
    if (n_total == n) cudaMemcpyAsync(dst, ob + o, n * elem * rows, cudaMemcpyDeviceToHost, h->stream);
    else cudaMemcpy2DAsync((char *)dst + lo * elem, n_total * elem, ob + o, n * elem, n * elem, rows,
                           cudaMemcpyDeviceToHost, h->stream);
  };
  if (s) {
    get(s->njobs, h->o_njobs, 8, 1); get(s->residence_sum, h->o_res, 8, 1);
    get(s->blocked_time, h->o_blk, 8, kp.M); get(s->starved_time, h->o_stv, 8, kp.M);
    get(s->occ_time, h->o_occ, 8, (size_t)kp.L); get(s->final_clock, h->o_clk, 8, 1);
    get(s->current_job, h->o_cur, 8, 1); get(s->n_events, h->o_nev, 8, 1);
    if (s->event_hash) {
      if (h->flags & MJP_CFG_EVENT_HASH) get(s->event_hash, h->o_hash, 8, 1);
      else std::memset(s->event_hash + lo, 0, n * 8);
    }
    get(s->status, h->o_status, 1, 1);
    if (s->aggregate)
      cudaMemcpyAsync(s->aggregate, h->d_agg.ptr, (size_t)mjp_aggregate_len(kp.M) * 8, cudaMemcpyDeviceToHost, h->stream);
  }
  int rc = MJP_OK;
  if (log) {
    log->count = 0; log->dropped = 0;
    if (!h->want_log && log->first_state)             // no capture ran: "nothing printed yet" for every instance
      std::memset(log->first_state + lo * 2 * (size_t)kp.M, 0xFF, (size_t)2 * kp.M * n * sizeof(uint32_t));
    if (h->want_log) {
      unsigned long long cur[2] = {0, 0};
      CK(cudaMemcpyAsync(cur, h->d_logcur.ptr, 16, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      uint64_t have = cur[0] < kp.log_capacity ? cur[0] : kp.log_capacity;
      uint64_t dropped = cur[1] + (cur[0] > kp.log_capacity ? cur[0] - kp.log_capacity : 0);
      uint64_t take = have < log->capacity ? have : log->capacity;
      if (take && log->records)
        CK(cudaMemcpyAsync(log->records, h->d_log.ptr, take * sizeof(mjp_log_record), cudaMemcpyDeviceToHost, h->stream));
      log->count = take;
      log->dropped = dropped + (have - take);
      if (log->dropped) rc = MJP_LOG_TRUNCATED;
      if (log->first_state)
        CK(cudaMemcpyAsync(log->first_state + lo * 2 * (size_t)kp.M, h->d_first.ptr, (size_t)2 * kp.M * n * sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, h->stream));
    }
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return rc;
}

extern "C" {

int mjp_run(mjp_handle *h, const mjp_line_desc *line, uint64_t n_inst, const mjp_rng_spec *rng,
            mjp_stats_out *stats, mjp_log_out *log) {
  int rc = mjp_upload(h, line, n_inst);
  if (rc != MJP_OK) return rc;
  if (log && line->log) {
    rc = mjp_reserve_log(h, log->capacity);
    if (rc != MJP_OK) return rc;
  }
  rc = mjp_launch(h, rng, log != nullptr);
  if (rc != MJP_OK) return rc;
  rc = mjp_sync(h);
  if (rc != MJP_OK) return rc;
  return mjp_download(h, stats, log);
}

int mjp_last_kernel_ms(mjp_handle *h, float *ms) {
  if (!h || !ms) return MJP_ERR_INVALID;
  *ms = h->last_ms;
  return MJP_OK;
}

int mjp_last_launch_count(mjp_handle *h, int *n) {
  if (!h || !n) return MJP_ERR_INVALID;
  *n = h->last_launches;
  return MJP_OK;
}

int mjp_probe_end_time(mjp_handle *h, const uint8_t *kinds, const double *times, int32_t n_events, double clock,
                       double dur, double *out) {
  if (!h || !kinds || !times || !out || n_events < 1) return MJP_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  DevBuf buf;
  const size_t tb = align_up((size_t)n_events * 8, 16);
  cudaError_t e = buf.reserve(tb + align_up(n_events, 16) + 16);
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return MJP_ERR_NOMEM; }
  char *base = (char *)buf.ptr;
  cudaMemcpyAsync(base, times, (size_t)n_events * 8, cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(base + tb, kinds, n_events, cudaMemcpyHostToDevice, h->stream);
  double *d_out = (double *)(base + tb + align_up(n_events, 16));
  mjp::mjp_end_time_probe<<<1, 32, 0, h->stream>>>((const uint8_t *)(base + tb), (const double *)base, n_events, clock,
                                                   dur, d_out);
  cudaMemcpyAsync(out, d_out, 8, cudaMemcpyDeviceToHost, h->stream);
  e = cudaStreamSynchronize(h->stream);
  buf.release();
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return MJP_ERR_CUDA; }
  return MJP_OK;
}

}  // extern "C"

// ================= multi-GPU: main-loop-multi's fan-out (core.clj:753-766) over the GPUs of one box =================
// One mjp_handle and one host thread per device; the batch is cut into contiguous ranges of instance ids (an
// instance's random substreams depend on its id only, so the results do not depend on the number of devices).
// There is no data-path exchange.  The ONE collective is the sum of the per-device aggregate vectors:
//   MJP_REDUCE_PEER  device 0 pulls every vector over NVLink (cudaMemcpyPeerAsync) and sums them in device order
//                    (deterministic: the result does not depend on timing);
//   MJP_REDUCE_NCCL  ncclAllReduce(sum, f64) over communicators made by ncclCommInitAll; libnccl.so.2 is bound at
//                    run time (dlopen), so the library has no link-time dependency on NCCL.
#include <dlfcn.h>

#include <thread>

namespace {

struct NcclApi {
  void *lib = nullptr;
  int (*CommInitAll)(void **comms, int ndev, const int *devlist) = nullptr;
  int (*CommDestroy)(void *comm) = nullptr;
  int (*AllReduce)(const void *send, void *recv, size_t count, int dtype, int op, void *comm, cudaStream_t st) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load(std::string &err) {
    if (lib) return true;
    // 1. the NCCL the host process has loaded already (a JVM or Python host that uses NCCL itself -- e.g. torch's
    //    bundled copy -- must not get a second, different libnccl.so.2 mapped under the same soname);
    // 2. the path in $MJP_NCCL_LIB;  3. the loader's default libnccl.so.2.  Symbols stay private (RTLD_LOCAL).
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
    if (!lib) if (const char *path = std::getenv("MJP_NCCL_LIB")) lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      if (lib) break;
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!lib) { err = std::string("MJP_REDUCE_NCCL: cannot load libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char *n) { return dlsym(lib, n); };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd) {
      err = "MJP_REDUCE_NCCL: libnccl lacks a required symbol";
      return false;
    }
    return true;
  }
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (stable across NCCL 2.x)

thread_local std::string g_multi_create_error;

}  // namespace

struct mjp_multi {
  std::vector<mjp_handle *> h;
  std::vector<int> dev;
  int reduce = MJP_REDUCE_PEER;
  NcclApi nccl;
  std::vector<void *> comms;
  DevBuf d_stage, d_total;        // on device dev[0]: [G][len] gathered vectors, [len] their sum
  std::vector<DevBuf> d_red;      // NCCL: per-device receive buffer
  cudaEvent_t r0 = nullptr, r1 = nullptr;
  float last_ms = 0.f, reduce_ms = 0.f;
  std::string err;
};

extern "C" {

int mjp_multi_create(const mjp_multi_config *cfg, mjp_multi **out) {
  if (!out) return MJP_ERR_INVALID;
  *out = nullptr;
  int visible = mjp_device_count();
  const int G = (cfg && cfg->n_devices > 0) ? cfg->n_devices : visible;
  if (G < 1) { g_multi_create_error = "mjp_multi_create: no CUDA device"; return MJP_ERR_CUDA; }
  mjp_multi *m = new (std::nothrow) mjp_multi();
  if (!m) return MJP_ERR_NOMEM;
  m->reduce = cfg ? cfg->reduce : MJP_REDUCE_PEER;
  for (int g = 0; g < G; ++g) {
    const int d = (cfg && cfg->devices) ? cfg->devices[g] : g;
    mjp_config c{d, cfg ? cfg->flags : 0, 0};
    mjp_handle *h = nullptr;
    int rc = (d >= 0 && d < visible) ? mjp_create(&c, &h) : MJP_ERR_INVALID;
    if (rc != MJP_OK) {
      g_multi_create_error = rc == MJP_ERR_INVALID ? "mjp_multi_create: device ordinal out of range" : mjp_last_error(nullptr);
      mjp_multi_destroy(m);
      return rc;
    }
    m->h.push_back(h); m->dev.push_back(d);
  }
  cudaSetDevice(m->dev[0]);
  cudaEventCreate(&m->r0); cudaEventCreate(&m->r1);
  if (m->reduce == MJP_REDUCE_NCCL) {
    if (!m->nccl.load(g_multi_create_error)) { mjp_multi_destroy(m); return MJP_ERR_UNSUPPORTED; }
    m->comms.assign(G, nullptr);
    const int rc = m->nccl.CommInitAll(m->comms.data(), G, m->dev.data());
    if (rc != 0) {
      g_multi_create_error = std::string("ncclCommInitAll: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(rc) : "error");
      m->comms.clear();
      mjp_multi_destroy(m);
      return MJP_ERR_CUDA;
    }
    m->d_red.resize(G);
  } else if (m->reduce != MJP_REDUCE_PEER) {
    g_multi_create_error = "mjp_multi_create: unknown reduce kind";
    mjp_multi_destroy(m);
    return MJP_ERR_INVALID;
  } else {
    for (int g = 1; g < G; ++g) {                 // device 0 reads its peers directly when the link allows it
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->dev[0], m->dev[g]);
      if (can) { cudaSetDevice(m->dev[0]); cudaDeviceEnablePeerAccess(m->dev[g], 0); }
    }
    cudaGetLastError();                           // (already enabled is not an error worth keeping)
  }
  *out = m;
  return MJP_OK;
}

void mjp_multi_destroy(mjp_multi *m) {
  if (!m) return;
  for (size_t g = 0; g < m->comms.size(); ++g) if (m->comms[g]) m->nccl.CommDestroy(m->comms[g]);
  for (size_t g = 0; g < m->d_red.size(); ++g) { cudaSetDevice(m->dev[g]); m->d_red[g].release(); }
  if (!m->dev.empty()) {
    cudaSetDevice(m->dev[0]);
    m->d_stage.release(); m->d_total.release();
    if (m->r0) cudaEventDestroy(m->r0);
    if (m->r1) cudaEventDestroy(m->r1);
  }
  for (mjp_handle *h : m->h) mjp_destroy(h);
  delete m;
}

const char *mjp_multi_last_error(const mjp_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int mjp_multi_device_count(const mjp_multi *m) { return m ? (int)m->h.size() : 0; }

int mjp_multi_run(mjp_multi *m, const mjp_line_desc *line, uint64_t n_inst, const mjp_rng_spec *rng, mjp_stats_out *stats) {
  if (!m || !line || !rng) return MJP_ERR_INVALID;
  if (n_inst == 0) { m->err = "n_inst must be >= 1"; return MJP_ERR_INVALID; }
  const int G = (int)m->h.size();
  const uint64_t per = (n_inst + G - 1) / G;           // ceil(n/G) contiguous instance ids per device
  std::vector<int> rc(G, MJP_OK);
  std::vector<uint64_t> lo(G), cnt(G);
  for (int g = 0; g < G; ++g) {
    lo[g] = std::min<uint64_t>((uint64_t)g * per, n_inst);
    cnt[g] = std::min<uint64_t>(lo[g] + per, n_inst) - lo[g];
  }
  // per device: upload its window, run, copy its columns of the per-instance planes back (no aggregate yet)
  mjp_stats_out planes{};
  if (stats) { planes = *stats; planes.aggregate = nullptr; }
  auto work = [&](int g) {
    if (!cnt[g]) return;
    mjp_handle *h = m->h[g];
    int r = upload_impl(h, line, cnt[g], lo[g], n_inst);
    mjp_rng_spec rg = *rng;
    rg.first_instance_id = rng->first_instance_id + lo[g];
    if (r == MJP_OK) r = mjp_launch(h, &rg, 0);
    if (r == MJP_OK) r = mjp_sync(h);
    if (r == MJP_OK && stats) r = download_impl(h, &planes, nullptr, lo[g], n_inst);
    rc[g] = r;
  };
  std::vector<std::thread> th;
  for (int g = 1; g < G; ++g) th.emplace_back(work, g);
  work(0);
  for (auto &t : th) t.join();
  m->last_ms = 0.f;
  for (int g = 0; g < G; ++g) {
    if (rc[g] != MJP_OK) { m->err = std::string("device ") + std::to_string(m->dev[g]) + ": " + mjp_last_error(m->h[g]); return rc[g]; }
    if (cnt[g]) m->last_ms = std::max(m->last_ms, m->h[g]->last_ms);
  }
  // ---- the one collective: sum of the aggregate vectors ----
  const int len = mjp_aggregate_len(line->M);
  const size_t bytes = (size_t)len * 8;
  int used = 0;
  for (int g = 0; g < G; ++g) used += cnt[g] != 0;
  mjp_handle *h0 = m->h[0];
  auto cuda_fail = [&](cudaError_t e, const char *what) { m->err = std::string(what) + ": " + cudaGetErrorString(e); return MJP_ERR_CUDA; };
  cudaError_t e;
  if ((e = cudaSetDevice(m->dev[0])) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  const double *result = nullptr;
  if (m->reduce == MJP_REDUCE_NCCL) {
    for (int g = 0; g < G; ++g) {
      cudaSetDevice(m->dev[g]);
      if ((e = m->d_red[g].reserve(bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
      if (!cnt[g]) {                                   // an empty shard contributes zeros
        if ((e = m->h[g]->d_agg.reserve(bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        cudaMemsetAsync(m->h[g]->d_agg.ptr, 0, bytes, m->h[g]->stream);
      }
    }
    cudaSetDevice(m->dev[0]);
    cudaEventRecord(m->r0, h0->stream);
    m->nccl.GroupStart();
    int nrc = 0;
    for (int g = 0; g < G; ++g) {
      cudaSetDevice(m->dev[g]);
      const int r = m->nccl.AllReduce(m->h[g]->d_agg.ptr, m->d_red[g].ptr, (size_t)len, kNcclFloat64, kNcclSum, m->comms[g],
                                      m->h[g]->stream);
      if (r) nrc = r;
    }
    const int r2 = m->nccl.GroupEnd();
    if (nrc || r2) { m->err = std::string("ncclAllReduce: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(nrc ? nrc : r2) : "error"); return MJP_ERR_CUDA; }
    cudaSetDevice(m->dev[0]);
    cudaEventRecord(m->r1, h0->stream);
    for (int g = 0; g < G; ++g) { cudaSetDevice(m->dev[g]); if ((e = cudaStreamSynchronize(m->h[g]->stream)) != cudaSuccess) return cuda_fail(e, "ncclAllReduce"); }
    cudaSetDevice(m->dev[0]);
    result = m->d_red[0].as<double>();
  } else {
    if ((e = m->d_stage.reserve(bytes * G)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = m->d_total.reserve(bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaEventRecord(m->r0, h0->stream);
    int rows = 0;
    for (int g = 0; g < G; ++g) {
      if (!cnt[g]) continue;
      e = cudaMemcpyPeerAsync((char *)m->d_stage.ptr + bytes * rows, m->dev[0], m->h[g]->d_agg.ptr, m->dev[g], bytes, h0->stream);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyPeerAsync");
      ++rows;
    }
    mjp::mjp_aggregate_final<<<(len + 127) / 128, 128, 0, h0->stream>>>(m->d_stage.as<double>(), rows, len, m->d_total.as<double>());
    cudaEventRecord(m->r1, h0->stream);
    if ((e = cudaStreamSynchronize(h0->stream)) != cudaSuccess) return cuda_fail(e, "aggregate sum");
    result = m->d_total.as<double>();
  }
  cudaEventElapsedTime(&m->reduce_ms, m->r0, m->r1);
  if (stats && stats->aggregate) {
    if ((e = cudaMemcpy(stats->aggregate, result, bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  }
  (void)used;
  return MJP_OK;
}

int mjp_multi_last_ms(mjp_multi *m, float *kernel_ms_max, float *reduce_ms) {
  if (!m) return MJP_ERR_INVALID;
  if (kernel_ms_max) *kernel_ms_max = m->last_ms;
  if (reduce_ms) *reduce_ms = m->reduce_ms;
  return MJP_OK;
}

}  // extern "C"
