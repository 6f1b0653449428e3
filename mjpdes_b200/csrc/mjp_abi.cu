// mjp_abi.cu -- host side of the C ABI declared in include/mjpdes_b200.h.
//
// The handle owns every device buffer, a stream and two timing events; the
// caller owns every host buffer.  No call throws or aborts: errors come back
// as MJP_ERR_* with text in mjp_last_error().
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/mjpdes_b200.h"
#include "mjp_line_kernel.cuh"
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
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  // uploaded line
  bool uploaded = false, launched = false, sweep = false;
  mjp::KParams kp{};
  // device buffers
  DevBuf d_const, d_planes, d_out, d_scratch, d_partials, d_agg, d_log, d_logcur, d_cold, d_park, d_tapes, d_stage;
  // out-plane offsets inside d_out (bytes)
  size_t o_njobs = 0, o_res = 0, o_blk = 0, o_stv = 0, o_occ = 0, o_clk = 0, o_cur = 0, o_nev = 0, o_hash = 0,
         o_status = 0, out_bytes = 0, zero_bytes = 0;
  int agg_blocks = 0;
  bool want_log = false;
  float last_ms = 0.f;
  int last_launches = 0;
  int block_threads = 0;
  size_t smem_bytes = 0;
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
static int validate(mjp_handle *h, const mjp_line_desc *L, uint64_t n_inst) {
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
    if (!(L->lambda[i] >= 0.0)) return fail(h, MJP_ERR_INVALID, ":lambda must be non-negative (core.clj:89)");
    if (!(L->mu[i] > 0.0)) return fail(h, MJP_ERR_INVALID, ":mu must be positive (core.clj:88)");
    if (!(L->W[i] > 0.0)) return fail(h, MJP_ERR_INVALID, ":W must be positive (core.clj:87)");
    if (L->discipline[i] > 1) return fail(h, MJP_ERR_INVALID, ":discipline must be :BAS or :BBS (core.clj:91)");
  }
  for (int b = 0; b < L->M - 1; ++b) {
    if (L->N[b] < 1) return fail(h, MJP_ERR_INVALID, ":N must be a positive integer (core.clj:95)");
    if (L->N[b] > 254) return fail(h, MJP_ERR_UNSUPPORTED, "buffer capacity above 254");
  }
  for (int k = 0; k < L->K; ++k)
    if (!(L->portion[k] > 0.0)) return fail(h, MJP_ERR_INVALID, ":portion must be positive (core.clj:109)");
  for (int q = 0; q < L->K * L->M; ++q)
    if (!(L->w[q] > 0.0)) return fail(h, MJP_ERR_INVALID, ":w values must be positive (core.clj:108)");
  if (!(L->warm_up >= 0.0)) return fail(h, MJP_ERR_INVALID, ":warm-up-time must be non-negative (core.clj:104)");
  if (!(L->run_to > 0.0)) return fail(h, MJP_ERR_INVALID, ":run-to-time must be positive (core.clj:105)");
  if (L->N_inst)
    for (int b = 0; b < L->M - 1; ++b)
      for (uint64_t g = 0; g < n_inst; ++g) {
        const int v = L->N_inst[(size_t)b * n_inst + g];
        if (v < 1 || v > L->N[b]) return fail(h, MJP_ERR_INVALID, "N_inst value outside 1..N[b]");
      }
  const uint64_t nm = n_inst * (uint64_t)L->M, nkm = n_inst * (uint64_t)L->M * (uint64_t)L->K;
  if (L->lambda_inst) for (uint64_t q = 0; q < nm; ++q) if (!(L->lambda_inst[q] >= 0.0)) return fail(h, MJP_ERR_INVALID, "lambda_inst must be non-negative");
  if (L->mu_inst) for (uint64_t q = 0; q < nm; ++q) if (!(L->mu_inst[q] > 0.0)) return fail(h, MJP_ERR_INVALID, "mu_inst must be positive");
  if (L->W_inst) for (uint64_t q = 0; q < nm; ++q) if (!(L->W_inst[q] > 0.0)) return fail(h, MJP_ERR_INVALID, "W_inst must be positive");
  if (L->w_inst) for (uint64_t q = 0; q < nkm; ++q) if (!(L->w_inst[q] > 0.0)) return fail(h, MJP_ERR_INVALID, "w_inst must be positive");
  return MJP_OK;
}

template <typename MaskT, int MT, bool REPLAY, bool HASH, bool SLICE, bool LOG, bool COLD>
static cudaError_t launch_variant(const mjp::KParams &kp, int blocks, int threads, size_t smem, cudaStream_t st) {
  auto k = mjp::mjp_line_kernel<MaskT, MT, REPLAY, HASH, SLICE, LOG, COLD>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<blocks, threads, smem, st>>>(kp);
  return cudaGetLastError();
}

template <typename MaskT, int MT, bool COLD = false>
static cudaError_t launch_flags(bool replay, bool hash, bool slice, bool log, const mjp::KParams &kp, int blocks,
                                int threads, size_t smem, cudaStream_t st) {
  const int v = (replay ? 4 : 0) | (hash ? 2 : 0) | (slice ? 1 : 0);
#define MJP_CASE(n, LG) \
  case n: return launch_variant<MaskT, MT, ((n) & 4) != 0, ((n) & 2) != 0, ((n) & 1) != 0, LG, COLD>(kp, blocks, threads, smem, st);
  if (log) {
    switch (v) {
      MJP_CASE(0, true) MJP_CASE(1, true) MJP_CASE(2, true) MJP_CASE(3, true)
      MJP_CASE(4, true) MJP_CASE(5, true) MJP_CASE(6, true) default: MJP_CASE(7, true)
    }
  }
  switch (v) {
    MJP_CASE(0, false) MJP_CASE(1, false) MJP_CASE(2, false) MJP_CASE(3, false)
    MJP_CASE(4, false) MJP_CASE(5, false) MJP_CASE(6, false) default: MJP_CASE(7, false)
  }
#undef MJP_CASE
}

static bool use_register_ends(const mjp::KParams &kp, bool log, int threads) {
  static const bool force_generic = std::getenv("MJP_FORCE_GENERIC") != nullptr;   // A/B measurements only
  (void)log;
  return !force_generic && threads == MJP_FIXED_NT && kp.M >= 2 && kp.M <= 8;
}

static cudaError_t launch_line_kernel(bool replay, bool hash, bool slice, bool log, bool cold, const mjp::KParams &kp,
                                      int blocks, int threads, size_t smem, cudaStream_t st) {
#define MJP_ARGS replay, hash, slice, log, kp, blocks, threads, smem, st
  if (cold) return kp.M > 32 ? launch_flags<uint64_t, 0, true>(MJP_ARGS) : launch_flags<uint32_t, 0, true>(MJP_ARGS);
  if (kp.M > 32) return launch_flags<uint64_t, 0>(MJP_ARGS);
  if (use_register_ends(kp, log, threads)) {
    switch (kp.M) {
      case 2: return launch_flags<uint32_t, 2>(MJP_ARGS);
      case 3: return launch_flags<uint32_t, 3>(MJP_ARGS);
      case 4: return launch_flags<uint32_t, 4>(MJP_ARGS);
      case 5: return launch_flags<uint32_t, 5>(MJP_ARGS);
      case 6: return launch_flags<uint32_t, 6>(MJP_ARGS);
      case 7: return launch_flags<uint32_t, 7>(MJP_ARGS);
      case 8: return launch_flags<uint32_t, 8>(MJP_ARGS);
      default: break;
    }
  }
  return launch_flags<uint32_t, 0>(MJP_ARGS);
#undef MJP_ARGS
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
  *out = h;
  return MJP_OK;
}

void mjp_destroy(mjp_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (DevBuf *b : {&h->d_const, &h->d_planes, &h->d_out, &h->d_scratch, &h->d_partials, &h->d_agg, &h->d_log,
                    &h->d_logcur, &h->d_cold, &h->d_park, &h->d_tapes, &h->d_stage})
    b->release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *mjp_last_error(const mjp_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mjp_upload(mjp_handle *h, const mjp_line_desc *L, uint64_t n_inst) {
  if (!h) return MJP_ERR_INVALID;
  h->uploaded = false; h->launched = false;
  int rc = validate(h, L, n_inst);
  if (rc != MJP_OK) return rc;
  CK(cudaSetDevice(h->device));
  const int M = L->M, K = L->K;
  mjp::LinePlan pl = mjp::build_plan(L, n_inst);
  mjp::KParams &kp = h->kp;
  kp = pl.kp;
  const int R = kp.R, levels = kp.L;

  // ---- constants: lam mu W w dur thr | N lvl ----
  const size_t dbytes = pl.cd.size() * sizeof(double), ibytes = pl.ci.size() * sizeof(int32_t);
  CK(h->d_const.reserve(dbytes + ibytes));
  CK(cudaMemcpyAsync(h->d_const.ptr, pl.cd.data(), dbytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync((char *)h->d_const.ptr + dbytes, pl.ci.data(), ibytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // pl is stack-owned
  const double *dc = h->d_const.as<double>();
  kp.lam = dc + pl.off_lam(); kp.mu = dc + pl.off_mu(); kp.W = dc + pl.off_W(); kp.w = dc + pl.off_w();
  kp.dur = dc + pl.off_dur(); kp.thr = dc + pl.off_thr();
  kp.N = reinterpret_cast<const int32_t *>((const char *)h->d_const.ptr + dbytes);
  kp.lvl = kp.N + (M - 1);

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
    auto put = [&](const void *src, size_t bytes, size_t padded) -> const void * {
      if (!src) return nullptr;
      void *dst = base + off;
      cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream);
      off += padded;
      return dst;
    };
    kp.lam_i = (const double *)put(L->lambda_inst, n_inst * M * sizeof(double), pM);
    kp.mu_i = (const double *)put(L->mu_inst, n_inst * M * sizeof(double), pM);
    kp.W_i = (const double *)put(L->W_inst, n_inst * M * sizeof(double), pM);
    kp.w_i = (const double *)put(L->w_inst, n_inst * (size_t)K * M * sizeof(double), pKM);
    kp.N_i = (const int32_t *)put(L->N_inst, n_inst * (size_t)(M - 1) * sizeof(int32_t), pN);
    CK(cudaGetLastError());
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
    auto cta_bytes = [&](bool c, int t) {
      mjp::KParams tmp = kp;
      const bool mt_ = !c && use_register_ends(kp, h->want_log, t);
      return (size_t)kp.const_bytes + mjp::plan_layout(tmp, mt_, replay, h->want_log, c) * t;
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
    const bool mt = !cold && use_register_ends(kp, h->want_log, threads);
    const size_t per_thread = mjp::plan_layout(kp, mt, replay, h->want_log, cold);
    if (cold) {
      const size_t fb = mjp::plan_align_up((size_t)mjp::cold_f64_slots(kp) * kp.n_inst * 8, 256);
      CK(h->d_cold.reserve(fb + (size_t)mjp::cold_u32_slots(kp, replay) * kp.n_inst * 4));
      kp.cold_f = h->d_cold.as<double>();
      kp.cold_u = reinterpret_cast<uint32_t *>((char *)h->d_cold.ptr + fb);
    }
    if (h->want_log) {                              // one tick of staged messages per instance; persists across slices
      CK(h->d_stage.reserve((size_t)kp.stage_cap * kp.n_inst * sizeof(unsigned long long)));
      kp.stage = h->d_stage.as<unsigned long long>();
    }
    h->smem_bytes = (size_t)kp.const_bytes + per_thread * threads;
    if (resume || until < std::numeric_limits<double>::infinity()) {   // == slice
      kp.park_words = 72 + 2 * kp.nf64 + kp.nu32;     // loop-carried registers (<= 72 words) + every shared slot
      CK(h->d_park.reserve((size_t)kp.park_words * kp.n_inst * 4));
      kp.park = h->d_park.as<uint32_t>();
    }
    const int blocks = (int)((kp.n_inst + threads - 1) / threads);
    const bool slice = resume || until < std::numeric_limits<double>::infinity();
    mjp::KParams kl = kp;                         // tapes are read by MJP_RNG_TAPE launches only
    if (rng->mode != MJP_RNG_TAPE) kl.tape_jt = kl.tape_u = kl.tape_e = nullptr;
    cudaError_t e = launch_line_kernel(replay, hash, slice, h->want_log, cold, kl, blocks, threads, h->smem_bytes, h->stream);
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
  if (h->launched) CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  return MJP_OK;
}

int mjp_download(mjp_handle *h, mjp_stats_out *s, mjp_log_out *log) {
  if (!h) return MJP_ERR_INVALID;
  if (!h->launched) return fail(h, MJP_ERR_INVALID, "mjp_download before mjp_launch");
  CK(cudaSetDevice(h->device));
  const mjp::KParams &kp = h->kp;
  const size_t n = kp.n_inst;
  const char *ob = (const char *)h->d_out.ptr;
  auto get = [&](void *dst, size_t o, size_t bytes) {
    if (dst) cudaMemcpyAsync(dst, ob + o, bytes, cudaMemcpyDeviceToHost, h->stream);
  };
  if (s) {
    get(s->njobs, h->o_njobs, n * 8); get(s->residence_sum, h->o_res, n * 8);
    get(s->blocked_time, h->o_blk, n * kp.M * 8); get(s->starved_time, h->o_stv, n * kp.M * 8);
    get(s->occ_time, h->o_occ, n * (size_t)kp.L * 8); get(s->final_clock, h->o_clk, n * 8);
    get(s->current_job, h->o_cur, n * 8); get(s->n_events, h->o_nev, n * 8);
    if (s->event_hash) {
      if (h->flags & MJP_CFG_EVENT_HASH) get(s->event_hash, h->o_hash, n * 8);
      else std::memset(s->event_hash, 0, n * 8);
    }
    get(s->status, h->o_status, n);
    if (s->aggregate)
      cudaMemcpyAsync(s->aggregate, h->d_agg.ptr, (size_t)mjp_aggregate_len(kp.M) * 8, cudaMemcpyDeviceToHost, h->stream);
  }
  int rc = MJP_OK;
  if (log) {
    log->count = 0; log->dropped = 0;
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
    }
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return rc;
}

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
