// asan_main.cpp -- TEST-ONLY: the kernel source under AddressSanitizer + UBSan on the host
// (compute-sanitizer is closed on the GPU pool).  Exact-size heap buffers for every plane, and the
// unused tail of the shared-memory arena is poisoned so a wrong slot offset traps.
#include <sanitizer/asan_interface.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hostsim.cpp"

static int run_case(int M, int K, int n, double run_to, int mode, bool log, bool cold, bool sweep, bool bbs) {
  std::vector<double> lam(M, 0.1), mu(M, 0.9), W(M, 1.0), portion(K, 1.0 / K), w((size_t)K * M);
  std::vector<uint8_t> disc(M, 0);
  std::vector<int32_t> N(M - 1);
  for (int i = 0; i < M; ++i) { W[i] = 1.0 + 0.05 * (i % 4); disc[i] = bbs && (i % 2); }
  for (int b = 0; b < M - 1; ++b) N[b] = 1 + b % 4;
  for (int q = 0; q < K * M; ++q) w[q] = 0.75 + 0.25 * (q % 3);
  std::vector<double> Wi; std::vector<int32_t> Ni;
  mjp_line_desc L{};
  L.M = M; L.K = K; L.lambda = lam.data(); L.mu = mu.data(); L.W = W.data(); L.discipline = disc.data();
  L.N = N.data(); L.portion = portion.data(); L.w = w.data(); L.warm_up = run_to / 10; L.run_to = run_to;
  L.run_to_job = -1; L.log = log; L.up_down = log; L.max_lines = ~0ull;
  if (sweep) {
    Wi.resize((size_t)M * n); Ni.resize((size_t)(M - 1) * n);
    for (size_t q = 0; q < Wi.size(); ++q) Wi[q] = 0.8 + 0.4 * ((q * 7) % 11) / 11.0;
    for (int b = 0; b < M - 1; ++b) for (int g = 0; g < n; ++g) Ni[(size_t)b * n + g] = 1 + (b + g) % N[b];
    L.W_inst = Wi.data(); L.N_inst = Ni.data();
  }
  int levels = 0; for (int b = 0; b < M - 1; ++b) levels += N[b] + 1;
  std::vector<int64_t> njobs(n), cur(n); std::vector<double> res(n), blk((size_t)M * n), stv((size_t)M * n),
      occ((size_t)levels * n), clk(n), agg(1 + 2 * (3 * M + 1));
  std::vector<uint64_t> nev(n), hash(n); std::vector<uint8_t> status(n, 255);
  mjp_stats_out s{njobs.data(), res.data(), blk.data(), stv.data(), occ.data(), clk.data(), cur.data(), nev.data(),
                  hash.data(), status.data(), agg.data()};
  std::vector<mjp_log_record> recs(log ? 400000 : 1);
  mjp_log_out lo{recs.data(), recs.size(), 0, 0, nullptr};
  mjp_rng_spec rng{mode, 0, 0x4D4A50646573ull, 0};
  // poison what the chosen layout does not use of the shared-memory arena
  {
    mjp::LinePlan pl = mjp::build_plan(&L, n);
    mjp::KParams kp = pl.kp;
    const bool mt = !cold && !log && M <= 8;
    const size_t per_thread = mjp::plan_layout(kp, mt, mode != MJP_RNG_COUNTER, log, cold);   // (sets kp.const_bytes)
    const size_t used = kp.const_bytes + per_thread;
    ASAN_UNPOISON_MEMORY_REGION(mjp::smem_raw, sizeof(mjp::smem_raw));
    ASAN_POISON_MEMORY_REGION(mjp::smem_raw + used, sizeof(mjp::smem_raw) - used);
  }
  int rc = hostsim_run(&L, n, &rng, &s, 1, log ? &lo : nullptr, cold);
  unsigned long long ev = 0; int bad = 0;
  for (int g = 0; g < n; ++g) { ev += nev[g]; bad += status[g] != 0; }
  std::printf("M=%d K=%d mode=%d log=%d cold=%d sweep=%d: rc=%d events=%llu bad=%d records=%llu\n", M, K, mode, log, cold,
              sweep, rc, ev, bad, (unsigned long long)lo.count);
  return rc != 0 || bad != 0;
}

int main() {
  int fail = 0;
  for (int mode = 0; mode < 2; ++mode) {
    fail |= run_case(5, 2, 4, 400, mode, false, false, false, false);
    fail |= run_case(5, 2, 3, 300, mode, true, false, false, true);
    fail |= run_case(2, 1, 3, 300, mode, false, false, false, false);
    fail |= run_case(8, 3, 3, 200, mode, false, false, true, true);
    fail |= run_case(10, 1, 3, 200, mode, false, true, true, false);
    fail |= run_case(20, 8, 2, 150, mode, true, true, false, true);
    fail |= run_case(50, 16, 2, 80, mode, false, true, false, false);
    fail |= run_case(64, 4, 2, 60, mode, true, false, false, true);
  }
  std::printf(fail ? "ASAN RUN FAILED\n" : "asan run clean\n");
  return fail;
}
