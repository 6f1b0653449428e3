"""GPU parity tests proper: the CUDA path, called through the C ABI
(libmjpdes_b200.so via mjpdes_b200.engine.Handle), against the CPU oracle on the same
seeded inputs -- event sequence (hash over every executed action) and every accumulator
bit-for-bit -- plus the reference's own golden vectors and size-independent properties at
BASELINE.json's full sizes.  Tolerance stated by the north star: 1e-9 relative on f64
statistics in RNG-replay mode (tests/parity.py); in practice the results are bit-identical."""
import numpy as np
import pytest

from lines import (GOLDENS, SEED, example_line, golden_two_machine, long50_line, mixed20_line, random_line,
                   submodel1_line, sweep10_line)
from mjpdes_b200 import abi
from parity import assert_same

pytestmark = pytest.mark.gpu
MODES = [abi.RNG_COUNTER, abi.RNG_REPLAY]
KIND = {":up": 0, ":down": 1}


@pytest.mark.parametrize("gid", ["G1", "G2"])
def test_end_time_known_answers_on_device(engine, gid):
    """core_test.clj job-tests-1/2 through the device end-time functions."""
    g = GOLDENS[gid]
    got = engine.probe_end_time([KIND[k] for _, k, _ in g["events"]], [t for _, _, t in g["events"]],
                                g["clock"], g["dur"])
    assert abs(got - g["expected"]) < g["tol"]
    assert got == g["expected"]


@pytest.mark.parametrize("mode", MODES)
def test_c1_example_single_run_replay(engine, oracle, mode):
    """C1: resources/example.clj, one instance, full length: identical event-sequence hash and stats."""
    line = example_line()
    got = engine.run(line, 1, mode=mode, seed=SEED)
    want = oracle.run(line, 1, mode=mode, seed=SEED)
    assert_same(got.stats, want.stats)
    assert 2.1e5 < got.total_events < 2.3e5


@pytest.mark.parametrize("mode", MODES)
def test_example_many_instances_bit_exact(engine, oracle, mode):
    line = example_line(run_to=5000.0, warm_up=500.0)
    n = 777                                  # ragged: not a multiple of the block size
    got = engine.run(line, n, mode=mode, seed=SEED, first_instance_id=12345)
    want = oracle.run(line, n, mode=mode, seed=SEED, first_instance_id=12345, n_threads=8)
    assert_same(got.stats, want.stats)
    assert np.allclose(got.stats.aggregate, want.stats.aggregate, rtol=1e-12)


@pytest.mark.parametrize("gid", ["G7", "G8", "G9"])
def test_reference_two_machine_models(engine, oracle, gid):
    """The whole-engine goldens (load-m1, reliable BAS / BBS cycles) are RNG-free: strongest parity test."""
    line = golden_two_machine(gid, lam=0.0, run_to=1000.0, log=False)
    got = engine.run(line, 3, mode=abi.RNG_REPLAY, seed=SEED)
    want = oracle.run(line, 3, mode=abi.RNG_REPLAY, seed=SEED)
    assert_same(got.stats, want.stats)


@pytest.mark.parametrize("trial", range(12))
def test_random_lines(engine, oracle, trial):
    rng = np.random.default_rng(2000 + trial)
    M, K = int(rng.integers(2, 13)), int(rng.integers(1, 5))
    line = random_line(rng, M, K, run_to=400.0, warm_up=40.0, jm2dm=bool(rng.random() < 0.3),
                       up_down=bool(rng.random() < 0.4))
    mode = MODES[trial % 2]
    got = engine.run(line, 65, mode=mode, seed=trial)
    want = oracle.run(line, 65, mode=mode, seed=trial, n_threads=8)
    assert_same(got.stats, want.stats)


@pytest.mark.parametrize("M", [20, 33, 50, 64])
def test_wide_lines(engine, oracle, M):
    line = random_line(np.random.default_rng(M), M, 3, run_to=150.0, warm_up=10.0)
    got = engine.run(line, 40, seed=5)
    want = oracle.run(line, 40, seed=5, n_threads=8)
    assert_same(got.stats, want.stats)


def test_identical_machines_many_time_ties(engine, oracle):
    line = abi.HostLine(lam=[.1] * 6, mu=[.9] * 6, W=[1.] * 6, discipline=[0, 1, 0, 1, 0, 0], N=[2] * 5,
                        portion=[1.], w=[[1.] * 6], warm_up=50, run_to=1500)
    for mode in MODES:
        assert_same(engine.run(line, 130, mode=mode, seed=3).stats, oracle.run(line, 130, mode=mode, seed=3, n_threads=8).stats)


def test_c3_sweep_planes(engine, oracle):
    """C3: per-instance W / N override planes (10-machine grid), a slice of the 1M-point sweep."""
    line = sweep10_line(200, warm_up=100, run_to=1000, first=654321)
    assert_same(engine.run(line, 200, seed=SEED).stats, oracle.run(line, 200, seed=SEED, n_threads=8).stats)


def test_c4_c5_shapes(engine, oracle):
    for line, n in ((mixed20_line(warm_up=50, run_to=500, log=False), 48), (long50_line(warm_up=50, run_to=400), 40)):
        assert_same(engine.run(line, n, seed=SEED).stats, oracle.run(line, n, seed=SEED, n_threads=8).stats)


def test_status_codes(engine, oracle):
    line = abi.HostLine(lam=[.1, .1], mu=[.9, .9], W=[1., 1.], discipline=[0, 0], N=[2], portion=[0.1, 0.2, 0.7],
                        w=np.ones((3, 2)), warm_up=0, run_to=500)
    got, want = engine.run(line, 33, seed=1).stats, oracle.run(line, 33, seed=1).stats
    assert set(want.status.tolist()) == {abi.INST_JOBTYPE_NIL}
    assert np.array_equal(got.status, want.status)
    # an instance that dies later leaves the same event count and accumulators behind (the tick before the fatal
    # add-job is never pushed)
    line = abi.HostLine(lam=[.1, .1, .1], mu=[.9, .9, .9], W=[1., 1., 1.], discipline=[0, 1, 0], N=[2, 1],
                        portion=[0.45, 0.55], w=np.ones((2, 3)), warm_up=0, run_to=500)
    got, want = engine.run(line, 70, seed=1).stats, oracle.run(line, 70, seed=1).stats
    assert set(want.status.tolist()) == {abi.INST_JOBTYPE_NIL} and want.n_events.max() > 50
    assert_same(got, want)


def test_invalid_descriptors_are_rejected(engine):
    bad = example_line()
    bad.mu[2] = 0.0
    with pytest.raises(abi.MjpError) as e:
        engine.run(bad, 4)
    assert e.value.code == abi.ERR_INVALID
    bad = example_line()
    bad.N[1] = 0
    with pytest.raises(abi.MjpError):
        engine.run(bad, 4)
    with pytest.raises(abi.MjpError):
        engine.run(example_line(), 0)


def test_results_do_not_depend_on_batching(engine):
    """Instance id -> RNG substream: [0,n) in one call == two calls over [0,n/2) and [n/2,n) (multi-GPU sharding)."""
    line = example_line(run_to=3000.0, warm_up=300.0)
    n = 512
    whole = engine.run(line, n, seed=SEED).stats
    a = engine.run(line, n // 2, seed=SEED, first_instance_id=0).stats
    b = engine.run(line, n // 2, seed=SEED, first_instance_id=n // 2).stats
    for f in ("n_events", "njobs", "event_hash", "residence_sum"):
        assert np.array_equal(getattr(whole, f), np.concatenate([getattr(a, f), getattr(b, f)]))
    assert np.array_equal(whole.blocked_time, np.concatenate([a.blocked_time, b.blocked_time], axis=1))
    again = engine.run(line, n, seed=SEED).stats
    assert np.array_equal(whole.event_hash, again.event_hash) and np.array_equal(whole.occ_time, again.occ_time)


def _welch_ok(x, y, z=3.3):
    """two-sample test at well beyond 99 %: |mean difference| <= z * standard error."""
    se = np.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y))
    return abs(x.mean() - y.mean()) <= z * se + 1e-12


def test_c2_counter_mode_statistics_agree_with_replay_oracle(engine, oracle):
    """C2 (reduced horizon so the oracle sample finishes in seconds): metrics over many COUNTER-mode GPU
    replications vs a seed-disjoint oracle sample drawn with the reference's literal accumulate-until
    loop (REPLAY): means agree within a 99 % two-sample confidence interval, metric by metric."""
    line = example_line(run_to=6000.0, warm_up=1000.0)
    run_time = line.run_to - line.warm_up
    g = engine.run(line, 6000, mode=abi.RNG_COUNTER, seed=SEED).stats
    o = oracle.run(line, 1500, mode=abi.RNG_REPLAY, seed=SEED + 1, first_instance_id=10 ** 7, n_threads=8).stats
    assert np.all(g.status == 0) and np.all(o.status == 0)
    lv = line.level_offsets
    checks = [("TP", g.njobs / run_time, o.njobs / run_time),
              ("residence", g.residence_sum / g.njobs, o.residence_sum / o.njobs)]
    for i in range(line.M):
        checks.append((f"blocked{i}", g.blocked_time[i] / run_time, o.blocked_time[i] / run_time))
        checks.append((f"starved{i}", g.starved_time[i] / run_time, o.starved_time[i] / run_time))
    for b in range(line.M - 1):
        lev = np.arange(line.N[b] + 1)[:, None]
        checks.append((f"occ{b}", (lev * g.occ_time[lv[b]:lv[b + 1]]).sum(0) / run_time,
                       (lev * o.occ_time[lv[b]:lv[b + 1]]).sum(0) / run_time))
    failed = [name for name, x, y in checks if not _welch_ok(x, y)]
    assert not failed, failed


def test_c2_full_size_properties(engine):
    """C2 at BASELINE.json's size: example.clj x 100 000 replications, full horizon, on one B200.
    Size-independent properties: every instance ends normally past the horizon; per-instance time
    balance of machine 1 (working + blocked <= run time); occupancy times of a buffer never exceed the
    run time; job conservation; the aggregate block equals the per-instance sums."""
    line = example_line()
    n = 100_000
    r = engine.run(line, n, mode=abi.RNG_COUNTER, seed=SEED)
    s = r.stats
    run_time = line.run_to - line.warm_up
    assert np.all(s.status == abi.INST_NORMAL_END)
    assert np.all(s.final_clock > line.run_to)
    assert abs(s.n_events.mean() - 2.20e5) < 0.005 * 2.20e5
    lv = line.level_offsets
    for b in range(line.M - 1):
        tot = s.occ_time[lv[b]:lv[b + 1]].sum(0)
        assert np.all(tot <= run_time + 1e-6) and np.all(tot > 0.95 * run_time)
    assert np.all(s.blocked_time >= 0) and np.all(s.starved_time >= 0)
    assert np.all(s.blocked_time[-1] == 0) and np.all(s.starved_time[0] == 0)
    assert np.all(s.current_job >= s.njobs) and np.all(s.current_job - s.njobs < 20000)
    tp = s.njobs / run_time
    assert 0.725 < tp.mean() < 0.745                      # recorded reference outputs: 0.7301, 0.7426
    agg = s.aggregate
    assert agg[0] == n
    assert np.isclose(agg[1], tp.sum(), rtol=1e-12) and np.isclose(agg[2], (tp * tp).sum(), rtol=1e-12)
    assert np.isclose(agg[5], (s.blocked_time[0] / run_time).sum(), rtol=1e-12)
    assert len(np.unique(s.event_hash)) == n              # every replication is a different realisation


@pytest.mark.parametrize("M,K,n,run_to", [(50, 16, 12001, 100.0), (10, 1, 50021, 100.0)])   # last warp partly idle
def test_cold_layout_variants_on_device(engine, oracle, M, K, n, run_to):
    """Batches of wide lines that do not fit one wave with all state in shared memory switch to the COLD
    variants (cold per-instance state in an L2-resident scratch, mjp_abi.cu): same results bit for bit."""
    if M == 50:
        line = long50_line(warm_up=10.0, run_to=run_to)
    else:
        line = sweep10_line(n, warm_up=10.0, run_to=run_to, first=7)
    got = engine.run(line, n, seed=SEED)
    want = oracle.run(line, n, seed=SEED, n_threads=8)
    assert_same(got.stats, want.stats)


def test_run_to_job_and_all_override_planes(engine, oracle):
    line = example_line(run_to=300.0, warm_up=30.0, run_to_job=600)
    assert_same(engine.run(line, 70, seed=11).stats, oracle.run(line, 70, seed=11, n_threads=8).stats)
    rng = np.random.default_rng(42)
    n, M, K = 90, 4, 2
    line = abi.HostLine(lam=[.1] * M, mu=[.9] * M, W=[1.] * M, discipline=[0, 1, 0, 0], N=[4, 4, 4], portion=[.5, .5],
                        w=np.ones((K, M)), warm_up=20, run_to=400,
                        lam_inst=rng.uniform(0.05, 0.2, (M, n)), mu_inst=rng.uniform(0.5, 1.5, (M, n)),
                        W_inst=rng.uniform(0.8, 1.2, (M, n)), N_inst=rng.integers(1, 5, (M - 1, n)),
                        w_inst=rng.uniform(0.5, 1.5, (K * M, n)))
    assert_same(engine.run(line, n, seed=2).stats, oracle.run(line, n, seed=2, n_threads=8).stats)


@pytest.mark.parametrize("mode", MODES)
def test_time_slicing_on_device(engine, oracle, mode):
    """mjp_launch_slice: park at slice boundaries, resume, and end with the results of one uninterrupted run."""
    for line, n in ((example_line(run_to=900.0, warm_up=90.0), 300),
                    (random_line(np.random.default_rng(10), 40, 2, run_to=200.0, warm_up=20.0), 70)):
        engine.upload(line, n)
        resume = False
        for until in (0.0, 37.5, 37.5, 120.0, 333.3, 899.0, 5000.0, float("inf")):
            engine.launch_slice(until, resume, mode=mode, seed=21, first_instance_id=5)
            engine.sync()
            resume = True
            st = engine.download_fields(("status",)).status
            if until < 100.0:
                assert np.all(st == abi.INST_PAUSED)
        got = engine.download().stats
        want = oracle.run(line, n, mode=mode, seed=21, first_instance_id=5, n_threads=8).stats
        assert_same(got, want)
