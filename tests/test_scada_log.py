"""SCADA record stream (what push-log hands to print-lines, util/log.clj:89-137): order after
clean-log-buf, :line numbering, :max-lines overshoot, :up&down? filtering, the :ends of :aj
records -- kernel source on the host (CPU) and the CUDA path through the C ABI (GPU) against
the oracle, record for record."""
import numpy as np
import pytest

from lines import GOLDENS, SEED, example_line, golden_two_machine, mixed20_line, pretty_act, random_line
from mjpdes_b200 import abi


def sorted_log(log):
    return np.sort(log, order=["instance", "line"])


def assert_same_log(got, want):
    a, b = sorted_log(got), sorted_log(want)
    assert len(a) == len(b)
    for f in a.dtype.names:
        if f == "aux":
            assert np.array_equal(a[f].view(np.uint64), b[f].view(np.uint64)), f
        else:
            assert np.array_equal(a[f], b[f]), f


CASES = [
    ("G7", lambda: golden_two_machine("G7", lam=0.0, run_to=200.0), abi.RNG_REPLAY),
    ("G8", lambda: golden_two_machine("G8", run_to=200.0), abi.RNG_REPLAY),
    ("G9", lambda: golden_two_machine("G9", run_to=200.0), abi.RNG_REPLAY),
    ("example-updown", lambda: example_line(run_to=1500, warm_up=100, log=True, up_down=True,
                                            max_lines=abi.UINT64_MAX), abi.RNG_COUNTER),
    ("example-max-lines", lambda: example_line(run_to=1500, warm_up=100, log=True, max_lines=500), abi.RNG_REPLAY),
    ("mixed20", lambda: mixed20_line(warm_up=20, run_to=300), abi.RNG_COUNTER),
]


@pytest.mark.parametrize("name,mk,mode", CASES, ids=[c[0] for c in CASES])
def test_kernel_source_log_matches_oracle(oracle, name, mk, mode):
    from hostsim import binding as H
    line = mk()
    want = oracle.run(line, 3, mode=mode, seed=SEED, log_capacity=500_000)
    got = H.run(line, 3, mode=mode, seed=SEED, log_capacity=500_000)
    assert_same_log(got.log, want.log)
    assert np.array_equal(got.event_hash, want.stats.event_hash)


@pytest.mark.parametrize("trial", range(8))
def test_kernel_source_log_random_lines(oracle, trial):
    from hostsim import binding as H
    rng = np.random.default_rng(300 + trial)
    line = random_line(rng, int(rng.integers(2, 9)), int(rng.integers(1, 4)), run_to=250.0, warm_up=25.0,
                       jm2dm=bool(rng.random() < .3), up_down=bool(rng.random() < .5), log=True,
                       max_lines=abi.UINT64_MAX)
    want = oracle.run(line, 2, mode=trial % 2, seed=trial, log_capacity=200_000)
    got = H.run(line, 2, mode=trial % 2, seed=trial, log_capacity=200_000)
    assert_same_log(got.log, want.log)


def test_max_lines_overshoots_by_whole_batches(oracle):
    """print-now? is tested before a batch and the whole batch is printed (util/log.clj:225,118-127)."""
    line = example_line(run_to=1500, warm_up=100, log=True, max_lines=100)
    r = oracle.run(line, 4, seed=SEED, log_capacity=10_000)
    per = np.bincount(r.log["instance"])
    assert np.all(per >= 100) and np.all(per < 100 + 3 * line.M)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mk,mode", CASES, ids=[c[0] for c in CASES])
def test_gpu_log_matches_oracle(engine, oracle, name, mk, mode):
    line = mk()
    n = 37
    want = oracle.run(line, n, mode=mode, seed=SEED, log_capacity=4_000_000, n_threads=8)
    got = engine.run(line, n, mode=mode, seed=SEED, log_capacity=4_000_000)
    assert got.rc == abi.OK
    assert_same_log(got.log, want.log)
    assert np.array_equal(got.stats.event_hash, want.stats.event_hash)
    assert np.array_equal(got.stats.occ_time, want.stats.occ_time)


@pytest.mark.gpu
@pytest.mark.parametrize("gid", ["G8", "G9"])
def test_gpu_reliable_cycles(engine, gid):
    """log_test.clj reliable-BAS-pattern / reliable-BBS-pattern straight from the device record stream."""
    g = GOLDENS[gid]
    got = engine.run(golden_two_machine(gid, run_to=1000.0), 1, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=10_000)
    log = sorted_log(got.log)[: g["pulled"]]
    log = log[log["clk"] >= g["from_clk"]]
    groups = [log[i:i + 6] for i in range(0, len(log) - len(log) % 6, 6)]
    assert len(groups) > 50
    for grp in groups:
        assert [pretty_act(x) for x in grp] == g["cycle"]
        assert abs((grp[-1]["clk"] - grp[0]["clk"]) - g["span"]) < g["tol"]


@pytest.mark.gpu
def test_gpu_log_truncation_is_reported_not_overrun(engine, oracle):
    line = example_line(run_to=1500, warm_up=100, log=True, up_down=True, max_lines=abi.UINT64_MAX)
    full = oracle.run(line, 16, seed=SEED, log_capacity=2_000_000, n_threads=8)
    cap = len(full.log) // 3
    got = engine.run(line, 16, seed=SEED, log_capacity=cap)
    assert got.rc == abi.LOG_TRUNCATED
    assert len(got.log) == cap
    # what was written is a subset of the true stream
    want = {(int(r["instance"]), int(r["line"])): r for r in full.log}
    for r in got.log[:: max(1, cap // 500)]:
        w = want[(int(r["instance"]), int(r["line"]))]
        assert r["clk"] == w["clk"] and r["act"] == w["act"] and r["machine"] == w["machine"] and r["job"] == w["job"]
    assert np.array_equal(got.stats.event_hash, full.stats.event_hash)      # statistics are unaffected


@pytest.mark.gpu
def test_gpu_log_count_without_copy_and_no_hash_variants(oracle):
    """The capture variants an un-instrumented handle launches (no event hash; register-resident :ends for
    M <= 8, generic above), and Handle.log_count(): the record count without moving the stream to the host."""
    from mjpdes_b200 import engine as E
    with E.Handle(device=0) as h:
        for line, n in ((example_line(run_to=900, warm_up=100, log=True, up_down=True, max_lines=abi.UINT64_MAX), 333),
                        (mixed20_line(warm_up=20, run_to=200), 70)):
            want = oracle.run(line, n, seed=SEED, log_capacity=3_000_000, n_threads=8)
            h.upload(line, n)
            h.reserve_log(3_000_000)
            h.launch(abi.RNG_COUNTER, SEED, 0, want_log=True)
            h.sync()
            assert h.log_count() == (len(want.log), 0)
            got = h.download(log_capacity=3_000_000)
            assert_same_log(got.log, want.log)
            assert np.array_equal(got.stats.blocked_time, want.stats.blocked_time)


@pytest.mark.gpu
def test_gpu_c4_shape_log_volume(engine):
    """C4 (20 machines, mixed BAS/BBS, 8 job types, SCADA on): record count per instance is what the
    executed actions imply; every instance's lines are 0..k-1 with non-decreasing clocks."""
    line = mixed20_line(warm_up=100, run_to=1100)
    n = 256
    got = engine.run(line, n, seed=SEED, log_capacity=40_000_000)
    assert got.rc == abi.OK
    log = sorted_log(got.log)
    per = np.bincount(log["instance"], minlength=n)
    assert np.all(per > 0)
    starts = np.concatenate([[0], np.cumsum(per)[:-1]])
    assert np.array_equal(log["line"], np.arange(len(log)) - np.repeat(starts, per))
    same = log["instance"][1:] == log["instance"][:-1]
    assert np.all(log["clk"][1:][same] >= log["clk"][:-1][same])
    assert np.all(log["clk"] >= line.warm_up)
    # every action logs one message; cancelled :bl/:ub and :st/:us pairs and the warm-up are the only losses
    assert np.all(per <= got.stats.n_events) and per.sum() > 0.75 * got.stats.n_events.sum() * (1000.0 / 1100.0)


def test_lattice_line_fills_a_tick_with_messages(oracle):
    """Identical machines on a w/W lattice: nearly every machine acts at the same instants, ~5 messages per
    machine per tick (found by differential fuzzing, case 957: the staging area must hold 9 per machine)."""
    from hostsim import binding as H
    rng = np.random.default_rng(100000 + 957)
    M = int(rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 20, 31, 32, 33, 40, 64]))
    K = int(rng.integers(1, 6))
    kw = dict(jm2dm=bool(rng.random() < .3), up_down=bool(rng.random() < .4), bbs_prob=float(rng.choice([0, .3, .7, 1.0])),
              reliable_prob=float(rng.choice([0, .1, .5])))
    run_to = float(rng.choice([60, 150, 400]))
    line = random_line(rng, M, K, run_to=run_to, warm_up=run_to * float(rng.choice([0, 0.1, 0.5])), **kw)
    line.W[:] = 1.0
    line.w[:] = 1.0
    line.log, line.max_lines = True, abi.UINT64_MAX
    assert M == 16
    want = oracle.run(line, 1, mode=abi.RNG_REPLAY, seed=957, log_capacity=300_000)
    assert want.extra["max_tick_msgs"].max() > 3 * M + 8
    got = H.run(line, 1, mode=abi.RNG_REPLAY, seed=957, log_capacity=300_000)
    assert set(got.status.tolist()) == {abi.INST_NORMAL_END}
    assert_same_log(got.log, want.log)
