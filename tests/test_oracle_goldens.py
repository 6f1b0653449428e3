"""The CPU oracle against every golden vector the reference's own tests hold for the
run path.  Fixtures: tests/golden/reference_goldens.json,
extracted from the reference test sources by tests/golden/extract_goldens.py."""
import math

import numpy as np
import pytest

from lines import GOLDENS, SEED, example_line, golden_two_machine, pretty_act
from mjpdes_b200 import abi

ACT = {":aj": abi.ACT_AJ, ":bj": abi.ACT_BJ, ":ej": abi.ACT_EJ, ":sm": abi.ACT_SM, ":st": abi.ACT_ST,
       ":us": abi.ACT_US, ":bl": abi.ACT_BL, ":ub": abi.ACT_UB, ":up": abi.ACT_UP, ":down": abi.ACT_DOWN}
KIND = {":up": 0, ":down": 1}


@pytest.mark.parametrize("gid", ["G1", "G2"])
def test_end_time_known_answer(oracle, gid):
    """core_test.clj job-tests-1/2: end-time over a literal :up&down schedule."""
    g = GOLDENS[gid]
    kinds = [KIND[k] for _, k, _ in g["events"]]
    times = [t for _, _, t in g["events"]]
    got = oracle.end_time(kinds, times, g["clock"], g["dur"])
    assert abs(got - g["expected"]) < g["tol"]
    assert got == g["expected"]          # in fact to the last bit


def test_end_time_comment_value():
    assert GOLDENS["G1"]["expected"] == GOLDENS["G1"]["expected_comment"]


@pytest.mark.parametrize("gid", ["G3", "G4"])
def test_block_starve_accounting(oracle, gid):
    """core_test.clj log-testing-1/2: 2 time units of 5 -> 0.4 (== in the reference)."""
    g = GOLDENS[gid]
    script = [("push",) if s[0] == "push" else ("log", ACT[s[1]], 0, s[3]) for s in g["script"]]
    stats, _ = oracle.log_script(1, [], g["warm_up"], script)
    acc = stats.blocked_time if g["metric"] == ":blocked" else stats.starved_time
    assert acc[0, 0] / (g["run_to"] - g["warm_up"]) == g["expected"]


def test_clean_log_buf_keeps_four(oracle):
    """core_test.clj:113-118: four same-tick :bl/:ub msgs of one machine are NOT cancelled, two are."""
    four = [("log", abi.ACT_BL, 0, 0.5), ("log", abi.ACT_UB, 0, 0.5), ("log", abi.ACT_BL, 0, 0.5),
            ("log", abi.ACT_UB, 0, 0.5), ("push",), ("log", abi.ACT_AJ, 0, 1.0), ("push",)]
    _, log = oracle.log_script(1, [], 0.0, four)
    assert [int(r["act"]) for r in log] == [abi.ACT_BL, abi.ACT_UB, abi.ACT_BL, abi.ACT_UB]
    two = [("log", abi.ACT_BL, 0, 0.5), ("log", abi.ACT_UB, 0, 0.5), ("log", abi.ACT_ST, 0, 0.5), ("push",),
           ("log", abi.ACT_AJ, 0, 1.0), ("push",)]
    _, log = oracle.log_script(1, [], 0.0, two)
    assert [int(r["act"]) for r in log] == [abi.ACT_ST]


def test_updown_generator_efficiency(oracle):
    """core_test.clj test-efficiency: strict alternation; up fraction in (0.89, 0.91) over 1e6 events."""
    for mode in (abi.RNG_REPLAY, abi.RNG_COUNTER):
        kinds, times, du, de = oracle.updown_events(0.1, 0.9, 1_000_001, mode, seed=SEED)
        assert np.all(kinds[1:] != kinds[:-1])
        assert np.all(np.diff(times) > 0)
        dt = np.diff(times)
        up = dt[kinds[1:] == 1].sum()          # interval ending in a :down event was up time
        down = dt[kinds[1:] == 0].sum()
        assert 0.89 < up / (up + down) < 0.91
        # Erlang-2 sojourns: mean 2/r, var 2/r^2
        for kind, rate in ((1, 0.1), (0, 0.9)):
            x = dt[kinds[1:] == kind]
            assert abs(x.mean() - 2 / rate) < 0.02 * 2 / rate
            assert abs(x.var() - 2 / rate ** 2) < 0.05 * 2 / rate ** 2
        if mode == abi.RNG_REPLAY:             # 1 uniform + ~2 exponentials per sojourn
            assert du == 1_000_001 + 0 and 1.95e6 < de < 2.05e6


def test_updown_reliable_machine(oracle):
    """core.clj:178-179: lambda == 0 -> the whole seq is ([0 :down ##Inf])."""
    for mode in (abi.RNG_REPLAY, abi.RNG_COUNTER):
        kinds, times, _, _ = oracle.updown_events(0.0, 0.9, 5, mode, seed=SEED)
        assert list(kinds) == [1] and math.isinf(times[0])


def test_calc_bneck(oracle):
    """core_test.clj bneck-test."""
    g = GOLDENS["G6"]
    for tp, exp in ((g["tp1"], g["tp1_expected"]), (g["tp2"], g["tp2_expected"])):
        bl = [tp[":blocked"][m] for m in g["machines"]]
        st = [tp[":starved"][m] for m in g["machines"]]
        assert g["machines"][oracle.calc_bneck(bl, st)] == exp


def _records_as_msgs(log, bufname=lambda b: f":b{b + 1}"):
    out = []
    for r in log:
        act, m = int(r["act"]), int(r["machine"])
        d = {":act": abi.ACT_KEYWORD[act], ":clk": float(r["clk"])}
        if act == abi.ACT_AJ:
            d.update({":j": int(r["job"]), ":jt": ":jobType1", ":ends": float(r["aux"])})
        elif act == abi.ACT_BJ:
            d.update({":bf": bufname(m), ":j": int(r["job"]), ":n": int(r["n"])})
        elif act == abi.ACT_SM:
            d.update({":bf": bufname(m - 1), ":j": int(r["job"]), ":n": int(r["n"])})
        elif act == abi.ACT_EJ:
            d.update({":m": f":m{m + 1}", ":j": int(r["job"]), ":ent": float(r["aux"])})
        else:
            d[":m"] = f":m{m + 1}"
        out.append(d)
    return out


def test_load_m1_messages(oracle):
    """core_test.clj:251-287: the documented 15-message log of the 2-machine w=(2,5) run to t=10
    (lambda := 0 to make it deterministic, as SURVEY G7 prescribes)."""
    line = golden_two_machine("G7", lam=0.0)
    r = oracle.run(line, 1, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=1000)
    got = [m for m in _records_as_msgs(r.log) if m[":clk"] <= 9.0]
    key = lambda d: sorted(d.items())
    want = GOLDENS["G7"]["messages"]
    assert sorted(map(key, got)) == sorted(map(key, want))
    assert len(got) == 15


@pytest.mark.parametrize("gid", ["G8", "G9"])
def test_reliable_cycle(oracle, gid):
    """log_test.clj reliable-BAS-pattern / reliable-BBS-pattern: the steady 6-tuple and its span."""
    g = GOLDENS[gid]
    line = golden_two_machine(gid, run_to=1000.0)
    r = oracle.run(line, 1, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=5000)
    log = r.log[: g["pulled"]]                       # (pull-data! model 500)
    log = log[log["clk"] >= g["from_clk"]]
    groups = [log[i:i + 6] for i in range(0, len(log) - len(log) % 6, 6)]
    assert len(groups) > 50
    for grp in groups:
        assert [pretty_act(x) for x in grp] == g["cycle"]
        assert abs((grp[-1]["clk"] - grp[0]["clk"]) - g["span"]) < g["tol"]
    assert np.array_equal(r.log["line"], np.arange(len(r.log)))


def test_example_event_counts(oracle):
    """~2.20e5 events, ~1.98e5 iterations, ~8.2e4 ticks per example.clj replication."""
    r = oracle.run(example_line(), 8, mode=abi.RNG_REPLAY, seed=SEED, n_threads=4)
    assert np.all(r.stats.status == abi.INST_NORMAL_END)
    ev = r.stats.n_events.astype(float)
    assert abs(ev.mean() - 2.20e5) < 0.01 * 2.20e5
    assert abs(r.extra["iterations"].mean() - 1.98e5) < 0.01 * 1.98e5
    assert abs(r.extra["ticks"].mean() - 8.2e4) < 0.01 * 8.2e4
    assert np.all(r.extra["act_counts"].sum(axis=0) == r.stats.n_events)
    # duplicate :up quirk: more ups than downs
    assert r.extra["act_counts"][abi.ACT_UP].mean() > 1.1 * r.extra["act_counts"][abi.ACT_DOWN].mean()


def test_example_statistics_vs_recorded_outputs(oracle):
    """G10, loose sanity vs data/new-out.clj and data/ex1-out-out.clj (excluding :blocked)."""
    line = example_line()
    r = oracle.run(line, 16, mode=abi.RNG_COUNTER, seed=SEED, n_threads=4)
    b = [oracle.calc_basics(line, r.stats, g) for g in range(16)]
    tp = np.mean([x["TP"] for x in b])
    assert 0.72 < tp < 0.75                     # recorded: 0.7426, 0.7301
    st = np.mean([x["starved"] for x in b], axis=0)
    assert np.allclose(st, [0.0, 0.0056, 0.111, 0.186, 0.305], atol=0.02)      # data/new-out.clj:15-21
    occ = np.mean([x["occupancy"] for x in b], axis=0)
    assert np.allclose(occ, [2.76, 1.32, 0.33, 0.144], atol=0.35)              # data/new-out.clj:8-12
    res = np.mean([x["residence"] for x in b])
    assert abs(res - 12.2) < 0.8                                               # data/new-out.clj:22
