""":job-detail? / :dets (util/log.clj:11-28,101; core.clj:251,274,295-325): which job is on which machine and in which
buffer just before each logged action.

The oracle takes the snapshots literally (Sim::detail at every log call).  The product reconstructs them on the host
(scada.DetsTracker) from three things the device provides: the records in cleaned order, each record's position in
the tick's EXECUTION order (``ord``), and the state at the instance's first printed message (``first_state``).
CPU tests: the reconstruction against the literal snapshots, and the kernel source (host shim) against the oracle for
``ord`` / ``first_state``.  GPU tests: the same through the C ABI and through model.main_loop's printed text.
"""
import io

import numpy as np
import pytest

from lines import GOLDENS, SEED, example_line, mixed20_line, random_line
from mjpdes_b200 import abi, model, scada


def as_model(d):
    """golden JSON (plain dicts) -> Model with record types, as eval of the file would give."""
    line = {k: (model.Buffer(v) if ":N" in v else model.ExpoMachine(v)) for k, v in d[":line"].items()}
    jm = {k: model.JobType(v) for k, v in d[":jobmix"].items()}
    return model.Model({**d, ":line": line, ":jobmix": jm})


class _CL:
    """The three name lists scada needs of a CompiledLine, for bare HostLine cases."""

    def __init__(self, M, mode="full"):
        self.machines = [f":m{i + 1}" for i in range(M)]
        self.buffers = [f":b{i + 1}" for i in range(M - 1)]
        self.jobtypes = [f":jobType{k + 1}" for k in range(64)]
        self.model = {":job-detail?": mode != "strip", ":report": {":job-detail?": mode == "full"}}


def lattice_line(M=12, **kw):
    """Identical reliable machines and unit work: nearly every tick holds several movements whose cleaned order
    (by machine) differs from their execution order, and ticks with more than 8 machines are regrouped by hash."""
    return abi.HostLine(lam=np.zeros(M), mu=np.ones(M), W=np.ones(M), discipline=(np.arange(M) % 3 == 0),
                        N=np.full(M - 1, 2), portion=[1.0], w=np.ones((1, M)), **kw)


CASES = [
    ("example", lambda: example_line(run_to=700, warm_up=100, log=True, up_down=True, max_lines=abi.UINT64_MAX), 3),
    ("example-max-lines", lambda: example_line(run_to=700, warm_up=100, log=True, max_lines=57), 2),
    ("mixed20", lambda: mixed20_line(warm_up=40, run_to=260, log=True, up_down=True, max_lines=abi.UINT64_MAX), 2),
    ("lattice12", lambda: lattice_line(12, warm_up=20.5, run_to=120, log=True, max_lines=abi.UINT64_MAX), 1),
    ("lattice5-no-warm-up", lambda: lattice_line(5, warm_up=0.0, run_to=60, log=True, max_lines=abi.UINT64_MAX), 1),
]


def literal(cl, want):
    """The oracle's literal snapshots as the maps log/detail returns."""
    from mjpdes_b200 import cljhash
    order = lambda d: {k: d[k] for k in cljhash.map_order(list(d.keys()))}
    return [{":run": order(dict(zip(cl.machines, run))), ":bufs": order(dict(zip(cl.buffers, bufs)))}
            for run, bufs in want.dets]


def reconstructed(cl, log, first_state):
    recs = np.sort(log, order=["instance", "line"])
    return [m[":dets"] for m in scada.message_maps(cl, recs, first_state)]


@pytest.mark.parametrize("name,mk,n", CASES, ids=[c[0] for c in CASES])
def test_reconstruction_equals_the_literal_snapshots(oracle, name, mk, n):
    line = mk()
    want = oracle.run(line, n, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=400_000, job_detail=True)
    assert want.rc == abi.OK and len(want.dets) == len(want.log) > 50
    cl = _CL(line.M)
    got = reconstructed(cl, want.log, want.first_state)
    assert got == literal(cl, want)
    # the snapshots are not trivially derivable from cleaned order: some tick was executed in another order
    recs = np.sort(want.log, order=["instance", "line"])
    if name.startswith("lattice") or name == "mixed20":
        same_tick = (recs["clk"][1:] == recs["clk"][:-1]) & (recs["instance"][1:] == recs["instance"][:-1])
        assert np.any(same_tick & (recs["ord"][1:] < recs["ord"][:-1]))


def test_random_lines_reconstruct(oracle):
    rng = np.random.default_rng(11)
    for trial in range(12):
        M = int(rng.integers(2, 15))
        line = random_line(rng, M, int(rng.integers(1, 4)), run_to=160.0, warm_up=float(rng.choice([0.0, 25.0])),
                           jm2dm=bool(rng.random() < .3), up_down=bool(rng.random() < .5), log=True,
                           max_lines=int(rng.choice([40, abi.UINT64_MAX])))
        want = oracle.run(line, 2, seed=100 + trial, log_capacity=200_000, job_detail=True)
        cl = _CL(M)
        assert reconstructed(cl, want.log, want.first_state) == literal(cl, want)


def test_kernel_source_gives_the_same_ordinals_and_first_state(oracle):
    from hostsim import binding as H
    for name, mk, n in CASES:
        line = mk()
        want = oracle.run(line, n, seed=SEED, log_capacity=400_000)
        for cold in ((False, True) if line.M > 8 else (False,)):
            got = H.run(line, n, seed=SEED, log_capacity=400_000, cold=cold)
            a, b = np.sort(got.log, order=["instance", "line"]), np.sort(want.log, order=["instance", "line"])
            assert np.array_equal(a["ord"], b["ord"]) and np.array_equal(a["n"], b["n"]), name
            assert np.array_equal(got.first_state, want.first_state), name
        # sliced: the state at the first printed message is taken once and survives the park
        got = H.run(line, n, seed=SEED, log_capacity=400_000, slices=[line.warm_up * 0.5, line.warm_up + 7.0, 100.0])
        assert np.array_equal(got.first_state, want.first_state), name


def test_nothing_printed_leaves_first_state_unset(oracle):
    line = example_line(run_to=50, warm_up=100, log=True, max_lines=abi.UINT64_MAX)      # ends before the warm-up
    want = oracle.run(line, 2, seed=SEED, log_capacity=1000)
    assert len(want.log) == 0 and np.all(want.first_state == 0xFFFFFFFF)
    with pytest.raises(ValueError):
        scada.DetsTracker(_CL(5), want.first_state[0])


def test_printed_form_of_dets():
    """print-lines (util/log.clj:111-137): ~A of the flattened message, :dets between :j and :line."""
    rec = np.zeros(1, abi.LOG_RECORD_DTYPE)[0]
    rec["clk"], rec["act"], rec["machine"], rec["n"], rec["job"], rec["line"] = 2012.25, abi.ACT_SM, 1, 2, 1715, 4
    cl = _CL(3)
    dets = {":run": {":m1": 1718, ":m2": None, ":m3": 1714}, ":bufs": {":b1": [1715, 1716], ":b2": []}}
    text = scada.format_line(scada.message_map(cl, rec, dets))
    assert text == ("{:clk 2012.2500 :act :m2-start-job      :m :m2 :mjpact :sm :bf :b1 :n 2 :j 1715 "
                    ":dets {:run {:m1 1718, :m2 nil, :m3 1714}, :bufs {:b1 [1715 1716], :b2 []}} :line 4}")
    assert scada.format_line(scada.message_map(cl, rec, None)).endswith(":j 1715 :dets nil :line 4}")
    assert ":dets" not in scada.format_line(scada.message_map(cl, rec))
    assert scada.job_detail_mode({":report": {":job-detail?": True}}) == "strip"          # util/log.clj:101
    assert scada.job_detail_mode({":job-detail?": True}) == "nil"
    assert scada.job_detail_mode({":job-detail?": True, ":report": {":job-detail?": True}}) == "full"


def test_wrapped_ordinal_is_refused():
    cl = _CL(2)
    t = scada.DetsTracker(cl, [0, 0, 0, 0])
    recs = np.zeros(2, abi.LOG_RECORD_DTYPE)
    recs["act"] = abi.ACT_ST
    with pytest.raises(ValueError):
        t.tick(recs)


# ---- GPU: through the C ABI -------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name,mk,n", CASES, ids=[c[0] for c in CASES])
def test_device_ordinals_first_state_and_dets(engine, oracle, name, mk, n):
    line = mk()
    n = n * 40 + 3                                          # several warps, a ragged last one
    want = oracle.run(line, n, seed=SEED, log_capacity=3_000_000, job_detail=True, n_threads=1)
    got = engine.run(line, n, seed=SEED, log_capacity=3_000_000)
    a, b = np.sort(got.log, order=["instance", "line"]), np.sort(want.log, order=["instance", "line"])
    assert len(a) == len(b)
    for f in a.dtype.names:
        assert np.array_equal(a[f].view(np.uint64) if f in ("clk", "aux") else a[f], b[f].view(np.uint64) if f in ("clk", "aux") else b[f]), f
    assert np.array_equal(got.first_state, want.first_state)
    cl = _CL(line.M)
    assert reconstructed(cl, got.log, got.first_state) == literal(cl, want)


@pytest.mark.gpu
def test_first_state_survives_slicing_on_the_device(engine, oracle):
    line = mixed20_line(warm_up=40, run_to=200, log=True, up_down=True, max_lines=abi.UINT64_MAX)
    n = 70
    want = oracle.run(line, n, seed=SEED, log_capacity=2_000_000)
    engine.upload(line, n)
    engine.reserve_log(2_000_000)
    parts, resume = [], False
    for until in (20.0, 41.0, 90.0, float("inf")):
        engine.launch_slice(until, resume, seed=SEED, want_log=True)
        engine.sync()
        res = engine.download(log_capacity=2_000_000)
        parts.append(res.log)
        resume = True
    assert np.array_equal(res.first_state, want.first_state)
    log = np.concatenate(parts)
    a, b = np.sort(log, order=["instance", "line"]), np.sort(want.log, order=["instance", "line"])
    assert np.array_equal(a["ord"], b["ord"]) and np.array_equal(a["job"], b["job"])


@pytest.mark.gpu
def test_main_loop_prints_dets_like_the_reference(engine, oracle):
    d = dict(GOLDENS["models"]["resources/example.clj"])
    d[":params"] = {":warm-up-time": 100, ":run-to-time": 400}
    d[":report"] = {":log?": True, ":max-lines": 300, ":up&down?": True, ":job-detail?": True, ":diag-log-buf?": True}
    d[":job-detail?"] = True
    out = io.StringIO()
    res = model.main_loop(as_model(d), handle=engine, seed=SEED, out_stream=out)
    lines = [l for l in out.getvalue().split("\n") if l.startswith("{:clk")]
    cl = model.preprocess_model(as_model(d), seed=SEED)
    want = oracle.run(cl.host, 1, seed=SEED, log_capacity=100_000, job_detail=True)
    lit = literal(cl, want)
    recs = np.sort(want.log, order="line")
    assert lines == [scada.format_line(scada.message_map(cl, r, dt)) for r, dt in zip(recs, lit)]
    assert len(lines) >= 300 and all(" :dets {:run {:m1 " in l for l in lines)
    assert [m[":dets"] for m in res[":diag-log-buf"]] == lit
    # top-level flag only: every message prints `:dets nil` (util/log.clj:14 vs :101)
    d[":report"] = {":log?": True, ":max-lines": 50}
    out = io.StringIO()
    model.main_loop(as_model(d), handle=engine, seed=SEED, out_stream=out)
    lines = [l for l in out.getvalue().split("\n") if l.startswith("{:clk")]
    assert len(lines) >= 50 and all(" :dets nil :line " in l for l in lines)


@pytest.mark.gpu
def test_streaming_carries_dets_from_slice_to_slice(engine, oracle):
    d = dict(GOLDENS["models"]["data/quick.clj"])
    d[":report"] = {":continuous?": True, ":up&down?": True, ":job-detail?": True}
    d[":job-detail?"] = True
    cm = model.main_loop(as_model(d), handle=engine, seed=SEED, mode=abi.RNG_REPLAY)
    msgs = model.pull_data(cm, 1500)
    line = cm.cl.host
    line.run_to = float(msgs[-1][":clk"]) + 50.0
    want = oracle.run(line, 1, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=100_000, job_detail=True)
    assert [m[":dets"] for m in msgs] == literal(cm.cl, want)[:1500]


def test_literal_snapshots_agree_with_the_messages_they_ride_on(oracle):
    """What any correct (log/detail model) must satisfy, whatever order a tick was executed in: the snapshot is taken
    BEFORE the action (core.clj:251,274,295-297), so a :bj / :ej of job j sees j on its machine, a :sm sees j at the
    head of the feed buffer and nothing on the machine, an :aj sees the entry machine free, and :n is the length of
    the buffer the job is about to enter or leave."""
    for name, mk, n in CASES:
        line = mk()
        want = oracle.run(line, n, seed=SEED + 1, log_capacity=400_000, job_detail=True)
        recs = np.sort(want.log, order=["instance", "line"])
        assert len(recs) == len(want.dets)
        for r, (run, bufs) in zip(recs, want.dets):
            act, m, j, nn = int(r["act"]), int(r["machine"]), int(r["job"]), int(r["n"])
            if act == abi.ACT_AJ:
                assert run[0] is None
            elif act == abi.ACT_BJ:
                assert run[m] == j and len(bufs[m]) == nn and j not in bufs[m]
            elif act == abi.ACT_SM:
                assert run[m] is None and bufs[m - 1][0] == j and len(bufs[m - 1]) == nn
            elif act == abi.ACT_EJ:
                assert run[m] == j and m == line.M - 1
            for b, ids in enumerate(bufs):                   # a buffer holds consecutive jobs, oldest first
                assert ids == list(range(ids[0], ids[0] + len(ids))) if ids else True
                assert len(ids) <= int(line.N[b])
