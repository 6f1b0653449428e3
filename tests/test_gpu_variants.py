"""Parity of the kernel instantiations that are actually BENCHMARKED and shipped by default: handles opened
WITHOUT the event hash (template flag HASH = 0, what bench.py and model.main_loop use).  Every accumulator,
n_events, current_job, final_clock and status is compared bit-for-bit with the oracle for each variant family
(<u32, MT = 2..8>, generic u32, COLD u32, u64, COLD u64, SLICE), in COUNTER and REPLAY mode; then the named
configurations at their FULL horizons, C2's confidence-interval test as SURVEY.md 8(d) states it, and the
multi-GPU entry points (mjp_multi_run behind the C ABI, dist.run_sharded)."""
import os

import numpy as np
import pytest

from lines import SEED, example_line, long50_line, mixed20_line, random_line, sweep10_line
from mjpdes_b200 import abi, dist
from mjpdes_b200 import engine as eng
from parity import assert_same

pytestmark = pytest.mark.gpu
MODES = [abi.RNG_COUNTER, abi.RNG_REPLAY]
CORES = max(1, len(os.sched_getaffinity(0)))


@pytest.fixture(scope="module")
def plain():
    """The handle bench.py opens: no event hash -> mjp_line_kernel<.., HASH = 0, ..>."""
    h = eng.Handle(device=0, event_hash=False)
    yield h
    h.close()


def same(got, want):
    assert_same(got, want, hash_checked=False)
    assert not got.event_hash.any()              # the plane is zero-filled when the hash is off


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7, 8])
def test_register_ends_variants_hash_off(plain, oracle, M, mode):
    """<u32, MT = M>: lines of 2..8 machines (M = 5 with example.clj itself: the headline instantiation)."""
    if M == 5:
        line = example_line(run_to=3000.0, warm_up=300.0)
    else:
        line = random_line(np.random.default_rng(100 + M), M, 1 + M % 3, run_to=600.0, warm_up=60.0,
                           up_down=bool(M & 1), jm2dm=M == 6)
    n = 333
    same(plain.run(line, n, mode=mode, seed=SEED, first_instance_id=99).stats,
         oracle.run(line, n, mode=mode, seed=SEED, first_instance_id=99, n_threads=CORES).stats)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", ["generic_u32", "cold_u32", "u64", "cold_u64"])
def test_wide_variants_hash_off(plain, oracle, shape, mode):
    if shape == "generic_u32":
        line, n = random_line(np.random.default_rng(7), 17, 3, run_to=300.0, warm_up=30.0, up_down=True), 150
    elif shape == "cold_u32":
        n = 50021                                               # more than one wave of the shared layout -> COLD
        line = sweep10_line(n, warm_up=10.0, run_to=60.0, first=11)
    elif shape == "u64":
        line, n = random_line(np.random.default_rng(8), 44, 4, run_to=150.0, warm_up=15.0), 90
    else:
        line, n = long50_line(warm_up=10.0, run_to=60.0), 12001
    same(plain.run(line, n, mode=mode, seed=5).stats, oracle.run(line, n, mode=mode, seed=5, n_threads=CORES).stats)


@pytest.mark.parametrize("mode", MODES)
def test_slice_variants_hash_off(plain, oracle, mode):
    for line, n in ((example_line(run_to=700.0, warm_up=70.0), 200),
                    (random_line(np.random.default_rng(12), 24, 2, run_to=160.0, warm_up=16.0), 70),
                    (random_line(np.random.default_rng(13), 40, 2, run_to=120.0, warm_up=12.0), 40)):
        plain.upload(line, n)
        resume = False
        for until in (25.0, 25.0, 111.1, 400.0, float("inf")):
            plain.launch_slice(until, resume, mode=mode, seed=3, first_instance_id=17)
            plain.sync()
            resume = True
        same(plain.download().stats, oracle.run(line, n, mode=mode, seed=3, first_instance_id=17, n_threads=CORES).stats)


def test_headline_batch_spot_check(plain, oracle):
    """The first 256 instances of the headline batch (example.clj, full horizon, the ids bench.py's first timed
    step uses) against the oracle -- the check bench.py itself repeats as `parity_spot`."""
    line = example_line()
    n = 100_000
    plain.upload(line, n)
    plain.launch(abi.RNG_COUNTER, SEED, 0)
    plain.sync()
    got = plain.download().stats
    want = oracle.run(line, 256, mode=abi.RNG_COUNTER, seed=SEED, n_threads=CORES).stats
    for f in ("status", "n_events", "current_job", "njobs", "final_clock", "residence_sum"):
        assert np.array_equal(getattr(got, f)[:256], getattr(want, f))
    for f in ("blocked_time", "starved_time", "occ_time"):
        assert np.array_equal(getattr(got, f)[:, :256], getattr(want, f))


# ---- C2 exactly as SURVEY.md 8(d) states it -------------------------------------------------------------

def _metrics(line, s):
    run_time = line.run_to - line.warm_up
    lv = line.level_offsets
    out = [("TP", s.njobs / run_time), ("residence", s.residence_sum / s.njobs)]
    for i in range(line.M):
        out.append((f"blocked{i}", s.blocked_time[i] / run_time))
        out.append((f"starved{i}", s.starved_time[i] / run_time))
    for b in range(line.M - 1):
        lev = np.arange(line.N[b] + 1)[:, None]
        out.append((f"occ{b}", (lev * s.occ_time[lv[b]:lv[b + 1]]).sum(0) / run_time))
    return out


def test_c2_full_spec_confidence_intervals(plain, oracle):
    """C2: example.clj x 100 000 COUNTER-mode replications on the GPU against 2 000 seed-disjoint oracle
    replications drawn with the reference's literal accumulate-until loop (REPLAY), full horizon 2000 / 20000.
    Two-sample Welch interval at 99 % (z = 2.576), metric by metric.  blocked[last] and starved[first] are
    identically zero on both sides and are compared for equality instead."""
    line = example_line()
    g = plain.run(line, 100_000, mode=abi.RNG_COUNTER, seed=SEED).stats
    o = oracle.run(line, 2000, mode=abi.RNG_REPLAY, seed=SEED + 1, first_instance_id=10 ** 7, n_threads=CORES).stats
    assert np.all(g.status == abi.INST_NORMAL_END) and np.all(o.status == abi.INST_NORMAL_END)
    z = 2.576
    failed = []
    for (name, x), (_, y) in zip(_metrics(line, g), _metrics(line, o)):
        se = np.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y))
        if se == 0.0:
            assert x.mean() == y.mean() == 0.0, name
            continue
        d = abs(x.mean() - y.mean())
        if d > z * se:
            failed.append((name, float(d / se)))
    assert not failed, failed


# ---- the wide configurations at their full horizons -------------------------------------------------------

def test_c3_full_horizon(plain, oracle):
    """C3: 256 points of the 10-machine sweep, warm-up 2000, run-to 20000."""
    n = 256
    line = sweep10_line(n, first=424242)
    same(plain.run(line, n, seed=SEED, first_instance_id=424242).stats,
         oracle.run(line, n, seed=SEED, first_instance_id=424242, n_threads=CORES).stats)


def test_c4_full_horizon_with_scada(plain, oracle):
    """C4: 20 machines, mixed BAS/BBS, 8 job types, :log? true :up&down? true, warm-up 2000, run-to 20000,
    64 instances: statistics AND the SCADA record stream (~3.5e7 records) record for record."""
    n = 64
    line = mixed20_line()
    cap = n * 700_000
    got = plain.run(line, n, seed=SEED, log_capacity=cap)
    want = oracle.run(line, n, seed=SEED, log_capacity=cap, n_threads=CORES)
    same(got.stats, want.stats)
    assert got.rc == abi.OK and want.rc == abi.OK and len(got.log) == len(want.log) > 3e7
    order = lambda r: r[np.argsort((r["instance"].astype(np.uint64) << np.uint64(32)) | r["line"].astype(np.uint64))]
    a, b = order(got.log), order(want.log)
    for f in ("instance", "line", "act", "machine", "n", "job", "clk"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["aux"].view(np.uint64), b["aux"].view(np.uint64))


def test_c5_time_boxed_horizon(plain, oracle):
    """C5: 50 machines, 16 job types, warm-up 2000, run-to 2e4 (the time-boxed slice bench.py measures), 32 instances."""
    n = 32
    line = long50_line(warm_up=2000.0, run_to=20000.0)
    same(plain.run(line, n, seed=SEED).stats, oracle.run(line, n, seed=SEED, n_threads=CORES).stats)


# ---- guards of the ABI ----------------------------------------------------------------------

def test_endless_and_non_finite_runs_are_rejected(plain):
    line = example_line(run_to=float("inf"))
    plain.upload(line, 8)                                   # allowed: main-loop-continuous drives it in slices
    with pytest.raises(abi.MjpError) as e:
        plain.launch(abi.RNG_COUNTER, 1, 0)                 # ... but never unsliced
    assert e.value.code == abi.ERR_INVALID
    plain.launch_slice(50.0, False, seed=1)
    plain.sync()
    assert np.all(plain.download_fields(("status",)).status == abi.INST_PAUSED)
    for field, bad in (("lam", float("inf")), ("mu", float("inf")), ("W", float("nan")), ("w", float("inf"))):
        line = example_line()
        getattr(line, field).flat[1] = bad
        with pytest.raises(abi.MjpError):
            plain.upload(line, 8)
    with pytest.raises(abi.MjpError):
        plain.upload(example_line(warm_up=float("inf")), 8)


def test_resume_must_match_the_parking_launch(plain):
    line = example_line(run_to=500.0, warm_up=50.0)
    plain.upload(line, 64)
    plain.launch_slice(100.0, False, mode=abi.RNG_COUNTER, seed=4, first_instance_id=0)
    plain.sync()
    for kw in (dict(mode=abi.RNG_REPLAY, seed=4), dict(mode=abi.RNG_COUNTER, seed=5),
               dict(mode=abi.RNG_COUNTER, seed=4, first_instance_id=64)):
        with pytest.raises(abi.MjpError) as e:
            plain.launch_slice(200.0, True, **kw)
        assert e.value.code == abi.ERR_INVALID
    plain.launch_slice(float("inf"), True, mode=abi.RNG_COUNTER, seed=4, first_instance_id=0)
    plain.sync()
    assert np.all(plain.download_fields(("status",)).status == abi.INST_NORMAL_END)


# ---- multi-GPU -----------------------------------------------------------------------------------------------

def _n_gpus():
    return abi.load_library().mjp_device_count()


@pytest.mark.parametrize("reduce", [abi.REDUCE_PEER, abi.REDUCE_NCCL])
def test_multi_run_behind_the_c_abi(oracle, reduce):
    """mjp_multi_run on every visible GPU (1 on a single-GPU box): per-instance planes and the reduced aggregate
    equal a single-device run of the whole batch; per-instance override planes are cut per device."""
    G = _n_gpus()
    n = 2 * 4099 + 7                                           # ragged shards
    line = sweep10_line(n, warm_up=20.0, run_to=200.0, first=31)
    try:
        m = eng.MultiHandle(reduce=reduce)
    except abi.MjpError as e:
        if reduce == abi.REDUCE_NCCL and e.code == abi.ERR_UNSUPPORTED:
            pytest.skip("libnccl.so.2 not loadable in this process")
        raise
    with m:
        assert m.n_devices == G
        got = m.run(line, n, seed=SEED, first_instance_id=5).stats
    want = oracle.run(line, n, seed=SEED, first_instance_id=5, n_threads=CORES).stats
    same(got, want)
    assert got.aggregate[0] == n
    assert np.allclose(got.aggregate, want.aggregate, rtol=1e-12)
    if G > 1:
        with eng.MultiHandle(devices=[G - 1]) as one:           # any single device gives the same planes
            again = one.run(line, n, seed=SEED, first_instance_id=5).stats
        same(again, want)


def _shard_worker(rank, world, port, n, q):
    import torch
    import torch.distributed as td
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        r = dist.run_sharded(example_line(run_to=800.0, warm_up=80.0), n, seed=SEED, device=torch.device("cuda", rank))
        q.put((rank, r.lo, r.hi, r.aggregate, r.local.n_events))
    finally:
        td.destroy_process_group()


def test_run_sharded_on_gpus(oracle):
    """dist.run_sharded with the CUDA runner: world 2 over NCCL when two GPUs are visible, otherwise world 1."""
    n = 1001
    line = example_line(run_to=800.0, warm_up=80.0)
    want = oracle.run(line, n, seed=SEED, n_threads=CORES).stats
    if _n_gpus() >= 2:
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_shard_worker, args=(r, 2, 29571, n, q)) for r in range(2)]
        [p.start() for p in procs]
        outs = sorted(q.get(timeout=600) for _ in procs)
        [p.join(60) for p in procs]
        assert [p.exitcode for p in procs] == [0, 0]
        assert (outs[0][1], outs[0][2], outs[1][1], outs[1][2]) == (0, 501, 501, 1001)
        for o in outs:
            assert np.allclose(o[3], want.aggregate, rtol=1e-12) and o[3][0] == n
        assert np.array_equal(np.concatenate([outs[0][4], outs[1][4]]), want.n_events)
    else:
        r = dist.run_sharded(line, n, seed=SEED, device=0)
        assert (r.lo, r.hi) == (0, n)
        assert np.allclose(r.aggregate, want.aggregate, rtol=1e-12)
        same(r.local, want)


def test_aggregate_tensor_aliases_the_device_vector(plain):
    import torch
    line = example_line(run_to=400.0, warm_up=40.0)
    plain.upload(line, 500)
    plain.launch(abi.RNG_COUNTER, SEED, 0)
    plain.sync()
    t = plain.aggregate_tensor()
    assert t.is_cuda and t.dtype == torch.float64 and t.numel() == abi.aggregate_len(line.M)
    assert np.array_equal(t.cpu().numpy(), plain.download_fields(("aggregate",)).aggregate)


# ---- rotation: batches a little larger than one wave ----------------------------------------------------------

@pytest.mark.parametrize("mode", MODES)
def test_rotation_on_device(plain, oracle, mode):
    """130 000 instances of example.clj need 1.25 waves of CTAs: the launch rotates them through the resident CTAs
    ((chunk, slice) work items, release/acquire flags between the slices of a chunk).  Results equal the oracle's."""
    n = 130_000
    line = example_line(run_to=200.0, warm_up=20.0)
    got = plain.run(line, n, mode=mode, seed=SEED, first_instance_id=77).stats
    want = oracle.run(line, n, mode=mode, seed=SEED, first_instance_id=77, n_threads=CORES).stats
    same(got, want)


def test_rotation_on_device_wide_and_logged(plain, oracle):
    n = 9472 * 2 + 100                                     # > one wave of the 50-machine COLD layout (6 CTAs/SM)
    line = long50_line(warm_up=10.0, run_to=60.0)
    same(plain.run(line, n, seed=5).stats, oracle.run(line, n, seed=5, n_threads=CORES).stats)
    n = 120_000
    line = example_line(run_to=120.0, warm_up=60.0, log=True, up_down=True, max_lines=abi.UINT64_MAX)
    cap = n * 1000
    got = plain.run(line, n, seed=9, log_capacity=cap)
    want = oracle.run(line, n, seed=9, log_capacity=cap, n_threads=CORES)
    same(got.stats, want.stats)
    assert got.rc == abi.OK and len(got.log) == len(want.log)
    order = lambda r: r[np.argsort((r["instance"].astype(np.uint64) << np.uint64(32)) | r["line"].astype(np.uint64))]
    a, b = order(got.log), order(want.log)
    for f in ("instance", "line", "act", "machine", "n", "job", "clk"):
        assert np.array_equal(a[f], b[f]), f
