"""The N>1 path on CPU: world_size-2 gloo process group, instances sharded by contiguous id range,
aggregate moments all-reduced.  The local runner is the oracle here (no GPU); what is under test is
the host-side sharding, plane slicing and the collective."""
import os
import socket

import numpy as np
import pytest

from lines import SEED, example_line, sweep10_line
from mjpdes_b200 import abi, dist as mdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_runner(shard, n, mode, seed, first_id):
    from oracle import binding
    return binding.run(shard, n, mode=mode, seed=seed, first_instance_id=first_id).stats


def _worker(rank, world, port, which, n_total, out_dir):
    import torch.distributed as td
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    line = example_line(run_to=800.0, warm_up=80.0) if which == "example" else sweep10_line(n_total, 50, 400, first=1000)
    r = mdist.run_sharded(line, n_total, seed=SEED, runner=_oracle_runner)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=r.lo, hi=r.hi, agg=r.aggregate,
             hash=r.local.event_hash if r.local is not None else np.zeros(0, np.uint64))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("which,n_total", [("example", 11), ("sweep", 7)])
def test_two_rank_sharding_and_reduction(tmp_path, oracle, which, n_total):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), which, n_total, str(tmp_path)), nprocs=world, join=True)
    line = example_line(run_to=800.0, warm_up=80.0) if which == "example" else sweep10_line(n_total, 50, 400, first=1000)
    whole = oracle.run(line, n_total, seed=SEED).stats
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert [(int(p["lo"]), int(p["hi"])) for p in parts] == [mdist.shard_range(n_total, world, r) for r in range(world)]
    assert np.array_equal(np.concatenate([p["hash"] for p in parts]), whole.event_hash)     # same instances, same streams
    for p in parts:
        assert p["agg"][0] == n_total
        assert np.allclose(p["agg"], whole.aggregate, rtol=1e-12)                           # every rank holds the sum


def test_shard_ranges_cover_the_batch():
    for n in (1, 7, 8, 100_000, 1_000_003):
        for world in (1, 2, 4, 8):
            spans = [mdist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-n // world)


def test_summary_from_moments():
    line = example_line()
    x = np.random.default_rng(0).normal(0.73, 0.01, 1000)
    agg = np.zeros(abi.aggregate_len(line.M))
    agg[0], agg[1], agg[2] = len(x), x.sum(), (x * x).sum()
    s = mdist.summarize_aggregate(agg, line)
    assert abs(s["TP"]["mean"] - x.mean()) < 1e-12 and abs(s["TP"]["stderr"] - x.std(ddof=1) / np.sqrt(len(x))) < 1e-9
