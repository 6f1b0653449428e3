"""Multi-GPU sharding of a batch of line instances.

Instances are independent, so the batch is split into contiguous ranges of instance ids,
``ceil(n / G)`` per rank; an instance's RNG substreams depend on its id only, so results do
not depend on G.  There is no data-path collective.  The ONE exchange is the final reduction
of the aggregate moments (count, sum x, sum x^2 per metric): ``all_reduce(SUM)`` over
``torch.distributed`` -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    per = -(-int(n_total) // int(world))
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def slice_line(line: abi.HostLine, lo: int, hi: int) -> abi.HostLine:
    """The descriptor of instances [lo, hi): shared vectors as they are, per-instance planes cut."""
    cut = lambda a: None if a is None else np.ascontiguousarray(a[:, lo:hi])
    return abi.HostLine(lam=line.lam, mu=line.mu, W=line.W, discipline=line.discipline, N=line.N,
                        portion=line.portion, w=line.w, warm_up=line.warm_up, run_to=line.run_to,
                        run_to_job=line.run_to_job, jm2dm=line.jm2dm, log=line.log, up_down=line.up_down,
                        max_lines=line.max_lines, lam_inst=cut(line.lam_inst), mu_inst=cut(line.mu_inst),
                        W_inst=cut(line.W_inst), N_inst=cut(line.N_inst), w_inst=cut(line.w_inst),
                        group_rank=line.group_rank)


def all_reduce_sum(vec: np.ndarray, group=None, device=None) -> np.ndarray:
    """Sum a small f64 vector over the ranks of the default (or given) process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.array(vec, dtype=np.float64)
    on_gpu = dist.get_backend(group) == "nccl"
    t = torch.from_numpy(np.array(vec, dtype=np.float64))
    if on_gpu:
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def summarize_aggregate(agg: np.ndarray, line: abi.HostLine) -> dict:
    """Mean and standard error of every metric from the reduced moment vector
    ``[count, (sum x, sum x^2) for TP, residence, blocked[M], starved[M], occupancy[M-1]]``."""
    n = float(agg[0])
    names = ["TP", "residence"] + [f"blocked[{i}]" for i in range(line.M)] + \
            [f"starved[{i}]" for i in range(line.M)] + [f"occupancy[{b}]" for b in range(line.M - 1)]
    out = {"count": int(n)}
    for j, name in enumerate(names):
        s1, s2 = float(agg[1 + 2 * j]), float(agg[2 + 2 * j])
        mean = s1 / n if n else math.nan
        var = max(0.0, (s2 - n * mean * mean) / (n - 1)) if n > 1 else math.nan
        out[name] = {"mean": mean, "stderr": math.sqrt(var / n) if n > 1 else math.nan}
    return out


class ShardResult:
    def __init__(self, lo, hi, local, aggregate, summary):
        self.lo, self.hi, self.local, self.aggregate, self.summary = lo, hi, local, aggregate, summary


def run_sharded(line: abi.HostLine, n_total: int, mode: int = abi.RNG_COUNTER, seed: int = 0,
                first_instance_id: int = 0, runner=None, group=None, rank: int | None = None,
                world: int | None = None, device=None) -> ShardResult:
    """Run this rank's range of the batch and reduce the aggregate moments over all ranks.

    ``runner(line_shard, n_local, mode, seed, first_id) -> abi.HostStats`` defaults to the CUDA path on
    ``cuda:<local rank>``; the CPU tests inject the oracle to exercise the sharding and the collective.
    """
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, world, rank)
    local = None
    agg = np.zeros(abi.aggregate_len(line.M))
    if hi > lo:
        shard = slice_line(line, lo, hi)
        if runner is None:
            import torch

            from . import engine
            dev = torch.cuda.current_device() if device is None else device
            dev = torch.device(dev).index if not isinstance(dev, int) else dev
            device = torch.device("cuda", int(dev))
            with engine.Handle(device=int(dev)) as h:
                local = h.run(shard, hi - lo, mode=mode, seed=seed, first_instance_id=first_instance_id + lo).stats
        else:
            local = runner(shard, hi - lo, mode, seed, first_instance_id + lo)
        agg = np.array(local.aggregate, dtype=np.float64)
    total = all_reduce_sum(agg, group=group, device=device)
    return ShardResult(lo, hi, local, total, summarize_aggregate(total, line))
