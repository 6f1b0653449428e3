"""Line descriptors used across the tests: the reference's own models and the
synthetic configs C1-C5 of BASELINE.json / SURVEY.md 8(d)."""
from __future__ import annotations

import json
import os

import numpy as np

from mjpdes_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x4D4A50646573  # "MJPdes"

with open(os.path.join(HERE, "golden", "reference_goldens.json")) as _fh:
    GOLDENS = json.load(_fh)


def example_line(warm_up=2000.0, run_to=20000.0, **kw) -> abi.HostLine:
    """resources/example.clj:2-19 (5 machines, 4 buffers, 2 job types)."""
    return abi.HostLine(lam=[0.1] * 5, mu=[0.9] * 5, W=[1.2, 1.0, 1.1, 1.05, 1.2], discipline=[0] * 5,
                        N=[3, 5, 1, 1], portion=[0.8, 0.2],
                        w=[[1.0, 1.0, 1.0, 1.0, 1.0], [1.0, 2.0, 1.5, 1.0, 1.0]],
                        warm_up=warm_up, run_to=run_to, **kw)


def example2_line(warm_up=2000.0, run_to=20000.0, **kw) -> abi.HostLine:
    """resources/example2.clj (single job type)."""
    return abi.HostLine(lam=[0.1] * 5, mu=[0.9] * 5, W=[1.2, 1.0, 1.1, 1.05, 1.2], discipline=[0] * 5,
                        N=[3, 5, 1, 1], portion=[1.0], w=[[1.0] * 5], warm_up=warm_up, run_to=run_to, **kw)


def submodel1_line(warm_up=2000.0, run_to=20000.0, **kw) -> abi.HostLine:
    """data/submodel-1.clj (2 machines, N=3)."""
    return abi.HostLine(lam=[0.1] * 2, mu=[0.9] * 2, W=[1.0, 1.0], discipline=[0, 0], N=[3], portion=[1.0],
                        w=[[1.0, 1.0]], warm_up=warm_up, run_to=run_to, **kw)


def golden_two_machine(gid: str, lam=None, **kw) -> abi.HostLine:
    """The 2-machine models of G7 (load-m1), G8 (BAS) and G9 (BBS), from the golden file."""
    m = GOLDENS[gid]["model"]
    names = [x for x in m[":topology"] if ":N" not in m[":line"][x]]
    bufs = [x for x in m[":topology"] if ":N" in m[":line"][x]]
    jt = list(m[":jobmix"].values())
    line = m[":line"]
    rep = m.get(":report", {})
    args = dict(
        lam=[line[x][":lambda"] if lam is None else lam for x in names],
        mu=[line[x][":mu"] for x in names], W=[line[x][":W"] for x in names],
        discipline=[1 if line[x].get(":discipline") == ":BBS" else 0 for x in names],
        N=[line[b][":N"] for b in bufs], portion=[j[":portion"] for j in jt],
        w=[[j[":w"][x] for x in names] for j in jt],
        warm_up=m[":params"][":warm-up-time"], run_to=m[":params"][":run-to-time"],
        log=True, up_down=bool(rep.get(":up&down?", False)), max_lines=abi.UINT64_MAX)
    args.update(kw)
    return abi.HostLine(**args)


def sweep10_line(n_inst: int, warm_up=2000.0, run_to=20000.0, first=0, total=1_000_000) -> abi.HostLine:
    """C3: 10-machine line, 4-axis grid 10 x 10 x 100 x 100 over (N_all, N_b5, W_all, W_m5);
    instance id = point index, this call holds points [first, first+n_inst)."""
    M = 10
    idx = np.arange(first, first + n_inst, dtype=np.int64)
    i_w5 = idx % 100
    i_w = (idx // 100) % 100
    i_n5 = (idx // 10_000) % 10
    i_n = (idx // 100_000) % 10
    grid = np.linspace(0.8, 1.2, 100)
    W = np.tile(grid[i_w], (M, 1))
    W[4] = grid[i_w5]
    N = np.tile((1 + i_n).astype(np.int32), (M - 1, 1))
    N[4] = (1 + i_n5).astype(np.int32)
    w = np.ones((1, M))
    w[0, M - 1] = 1.01
    return abi.HostLine(lam=[0.1] * M, mu=[0.9] * M, W=[1.0] * M, discipline=[0] * M, N=[10] * (M - 1),
                        portion=[1.0], w=w, warm_up=warm_up, run_to=run_to, W_inst=W, N_inst=N)


def mixed20_line(warm_up=2000.0, run_to=20000.0, log=True, up_down=True, max_lines=abi.UINT64_MAX) -> abi.HostLine:
    """C4: 20 machines, BBS on the even-numbered names (m2, m4, ... = odd indices), 8 job types, SCADA on."""
    M, K = 20, 8
    i = np.arange(M)
    return abi.HostLine(lam=[0.1] * M, mu=[0.9] * M, W=1.0 + 0.05 * ((7 * i) % 5), discipline=(i % 2 == 1),
                        N=1 + (np.arange(M - 1) % 5), portion=[0.125] * K,
                        w=[[1.0 + 0.25 * ((k + j) % 4) for j in range(M)] for k in range(K)],
                        warm_up=warm_up, run_to=run_to, log=log, up_down=up_down, max_lines=max_lines)


def long50_line(warm_up=1e5, run_to=1e6) -> abi.HostLine:
    """C5: 50 machines, 16 job types."""
    M, K = 50, 16
    i = np.arange(M)
    return abi.HostLine(lam=[0.1] * M, mu=[0.9] * M, W=1.0 + 0.02 * ((11 * i) % 11), discipline=[0] * M,
                        N=1 + (np.arange(M - 1) % 4), portion=[1.0 / K] * K,
                        w=[[1.0 + 0.125 * ((3 * k + 5 * j) % 8) for j in range(M)] for k in range(K)],
                        warm_up=warm_up, run_to=run_to)


def random_line(rng: np.random.Generator, M: int, K: int, run_to=300.0, warm_up=30.0, bbs_prob=0.3,
                reliable_prob=0.1, **kw) -> abi.HostLine:
    """A random but valid line for property/parity sweeps."""
    lam = np.where(rng.random(M) < reliable_prob, 0.0, rng.uniform(0.02, 0.3, M))
    mu = rng.uniform(0.5, 1.5, M)
    W = rng.uniform(0.8, 1.2, M)
    disc = rng.random(M) < bbs_prob
    N = rng.integers(1, 6, M - 1)
    # equal portions keep the shifted thresholds of new-job (core.clj:332-339) inside [0,1]
    portion = np.full(K, 1.0 / K)
    w = rng.choice([0.5, 0.75, 1.0, 1.0, 1.25, 1.5, 2.0], size=(K, M))
    return abi.HostLine(lam=lam, mu=mu, W=W, discipline=disc, N=N, portion=portion, w=w, warm_up=warm_up,
                        run_to=run_to, **kw)


PRETTY = {abi.ACT_AJ: "start-job", abi.ACT_SM: "start-job", abi.ACT_BJ: "complete-job", abi.ACT_EJ: "complete-job",
          abi.ACT_BL: "blocked", abi.ACT_UB: "unblocked", abi.ACT_ST: "starved", abi.ACT_US: "unstarved",
          abi.ACT_UP: "up", abi.ACT_DOWN: "down"}


def pretty_act(rec) -> str:
    """mjp2pretty-name, util/log.clj:248-261."""
    return f":m{int(rec['machine']) + 1}-{PRETTY[int(rec['act'])]}"
