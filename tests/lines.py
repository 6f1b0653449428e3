"""Line descriptors used across the tests: the reference's own models and the
synthetic configs C1-C5 of BASELINE.json / SURVEY.md 8(d)."""
from __future__ import annotations

import json
import os

import numpy as np

from mjpdes_b200 import abi
from mjpdes_b200.workloads import (SEED, example2_line, example_line, long50_line, mixed20_line,  # noqa: F401
                                    submodel1_line, sweep10_line)

HERE = os.path.dirname(os.path.abspath(__file__))

with open(os.path.join(HERE, "golden", "reference_goldens.json")) as _fh:
    GOLDENS = json.load(_fh)


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
