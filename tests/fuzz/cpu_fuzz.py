#!/usr/bin/env python
"""Differential fuzzing on the CPU: the kernel SOURCE (compiled as host C++ through tests/hostsim) against the
oracle, on random lines -- statistics, event hash and SCADA records must be identical.

    python tests/fuzz/cpu_fuzz.py [seconds] [first_trial] [wild]

Trial t is fully determined by t (line shape, RNG mode incl. tapes, COLD layout, slice boundaries, SCADA capture), so a
failure is reproduced by `python tests/fuzz/cpu_fuzz.py 1 <t>`.  Companion of tests/fuzz/gpu_fuzz.py (same idea through the
C ABI on a GPU).  Test tooling: imports oracle/ and tests/."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mjpdes_b200 import abi
from oracle import binding as O
from hostsim import binding as H
from lines import random_line
from parity import assert_same


def wild_line(rng, M, K, run_to, warm_up, n, **kw):
    """Lines outside the comfortable range of tests/lines.py::random_line: deep buffers, unequal portions (the
    shifted thresholds of new-job, core.clj:332-339, can then yield no job type at all), failure-prone or slow-to-
    repair machines, wide spreads of processing times, per-instance override planes."""
    lam = np.where(rng.random(M) < 0.1, 0.0, rng.choice([0.005, 0.05, 0.3, 1.0, 3.0], M) * rng.uniform(0.5, 1.5, M))
    mu = rng.choice([0.05, 0.3, 1.0, 5.0, 50.0], M) * rng.uniform(0.5, 1.5, M)
    W = rng.choice([0.2, 1.0, 1.0, 3.0], M) * rng.uniform(0.8, 1.2, M)
    N = rng.choice([1, 1, 2, 3, 8, 40, 254], M - 1)
    portion = rng.dirichlet(np.ones(K)) if rng.random() < 0.5 else np.full(K, 1.0 / K)
    portion = np.maximum(portion, 1e-3)
    w = rng.choice([0.05, 0.5, 1.0, 1.0, 2.0, 7.5], size=(K, M))
    planes = {}
    if rng.random() < 0.3:
        planes = dict(W_inst=rng.uniform(0.5, 2.0, (M, n)), N_inst=rng.integers(1, N[:, None] + 1, (M - 1, n)).astype(np.int32))
        if rng.random() < 0.5:
            planes.update(lam_inst=rng.uniform(0.0, 0.5, (M, n)), mu_inst=rng.uniform(0.2, 3.0, (M, n)),
                          w_inst=rng.uniform(0.3, 3.0, (K * M, n)))
    return abi.HostLine(lam=lam, mu=mu, W=W, discipline=rng.random(M) < rng.choice([0, .3, 1.0]), N=N, portion=portion, w=w,
                        warm_up=warm_up, run_to=run_to, **planes, **kw)


WILD = False


def one(trial: int):
    rng = np.random.default_rng(1_700_000 + trial)
    M = int(rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 20, 31, 32, 33, 40, 64]))
    K = int(rng.integers(1, 6)) if not WILD else int(rng.choice([1, 2, 3, 7, 16, 40]))
    kw = dict(jm2dm=bool(rng.random() < .3), up_down=bool(rng.random() < .4),
              bbs_prob=float(rng.choice([0, .3, .7, 1.0])), reliable_prob=float(rng.choice([0, .1, .5])))
    run_to = float(rng.choice([60, 150, 400]))
    if WILD:
        n_w = int(rng.integers(1, 4))
        line = wild_line(rng, M, K, run_to, run_to * float(rng.choice([0, 0.1, 0.5])), n_w,
                         jm2dm=kw["jm2dm"], up_down=kw["up_down"])
    else:
        line = random_line(rng, M, K, run_to=run_to, warm_up=run_to * float(rng.choice([0, 0.1, 0.5])), **kw)
    if rng.random() < 0.3:                       # lattice-time line: identical machines -> many ties
        line.W[:] = 1.0
        line.w[:] = 1.0
    if rng.random() < 0.2:
        line.run_to_job = int(rng.integers(10, 200))
    mode, n = int(rng.integers(0, 2)), int(rng.integers(1, 4))
    if WILD:
        n = n_w
    cold = bool(rng.random() < 0.4)
    slices = sorted(rng.uniform(0, run_to * 1.2, int(rng.integers(0, 4))).tolist()) if rng.random() < 0.4 else []
    log = rng.random() < 0.3
    if log:
        line.log = True
        line.max_lines = int(rng.choice([abi.UINT64_MAX, 50, 500]))
    cap = 300_000 if log else 0
    if rng.random() < 0.1:                       # MJP_RNG_TAPE: replay recorded draws, sometimes from tapes cut short
        tapes, _ = O.record_tapes(line, n, trial, 1024, 256, 1024)
        if rng.random() < 0.5:
            tapes = abi.DrawTapes(tapes.jobtype[:, :int(rng.integers(1, 200))], tapes.uniform[:, :, :int(rng.integers(1, 64))],
                                  tapes.exp[:, :, :int(rng.integers(1, 200))])
        want = O.run(line, n, mode=abi.RNG_TAPE, tapes=tapes, log_capacity=cap)
        got = H.run(line, n, mode=abi.RNG_TAPE, tapes=tapes, cold=cold, slices=slices, log_capacity=cap)
    else:
        want = O.run(line, n, mode=mode, seed=trial, log_capacity=cap)
        got = H.run(line, n, mode=mode, seed=trial, cold=cold, slices=slices, log_capacity=cap)
    if np.any(want.stats.status == abi.INST_TAPE_EXHAUSTED):
        # The oracle realises up/down events ahead of time, as the reference's end-time does; the kernel draws
        # them lazily.  On a tape cut short the oracle therefore runs out first: the kernel may stop later with
        # the same status, or not need the missing draws at all.  Instances the oracle completes must agree.
        ok = (want.stats.status == abi.INST_TAPE_EXHAUSTED) | (np.asarray(got.status) == want.stats.status)
        assert np.all(ok), "status"
        assert np.all(np.asarray(got.n_events) >= want.stats.n_events), "the kernel stopped before the oracle"
        return M, K, mode, cold, slices, log
    assert_same(got, want.stats)
    if log:
        a, b = np.sort(got.log, order=["instance", "line"]), np.sort(want.log, order=["instance", "line"])
        assert len(a) == len(b)
        assert all(np.array_equal(a[f], b[f]) for f in ("clk", "job", "line", "n", "ord", "act", "machine"))
        assert np.array_equal(a["aux"].view(np.uint64), b["aux"].view(np.uint64))
        assert np.array_equal(got.first_state, want.first_state), "first_state"
    return M, K, mode, cold, slices, log


def main():
    global WILD
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
    trial = first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    WILD = len(sys.argv) > 3 and sys.argv[3] == "wild"
    t0, ok, fails = time.time(), 0, []
    while time.time() - t0 < budget:
        try:
            one(trial)
            ok += 1
        except Exception as e:                   # noqa: BLE001 -- every kind of disagreement is a finding
            fails.append((trial, str(e)[:200]))
            print("FAIL", fails[-1], flush=True)
        trial += 1
    print(f"cpu fuzz: trials {first}..{trial - 1} ok {ok} fails {len(fails)}")
    print(fails[:10])
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
