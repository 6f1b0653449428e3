#!/usr/bin/env python
"""Differential fuzz on a GPU: random lines through the C ABI vs the CPU oracle (bit-exact compare of status,
event hash, statistics and SCADA records).   python tests/fuzz/gpu_fuzz.py [seconds] [first_trial] [wild]
`wild` draws the lines from cpu_fuzz.wild_line (deep buffers, unequal portions, extreme rates, override planes)."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lines import random_line
from mjpdes_b200 import abi, engine
from oracle import binding as O
from parity import assert_same
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cpu_fuzz import wild_line

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
trial = int(sys.argv[2]) if len(sys.argv) > 2 else 0
wild = len(sys.argv) > 3 and sys.argv[3] == "wild"
t0, ok, fails = time.time(), 0, []
with engine.Handle(device=0, event_hash=True) as h:
    while time.time() - t0 < budget:
        rng = np.random.default_rng(500000 + trial)
        M = int(rng.choice([2, 3, 4, 5, 5, 5, 6, 7, 8, 9, 10, 12, 16, 20, 31, 32, 33, 40, 64]))
        K = int(rng.integers(1, 6))
        kw = dict(jm2dm=bool(rng.random() < .3), up_down=bool(rng.random() < .4), bbs_prob=float(rng.choice([0, .3, .7, 1.0])),
                  reliable_prob=float(rng.choice([0, .1, .5])))
        run_to = float(rng.choice([60, 150, 400]))
        n = int(rng.choice([1, 31, 33, 64, 97, 200]))
        if wild:
            K = int(rng.choice([1, 2, 3, 7, 16, 40]))
            line = wild_line(rng, M, K, run_to, run_to * float(rng.choice([0, 0.1, 0.5])), n, jm2dm=kw["jm2dm"],
                             up_down=kw["up_down"])
        else:
            line = random_line(rng, M, K, run_to=run_to, warm_up=run_to * float(rng.choice([0, 0.1, 0.5])), **kw)
        if rng.random() < 0.3:
            line.W[:] = 1.0; line.w[:] = 1.0
        if rng.random() < 0.2:
            line.run_to_job = int(rng.integers(10, 200))
        mode = int(rng.integers(0, 2))
        log = rng.random() < 0.3
        slices = sorted(rng.uniform(0, run_to * 1.2, int(rng.integers(1, 4))).tolist()) if rng.random() < 0.3 else []
        cap = 0
        if log:
            line.log = True; line.max_lines = int(rng.choice([abi.UINT64_MAX, 50, 500]))
            cap = int(n * (12 * M * run_to + 1000)) if line.max_lines == abi.UINT64_MAX else n * (line.max_lines + 10 * M)
            cap = min(cap, 20_000_000)
        try:
            want = O.run(line, n, mode=mode, seed=trial, log_capacity=cap, n_threads=8)
            if slices and not log:
                h.upload(line, n)
                resume = False
                for until in slices + [float("inf")]:
                    h.launch_slice(until, resume, mode=mode, seed=trial); h.sync(); resume = True
                got = h.download()
            else:
                got = h.run(line, n, mode=mode, seed=trial, log_capacity=cap)
            assert_same(got.stats, want.stats)
            if log:
                assert got.rc == want.rc, (got.rc, want.rc)
                if got.rc == abi.OK:
                    a = np.sort(got.log, order=["instance", "line"]); b = np.sort(want.log, order=["instance", "line"])
                    assert len(a) == len(b)
                    for f in ("clk", "job", "line", "n", "ord", "act", "machine"):
                        assert np.array_equal(a[f], b[f]), f
                    assert np.array_equal(a["aux"].view(np.uint64), b["aux"].view(np.uint64))
                    assert np.array_equal(got.first_state, want.first_state), "first_state"
            ok += 1
        except Exception as e:
            fails.append((trial, M, K, mode, n, log, slices, str(e)[:160]))
            print("FAIL", fails[-1], flush=True)
        trial += 1
print(f"gpu fuzz: trials {trial} ok {ok} fails {len(fails)}")
sys.exit(1 if fails else 0)
