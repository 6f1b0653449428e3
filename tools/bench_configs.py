#!/usr/bin/env python
"""Throughput of the named synthetic configurations of BASELINE.json / SURVEY.md 8(d) other than the
headline C2 (that one is bench.py): slices sized to run in seconds on ONE B200.

    python tools/bench_configs.py [--out profiles/xyz.json]

C3  10-machine sweep (per-instance W / N planes), a 104192-point slice of the 1M-point grid
C4  20-machine mixed BAS/BBS, 8 job types, SCADA capture on (record stream written to HBM)
C5  50-machine, 16 job types: a time-boxed slice (run-to 2e4 instead of 1e6), aggregate only
"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lines import SEED, sweep10_line, mixed20_line, long50_line, example2_line, submodel1_line
from mjpdes_b200 import abi, engine


def timed(h, line, n, log_capacity=0, reps=2):
    h.upload(line, n)
    if log_capacity:
        h.reserve_log(log_capacity)
    best = None
    for r in range(reps + 1):
        h.launch(abi.RNG_COUNTER, SEED, r * n, want_log=bool(log_capacity))
        h.sync()
        ms = h.kernel_ms()
        if r:                                  # first launch is the warm-up
            best = ms if best is None else min(best, ms)
    res = h.download(log_capacity=log_capacity)
    ev = int(res.stats.n_events.sum())
    ok = bool(np.all(res.stats.status == abi.INST_NORMAL_END))
    return dict(instances=n, events=ev, ms=best, events_per_sec=ev / (best / 1e3), all_normal_end=ok,
                records=int(len(res.log)) if res.log is not None else None,
                log_rc=int(res.rc))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    a = ap.parse_args()
    out = {}
    with engine.Handle(device=0) as h:
        out["C3_sweep10"] = timed(h, sweep10_line(104192, first=0), 104192)
        out["C3_sweep10"]["note"] = "M=10, K=1, W/N per-instance planes, warm-up 2000 run-to 20000, 104192 of the 1e6 grid points"
        n4 = 9472
        cap = int(n4 * 40_000)
        out["C4_mixed20_scada"] = timed(h, mixed20_line(warm_up=2000, run_to=3000), n4, log_capacity=cap)
        r = out["C4_mixed20_scada"]
        r["record_bytes_per_sec"] = 32.0 * r["records"] / (r["ms"] / 1e3)
        r["note"] = "M=20 mixed BAS/BBS, K=8, :log? true :up&down? true :max-lines Inf; horizon 2000..3000 so the stream fits HBM"
        out["C4_mixed20_stats_only"] = timed(h, mixed20_line(warm_up=2000, run_to=20000, log=False), 16384)
        out["C5_long50_slice"] = timed(h, long50_line(warm_up=2000, run_to=20000), 9472)
        out["C5_long50_slice"]["note"] = "M=50, K=16; time-boxed slice (run-to 2e4, not 1e6), 9472 = 148 x 64 instances"
        out["C4_mixed20_stats_many"] = timed(h, mixed20_line(warm_up=2000, run_to=6000, log=False), 75776, reps=1)
        out["C5_long50_many"] = timed(h, long50_line(warm_up=2000, run_to=5000), 37888, reps=1)
        out["C5_long50_many"]["note"] = "M=50, K=16; 37888 instances, run-to 5000: enough instances to need more than one wave without the COLD layout"
        out["example2"] = timed(h, example2_line(), 104192)
        out["submodel1"] = timed(h, submodel1_line(), 104192)
    print(json.dumps(out, indent=1))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
