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
from lines import SEED, example_line, sweep10_line, mixed20_line, long50_line, example2_line, submodel1_line
from mjpdes_b200 import abi, engine


def timed(h, line, n, log_capacity=0, reps=2, keep_log_on_device=False):
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
    if keep_log_on_device:                     # count the records; a stream of tens of GB stays in HBM
        st = h.download_fields(("n_events", "status"))
        records, dropped = h.log_count()
        ev, ok = int(st.n_events.sum()), bool(np.all(st.status == abi.INST_NORMAL_END))
        return dict(instances=n, events=ev, ms=best, events_per_sec=ev / (best / 1e3), all_normal_end=ok,
                    records=records, dropped=dropped)
    res = h.download(log_capacity=log_capacity)
    ev = int(res.stats.n_events.sum())
    ok = bool(np.all(res.stats.status == abi.INST_NORMAL_END))
    return dict(instances=n, events=ev, ms=best, events_per_sec=ev / (best / 1e3), all_normal_end=ok,
                records=int(len(res.log)) if res.log is not None else None,
                log_rc=int(res.rc))


def scada(h, n, reps, on_device, line=None, per_instance=40_000):
    r = timed(h, line or mixed20_line(warm_up=2000, run_to=3000), n, log_capacity=int(n * per_instance), reps=reps,
              keep_log_on_device=on_device)
    r["record_bytes_per_sec"] = 32.0 * r["records"] / (r["ms"] / 1e3)
    # the warm-up (0..2000) is simulated without capture; the same line without :log? gives its cost
    return r


CONFIGS = {
    "C3_sweep10": (lambda h: timed(h, sweep10_line(104192, first=0), 104192),
                   "M=10, K=1, W/N per-instance planes, warm-up 2000 run-to 20000, 104192 of the 1e6 grid points"),
    "C4_mixed20_scada": (lambda h: scada(h, 9472, 2, False),
                         "M=20 mixed BAS/BBS, K=8, :log? true :up&down? true :max-lines Inf; horizon 2000..3000 so the "
                         "stream fits HBM"),
    "C4_mixed20_scada_many": (lambda h: scada(h, 75776, 1, True),
                              "same, 75776 instances: 2.6e9 records = 83 GB of HBM, counted on the device, not copied "
                              "to the host"),
    "example_scada": (lambda h: scada(h, 104192, 1, True, per_instance=12_500,
                                      line=example_line(warm_up=2000, run_to=3000, log=True, up_down=True,
                                                        max_lines=abi.UINT64_MAX)),
                      "resources/example.clj with :log? true :up&down? true, horizon 2000..3000, 104192 instances"),
    "example_warmup_only": (lambda h: timed(h, example_line(warm_up=2000, run_to=2000.001, up_down=True), 104192, reps=1),
                            "the capture-free part (0..2000) of example_scada"),
    "C4_mixed20_warmup_only_many": (lambda h: timed(h, mixed20_line(warm_up=2000, run_to=2000.001, log=False), 75776,
                                                    reps=1),
                                    "the capture-free part (0..2000) of C4_mixed20_scada_many, to separate the two"),
    "C4_mixed20_stats_only": (lambda h: timed(h, mixed20_line(warm_up=2000, run_to=20000, log=False), 16384), ""),
    "C5_long50_slice": (lambda h: timed(h, long50_line(warm_up=2000, run_to=20000), 9472),
                        "M=50, K=16; time-boxed slice (run-to 2e4, not 1e6), 9472 = 148 x 64 instances"),
    "C4_mixed20_stats_many": (lambda h: timed(h, mixed20_line(warm_up=2000, run_to=6000, log=False), 75776, reps=1), ""),
    "C5_long50_many": (lambda h: timed(h, long50_line(warm_up=2000, run_to=5000), 37888, reps=1),
                       "M=50, K=16; 37888 instances, run-to 5000: enough instances to need more than one wave without "
                       "the COLD layout"),
    "example2": (lambda h: timed(h, example2_line(), 104192), ""),
    "submodel1": (lambda h: timed(h, submodel1_line(), 104192), ""),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--only", default="", help="comma-separated names of the configurations to run (default: all)")
    a = ap.parse_args()
    want = [w for w in a.only.split(",") if w]
    out = {}
    with engine.Handle(device=0) as h:
        for name, (fn, note) in CONFIGS.items():
            if want and name not in want:
                continue
            out[name] = fn(h)
            if note:
                out[name]["note"] = note
    print(json.dumps(out, indent=1))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
