#!/usr/bin/env python
"""One launch of a named workload at a profiling-sized horizon, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:mjp_line_kernel --launch-skip 1 --launch-count 1 \\
        -o gpurun_out/<name> python tools/profile_config.py <name> --out gpurun_out/<name>_events.json

The script launches the variant once on a small batch (warm-up: launch #0, skipped by ncu) and once on the
profiled batch (launch #1).  It writes the number of events the profiled launch executed, which
tools/issue_model.py divides the instruction count by.  Horizons are shortened so that ~40 ncu replays finish in
seconds; ratios (warp-instructions per event, lanes per instruction, stall mix) carry over, absolute times do not.
"""
import argparse, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from mjpdes_b200 import abi, engine, workloads as wl

CONFIGS = {
    # name: (line factory(n), n, log records per instance, sliced quantum (q0, q1) or None)
    "C2_example": (lambda n: wl.example_line(warm_up=200.0, run_to=2000.0), 104192, 0, None),
    "C2_small": (lambda n: wl.example_line(warm_up=500.0, run_to=5000.0), 10000, 0, None),
    "C2_example_full": (lambda n: wl.example_line(), 100000, 0, None),
    "C3_sweep10": (lambda n: wl.sweep10_line(n, warm_up=100.0, run_to=1000.0, first=0), 104192, 0, None),
    "C4_mixed20_stats": (lambda n: wl.mixed20_line(warm_up=100.0, run_to=1000.0, log=False), 10000, 0, None),
    "C4_mixed20_stats_many": (lambda n: wl.mixed20_line(warm_up=100.0, run_to=600.0, log=False), 75776, 0, None),
    "C4_mixed20_scada": (lambda n: wl.mixed20_line(warm_up=100.0, run_to=700.0), 10000, 22000, None),
    "C4_mixed20_scada_many": (lambda n: wl.mixed20_line(warm_up=100.0, run_to=400.0), 75776, 11000, None),
    "example_scada": (lambda n: wl.example_line(warm_up=100.0, run_to=400.0, log=True, up_down=True,
                                                max_lines=abi.UINT64_MAX), 104192, 4000, None),
    "C5_long50_slice": (lambda n: wl.long50_line(warm_up=1e5, run_to=1e6), 75776, 0, (20.0, 120.0)),
    "C5_long50": (lambda n: wl.long50_line(warm_up=20.0, run_to=200.0), 37888, 0, None),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", choices=sorted(CONFIGS))
    ap.add_argument("--out")
    ap.add_argument("--instances", type=int, default=0)
    a = ap.parse_args()
    mk, n, per_log, quantum = CONFIGS[a.name]
    n = a.instances or n
    with engine.Handle(device=0) as h:
        ev0 = 0
        if quantum:
            h.upload(mk(n), n)
            h.launch_slice(quantum[0], False, seed=wl.SEED); h.sync()          # launch #0 (start-up transient)
            ev0 = int(h.download_fields(("n_events",)).n_events.sum())
            h.launch_slice(quantum[1], True, seed=wl.SEED); h.sync()           # launch #1: the profiled quantum
        else:
            small = min(n, 2048)
            h.upload(mk(small), small)
            if per_log:
                h.reserve_log(small * per_log)
            h.launch(abi.RNG_COUNTER, wl.SEED, 0, want_log=bool(per_log)); h.sync()   # launch #0
            h.upload(mk(n), n)
            if per_log:
                h.reserve_log(n * per_log)
            h.launch(abi.RNG_COUNTER, wl.SEED, 0, want_log=bool(per_log)); h.sync()   # launch #1: profiled
        st = h.download_fields(("n_events", "status"))
        res = {"name": a.name, "instances": n, "events": int(st.n_events.sum()) - ev0, "ms": h.kernel_ms(),
               "status_counts": {str(k): int(v) for k, v in zip(*np.unique(st.status, return_counts=True))}}
        if per_log:
            res["records"], res["dropped"] = h.log_count()
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh)


if __name__ == "__main__":
    main()
