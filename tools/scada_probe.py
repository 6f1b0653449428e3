#!/usr/bin/env python
"""One launch of the C4 line (20 machines, mixed BAS/BBS, 8 job types) with SCADA capture, for ncu:

    python tools/scada_probe.py [--instances N] [--warm-up T0] [--run-to T1] [--no-log]

Prints events, records and the CUDA-event time of the launch."""
import argparse, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lines import SEED, mixed20_line
from mjpdes_b200 import abi, engine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=75776)
    ap.add_argument("--warm-up", type=float, default=100.0)
    ap.add_argument("--run-to", type=float, default=400.0)
    ap.add_argument("--no-log", action="store_true")
    ap.add_argument("--launches", type=int, default=2)
    a = ap.parse_args()
    line = mixed20_line(warm_up=a.warm_up, run_to=a.run_to, log=not a.no_log)
    with engine.Handle(device=0) as h:
        h.upload(line, a.instances)
        if not a.no_log:
            h.reserve_log(int(a.instances * 45 * (a.run_to - a.warm_up)))
        for r in range(a.launches):
            h.launch(abi.RNG_COUNTER, SEED, r * a.instances, want_log=not a.no_log)
            h.sync()
        st = h.download_fields(("n_events", "status"))
        rec = h.log_count() if not a.no_log else (0, 0)
        print(json.dumps(dict(instances=a.instances, events=int(st.n_events.sum()), records=rec[0], dropped=rec[1],
                              ms=h.kernel_ms(), all_normal_end=bool(np.all(st.status == abi.INST_NORMAL_END)))))


if __name__ == "__main__":
    main()
