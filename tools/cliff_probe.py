#!/usr/bin/env python
"""Throughput of the headline line as a function of the batch size around one wave of CTAs (104 192 lanes):

    python tools/cliff_probe.py [--run-to 5000] [--sizes 90000,100000,...] [--ragged]

Run once with MJP_ROTATE=0 (every CTA runs its instances from start to end: a batch just over one wave pays for a
second, nearly empty wave) and once with the default (rotation, mjp_line_kernel.cuh).  --ragged: a mixed batch in
which half of the instances stop early on :run-to-job."""
import argparse, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from mjpdes_b200 import abi, engine, workloads as wl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--run-to", type=float, default=5000.0)
    ap.add_argument("--sizes", default="90000,100000,104192,106000,110000,120000,130000,160000,208384")
    a = ap.parse_args()
    out = {"rotate_env": os.environ.get("MJP_ROTATE", "default"), "run_to": a.run_to, "sizes": {}}
    with engine.Handle(device=0) as h:
        for n in (int(x) for x in a.sizes.split(",")):
            line = wl.example_line(warm_up=a.run_to / 10, run_to=a.run_to)
            h.upload(line, n)
            best = None
            for r in range(3):
                h.launch(abi.RNG_COUNTER, wl.SEED, r * n); h.sync()
                if r:
                    best = h.kernel_ms() if best is None else min(best, h.kernel_ms())
            st = h.download_fields(("n_events", "status"))
            ev = float(st.n_events.sum())
            out["sizes"][n] = {"ms": best, "events_per_sec": ev / (best / 1e3), "all_normal_end": bool(np.all(st.status == 0))}
            print(n, out["sizes"][n], file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
