"""Smallest end-to-end case for compute-sanitizer: a few instances of every kernel variant family."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from lines import SEED, example_line, mixed20_line, long50_line, sweep10_line
from mjpdes_b200 import abi, engine
with engine.Handle(device=0, event_hash=True) as h:
    for line, n, cap in ((example_line(run_to=150, warm_up=15), 70, 0),
                         (example_line(run_to=150, warm_up=15, log=True, up_down=True, max_lines=abi.UINT64_MAX), 70, 200000),
                         (sweep10_line(70, warm_up=10, run_to=80), 70, 0), (mixed20_line(warm_up=10, run_to=60), 40, 200000),
                         (long50_line(warm_up=10, run_to=40), 33, 0)):
        for mode in (abi.RNG_COUNTER, abi.RNG_REPLAY):
            r = h.run(line, n, mode=mode, seed=SEED, log_capacity=cap)
            print(line.M, mode, r.total_events, set(r.stats.status.tolist()))
print("sanitize smoke done")
