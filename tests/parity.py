"""Shared comparison helpers: a result of the CUDA path (or of the host-shimmed kernel source)
against the oracle on the same seeded inputs.  Integer fields and the event hash must be equal;
f64 accumulators are compared bit-for-bit and, failing that, at the 1e-9 relative tolerance the
north star states (the failure message says which)."""
import numpy as np

INT_FIELDS = ["status", "n_events", "current_job", "njobs", "event_hash"]
F64_FIELDS = ["final_clock", "residence_sum", "blocked_time", "starved_time", "occ_time"]
RTOL = 1e-9


def assert_same(got, want, hash_checked=True, bit_exact=True):
    for f in INT_FIELDS:
        if f == "event_hash" and not hash_checked:
            continue
        x, y = getattr(got, f), getattr(want, f)
        if not np.array_equal(x, y):
            bad = np.argwhere(np.asarray(x) != np.asarray(y))
            raise AssertionError(f"{f}: {len(bad)} instances differ, first at {bad[0]}: "
                                 f"got {x[tuple(bad[0])]} want {y[tuple(bad[0])]}")
    for f in F64_FIELDS:
        x, y = getattr(got, f), getattr(want, f)
        if np.array_equal(x, y):
            continue
        if not np.allclose(x, y, rtol=RTOL, atol=0.0):
            bad = np.argwhere(~np.isclose(x, y, rtol=RTOL, atol=0.0))
            raise AssertionError(f"{f}: {len(bad)} values outside rtol {RTOL}, first at {bad[0]}: "
                                 f"got {x[tuple(bad[0])]!r} want {y[tuple(bad[0])]!r}")
        if bit_exact:
            raise AssertionError(f"{f}: within {RTOL} but not bit-identical")
