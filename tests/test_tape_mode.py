"""MJP_RNG_TAPE: every random draw is read from per-consumer tapes -- the mechanism by which the
reference's OWN stream can be replayed (record what clojure.core/rand and incanter's sample-exp return in a
JVM run, feed the tapes here).  No JVM exists in this image, so the recording side is played by the oracle
in REPLAY mode; a TAPE run of the recorded draws must then reproduce that run exactly."""
import numpy as np
import pytest

from lines import SEED, example_line, random_line
from mjpdes_b200 import abi
from parity import assert_same


def _cases():
    yield example_line(run_to=400.0, warm_up=40.0), 3
    yield random_line(np.random.default_rng(3), 7, 3, run_to=300.0, warm_up=30.0, bbs_prob=0.5, jm2dm=True), 2
    yield random_line(np.random.default_rng(4), 36, 2, run_to=120.0, warm_up=10.0), 2


def test_oracle_tape_run_reproduces_the_recorded_run(oracle):
    for line, n in _cases():
        tapes, used = oracle.record_tapes(line, n, SEED, 2048, 512, 2048)
        assert used.max() < 512 or used[:, 1 + line.M:].max() < 2048
        want = oracle.run(line, n, mode=abi.RNG_REPLAY, seed=SEED).stats
        got = oracle.run(line, n, mode=abi.RNG_TAPE, tapes=tapes).stats
        assert_same(got, want)


def test_kernel_source_tape_run(oracle):
    from hostsim import binding as H
    for line, n in _cases():
        tapes, _ = oracle.record_tapes(line, n, SEED, 2048, 512, 2048)
        want = oracle.run(line, n, mode=abi.RNG_REPLAY, seed=SEED).stats
        assert_same(H.run(line, n, mode=abi.RNG_TAPE, tapes=tapes), want)
        assert_same(H.run(line, n, mode=abi.RNG_TAPE, tapes=tapes, cold=True, slices=[50.0, 100.0]), want)


def test_exhausted_tapes_are_reported(oracle):
    from hostsim import binding as H
    line = example_line(run_to=400.0, warm_up=40.0)
    tapes, used = oracle.record_tapes(line, 2, SEED, 2048, 512, 2048)
    short_e = abi.DrawTapes(tapes.jobtype, tapes.uniform, tapes.exp[:, :, :20])
    short_j = abi.DrawTapes(tapes.jobtype[:, :50], tapes.uniform, tapes.exp)
    for t in (short_e, short_j):
        a, b = oracle.run(line, 2, mode=abi.RNG_TAPE, tapes=t).stats, H.run(line, 2, mode=abi.RNG_TAPE, tapes=t)
        assert set(a.status.tolist()) == {abi.INST_TAPE_EXHAUSTED} and np.array_equal(a.status, b.status)


@pytest.mark.gpu
def test_gpu_tape_run(engine, oracle):
    for line, n in _cases():
        tapes, _ = oracle.record_tapes(line, n, SEED, 2048, 512, 2048)
        want = oracle.run(line, n, mode=abi.RNG_REPLAY, seed=SEED).stats
        assert_same(engine.run_tape(line, n, tapes).stats, want)
    line = example_line(run_to=400.0, warm_up=40.0)
    tapes, _ = oracle.record_tapes(line, 2, SEED, 2048, 512, 2048)
    short = abi.DrawTapes(tapes.jobtype, tapes.uniform, tapes.exp[:, :, :20])
    assert set(engine.run_tape(line, 2, short).stats.status.tolist()) == {abi.INST_TAPE_EXHAUSTED}
    with pytest.raises(abi.MjpError):
        engine.upload(line, 2)
        engine.launch(abi.RNG_TAPE, 0, 0)            # tapes not set for this upload
