"""CPU-side check of the KERNEL SOURCE: mjp_line_kernel.cuh compiled as host C++ through
tests/hostsim/cuda_shim.h (one thread per block) must reproduce the oracle's event sequence
(hash) and accumulators bit-for-bit.  This guards the kernel's control logic where no GPU
exists; the GPU parity tests proper are in test_gpu_parity.py."""
import numpy as np
import pytest

from hostsim import binding as H
from lines import (GOLDENS, SEED, example_line, golden_two_machine, long50_line, mixed20_line, random_line,
                   sweep10_line)
from mjpdes_b200 import abi
from parity import assert_same

MODES = [abi.RNG_COUNTER, abi.RNG_REPLAY]
KIND = {":up": 0, ":down": 1}


@pytest.mark.parametrize("gid", ["G1", "G2"])
def test_lazy_end_time_matches_reference_known_answers(gid):
    g = GOLDENS[gid]
    got = H.end_time([KIND[k] for _, k, _ in g["events"]], [t for _, _, t in g["events"]], g["clock"], g["dur"])
    assert got == g["expected"]


@pytest.mark.parametrize("mode", MODES)
def test_example_line(oracle, mode):
    line = example_line(run_to=4000.0, warm_up=400.0)
    assert_same(H.run(line, 6, mode=mode, seed=SEED), oracle.run(line, 6, mode=mode, seed=SEED).stats)


@pytest.mark.parametrize("gid", ["G7", "G8", "G9"])
def test_reference_two_machine_models(oracle, gid):
    line = golden_two_machine(gid, lam=0.0, run_to=300.0)
    assert_same(H.run(line, 2, mode=abi.RNG_REPLAY, seed=SEED), oracle.run(line, 2, mode=abi.RNG_REPLAY, seed=SEED).stats)


@pytest.mark.parametrize("trial", range(16))
def test_random_lines(oracle, trial):
    rng = np.random.default_rng(1000 + trial)
    M, K = int(rng.integers(2, 13)), int(rng.integers(1, 5))
    line = random_line(rng, M, K, run_to=300.0, warm_up=30.0, jm2dm=bool(rng.random() < 0.3),
                       up_down=bool(rng.random() < 0.4))
    mode = MODES[trial % 2]
    assert_same(H.run(line, 3, mode=mode, seed=trial), oracle.run(line, 3, mode=mode, seed=trial).stats)


@pytest.mark.parametrize("M", [20, 33, 64])
def test_wide_lines_use_64bit_masks(oracle, M):
    line = random_line(np.random.default_rng(M), M, 3, run_to=120.0, warm_up=10.0)
    assert_same(H.run(line, 2, seed=5), oracle.run(line, 2, seed=5).stats)


def test_identical_machines_many_time_ties(oracle):
    """38 % of ticks are genuine cross-machine time ties on a line of identical machines."""
    line = abi.HostLine(lam=[.1] * 6, mu=[.9] * 6, W=[1.] * 6, discipline=[0, 1, 0, 1, 0, 0], N=[2] * 5,
                        portion=[1.], w=[[1.] * 6], warm_up=50, run_to=1200)
    for mode in MODES:
        assert_same(H.run(line, 4, mode=mode, seed=3), oracle.run(line, 4, mode=mode, seed=3).stats)


def test_named_configs(oracle):
    for line, n in ((sweep10_line(5, warm_up=50, run_to=500, first=123456), 5),
                    (mixed20_line(warm_up=50, run_to=300, log=False), 2),
                    (long50_line(warm_up=50, run_to=250), 2)):
        assert_same(H.run(line, n, seed=SEED), oracle.run(line, n, seed=SEED).stats)


def test_jobtype_nil_status(oracle):
    """p0 < p_{K-1} lets the draw fall off the shifted thresholds -> the reference dies in add-job,
    before anything of that batch is executed or logged: the tick before it is never pushed (found by the wild
    profile of tests/fuzz/cpu_fuzz.py: event count and accumulators of the dying instance)."""
    for portion, warm_up in (([0.1, 0.2, 0.7], 0), ([0.45, 0.55], 0), ([0.48, 0.52], 5)):
        K = len(portion)
        line = abi.HostLine(lam=[.1, .1, .1], mu=[.9, .9, .9], W=[1., 1., 1.], discipline=[0, 1, 0], N=[2, 1], portion=portion,
                            w=np.ones((K, 3)), warm_up=warm_up, run_to=500)
        for mode in (abi.RNG_COUNTER, abi.RNG_REPLAY):
            a, b = H.run(line, 8, seed=1, mode=mode), oracle.run(line, 8, seed=1, mode=mode).stats
            assert set(b.status.tolist()) == {abi.INST_JOBTYPE_NIL}
            assert_same(a, b)
        if K == 2:
            assert b.n_events.max() > 50 and b.blocked_time.sum() + b.starved_time.sum() > 0


def test_aggregate_matches_oracle(oracle):
    line = example_line(run_to=1500.0, warm_up=100.0)
    a, b = H.run(line, 7, seed=9), oracle.run(line, 7, seed=9).stats
    assert a.aggregate[0] == 7
    assert np.allclose(a.aggregate, b.aggregate, rtol=1e-12)


@pytest.mark.parametrize("trial", range(6))
def test_cold_state_layout(oracle, trial):
    """COLD variants (wide lines): the cold per-instance state lives in a global scratch instead of shared
    memory; same results bit for bit."""
    rng = np.random.default_rng(500 + trial)
    M = [3, 10, 20, 33, 50, 64][trial]
    line = random_line(rng, M, int(rng.integers(1, 5)), run_to=150.0, warm_up=15.0, up_down=bool(trial % 2))
    mode = MODES[trial % 2]
    want = oracle.run(line, 3, mode=mode, seed=trial).stats
    for started_hot in (False, True):            # STARTED[M] in the scratch / in shared memory (host's choice, mjp_abi.cu)
        assert_same(H.run(line, 3, mode=mode, seed=trial, cold=True, started_hot=started_hot), want)


def test_cold_state_layout_with_scada_log(oracle):
    line = mixed20_line(warm_up=20, run_to=200)
    want = oracle.run(line, 2, seed=SEED, log_capacity=100_000)
    got = H.run(line, 2, seed=SEED, log_capacity=100_000, cold=True)
    a, b = np.sort(got.log, order=["instance", "line"]), np.sort(want.log, order=["instance", "line"])
    assert len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in ("clk", "job", "line", "n", "act", "machine"))


def test_run_to_job_keeps_the_loop_going(oracle):
    """end-main-loop? (core.clj:722-727): the loop continues while current-job <= :run-to-job OR clock <= :run-to-time."""
    line = example_line(run_to=200.0, warm_up=20.0, run_to_job=400)
    a, b = H.run(line, 4, seed=11), oracle.run(line, 4, seed=11).stats
    assert_same(a, b)
    assert np.all(b.current_job > 400) and np.all(b.final_clock > 200.0)
    short = oracle.run(example_line(run_to=200.0, warm_up=20.0), 4, seed=11).stats
    assert np.all(short.current_job < 400)


def test_per_instance_capacity_and_rate_planes(oracle):
    """All five override planes at once (lambda, mu, W, N, w per instance)."""
    rng = np.random.default_rng(42)
    n, M, K = 6, 4, 2
    line = abi.HostLine(lam=[.1] * M, mu=[.9] * M, W=[1.] * M, discipline=[0, 1, 0, 0], N=[4, 4, 4], portion=[.5, .5],
                        w=np.ones((K, M)), warm_up=20, run_to=400,
                        lam_inst=rng.uniform(0.05, 0.2, (M, n)), mu_inst=rng.uniform(0.5, 1.5, (M, n)),
                        W_inst=rng.uniform(0.8, 1.2, (M, n)), N_inst=rng.integers(1, 5, (M - 1, n)),
                        w_inst=rng.uniform(0.5, 1.5, (K * M, n)))
    assert_same(H.run(line, n, seed=2), oracle.run(line, n, seed=2).stats)


def test_kernel_source_under_address_and_ub_sanitizers(tmp_path):
    """compute-sanitizer is closed on the GPU pool; the kernel source runs under ASan + UBSan on the host
    instead (tests/hostsim/asan_main.cpp): every plane is an exact-size heap buffer and the unused tail of
    the shared-memory arena is poisoned, so a wrong slot offset or plane index traps."""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
    exe = str(tmp_path / "asan_main")
    # -DMJP_HOSTSIM_MIN: only the kernel instantiations asan_main.cpp runs (hash on, no slicing; lines of 2, 5 and 8
    # machines, the generic, cold and 64-bit kernels, with and without capture, both RNG modes)
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=address,undefined",
                           "-fno-sanitize-recover=undefined", "-DMJP_HOSTSIM_MIN", "-o", exe,
                           os.path.join(here, "asan_main.cpp")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and "asan run clean" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("cold", [False, True])
def test_time_slicing_is_invisible(oracle, mode, cold):
    """mjp_launch_slice: parking the complete instance state at arbitrary slice boundaries and resuming gives
    the results of one uninterrupted run, bit for bit (statistics, event hash, SCADA records)."""
    slices = [0.0, 37.5, 37.5, 120.0, 121.0, 333.3, 800.0, 5000.0]      # repeated and past-the-end boundaries too
    for line in (example_line(run_to=900.0, warm_up=90.0),
                 random_line(np.random.default_rng(9), 11, 3, run_to=500.0, warm_up=50.0, up_down=True),
                 random_line(np.random.default_rng(10), 40, 2, run_to=200.0, warm_up=20.0)):
        want = oracle.run(line, 3, mode=mode, seed=21).stats
        assert_same(H.run(line, 3, mode=mode, seed=21, cold=cold, slices=slices), want)


def test_time_slicing_with_scada_log(oracle):
    line = example_line(run_to=600.0, warm_up=60.0, log=True, up_down=True, max_lines=abi.UINT64_MAX)
    want = oracle.run(line, 2, mode=abi.RNG_REPLAY, seed=5, log_capacity=100_000)
    got = H.run(line, 2, mode=abi.RNG_REPLAY, seed=5, log_capacity=100_000, slices=[50.0, 61.0, 300.0, 599.9])
    a, b = np.sort(got.log, order=["instance", "line"]), np.sort(want.log, order=["instance", "line"])
    assert len(a) == len(b)
    for f in ("clk", "job", "line", "n", "act", "machine"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["aux"].view(np.uint64), b["aux"].view(np.uint64))
    assert_same(got, want.stats)


def test_fuzz_profiles_sample(oracle):
    """A fixed sample of both profiles of tests/fuzz/cpu_fuzz.py (standard: trials 0.., incl. MJP_RNG_TAPE, COLD, slices
    and SCADA capture; wild: deep buffers, unequal portions, extreme rates, override planes)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz"))
    import cpu_fuzz
    for wild, first in ((False, 0), (True, 3_000_000)):
        cpu_fuzz.WILD = wild
        for t in range(first, first + 300):
            cpu_fuzz.one(t)
    cpu_fuzz.WILD = False


@pytest.mark.parametrize("mode", MODES)
def test_rotation_is_invisible(oracle, mode):
    """Rotation (the work queue of (chunk, slice) items the SLICE variants can pull from, mjp_line_kernel.cuh): with
    one thread per block a single "CTA" drains the queue in order, i.e. every instance is restored, advanced one
    slice and parked per item -- results equal an uninterrupted run, whatever the number of slices."""
    for line, n, cold in ((example_line(run_to=900.0, warm_up=90.0), 12, False),
                          (mixed20_line(warm_up=50, run_to=400, log=False), 5, False),
                          (long50_line(warm_up=50, run_to=300), 4, True),
                          (example_line(run_to=300.0, warm_up=30.0, run_to_job=500), 7, False)):
        want = oracle.run(line, n, mode=mode, seed=SEED, n_threads=4).stats
        for slices in (2, 5, 16):
            assert_same(H.run(line, n, mode=mode, seed=SEED, cold=cold, rotation=slices), want)


def test_rotation_with_scada_log(oracle):
    line = example_line(run_to=600.0, warm_up=60.0, log=True, up_down=True, max_lines=abi.UINT64_MAX)
    want = oracle.run(line, 6, seed=SEED, log_capacity=60000)
    got = H.run(line, 6, seed=SEED, log_capacity=60000, rotation=7)
    assert_same(got, want.stats)
    key = lambda r: np.argsort((r["instance"].astype(np.uint64) << np.uint64(32)) | r["line"].astype(np.uint64))
    a, b = got.log[key(got.log)], want.log[key(want.log)]
    assert len(a) == len(b) > 30000
    for f in ("instance", "line", "act", "machine", "n", "job", "clk"):
        assert np.array_equal(a[f], b[f]), f
