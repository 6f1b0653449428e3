"""TEST-ONLY: run the kernel SOURCE (mjp_line_kernel.cuh) on the host through
tests/hostsim/cuda_shim.h, one thread per block.  Used by the CPU test-suite to
check the kernel's control logic against the oracle where no GPU exists; never
imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from mjpdes_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.join(_HERE, "..", "..")
LIB_PATH = os.path.join(_HERE, "libhostsim.so")
_LIB = None


def build():
    deps = [os.path.join(_HERE, f) for f in ("hostsim.cpp", "cuda_shim.h")] + [
        os.path.join(_ROOT, "mjpdes_b200", "csrc", f)
        for f in ("mjp_line_kernel.cuh", "mjp_aux_kernels.cuh", "mjp_device.cuh", "mjp_params.h", "mjp_plan.h")] + [
        os.path.join(_ROOT, "include", "mjpdes_b200.h")]
    if (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        extra = os.environ.get("MJP_HOSTSIM_FLAGS", "").split()       # e.g. -DMJP_FLUSH_MERGED=1: an experiment's code path
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *extra, "-o", LIB_PATH,
                               os.path.join(_HERE, "hostsim.cpp")])


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(LIB_PATH)
        _LIB.hostsim_run.argtypes = [C.POINTER(abi.LineDesc), C.c_uint64, C.POINTER(abi.RngSpec),
                                     C.POINTER(abi.StatsOut), C.c_int, C.POINTER(abi.LogOut), C.c_int]
        _LIB.hostsim_run_sliced.argtypes = [C.POINTER(abi.LineDesc), C.c_uint64, C.POINTER(abi.RngSpec),
                                            C.POINTER(abi.StatsOut), C.c_int, C.POINTER(abi.LogOut), C.c_int,
                                            C.POINTER(C.c_double), C.c_int, C.POINTER(abi.DrawTapesStruct)]
        _LIB.hostsim_end_time.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double]
        _LIB.hostsim_end_time.restype = C.c_double
    return _LIB


def run(line, n_inst, mode=abi.RNG_COUNTER, seed=0, first_instance_id=0, want_hash=True, log_capacity=0, cold=False,
        slices=(), tapes=None, rotation=0, started_hot=None):
    stats = abi.HostStats(line, n_inst)
    d, r, s = line.desc(), abi.RngSpec(mode, 0, seed, first_instance_id), stats.struct()
    hlog = abi.HostLog(log_capacity, n_inst, line.M) if log_capacity else None
    sl = np.ascontiguousarray(slices, np.float64)
    lib().hostsim_set_rotation(int(rotation))
    # cold layouts: STARTED[M] in shared memory or in the scratch; by default alternate with the batch size so that
    # every caller (the fuzzers included) exercises both
    lib().hostsim_set_started_hot(int((n_inst & 1) if started_hot is None else started_hot))
    rc = lib().hostsim_run_sliced(C.byref(d), n_inst, C.byref(r), C.byref(s), int(want_hash),
                                  C.byref(hlog.struct()) if hlog else None, int(cold),
                                  abi.ptr(sl, C.c_double) if len(sl) else None, len(sl),
                                  C.byref(tapes.struct()) if tapes is not None else None)
    lib().hostsim_set_rotation(0)
    assert rc == 0, rc
    if hlog:
        stats.log = hlog.view().copy()
        stats.log_dropped = hlog.dropped
        stats.first_state = hlog.first_state
    return stats


def end_time(kinds, times, clock, dur):
    k = np.ascontiguousarray(kinds, np.uint8)
    t = np.ascontiguousarray(times, np.float64)
    return lib().hostsim_end_time(abi.ptr(k, C.c_uint8), abi.ptr(t, C.c_double), len(k), clock, dur)
