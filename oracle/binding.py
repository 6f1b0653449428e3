"""ctypes binding of the CPU ORACLE (``oracle/libmjp_oracle.so``).

TEST INFRASTRUCTURE.  Import this only from ``tests/``, ``__graft_entry__.smoke``
and ``bench.py`` (cpu_baseline / ``--impl reference``).  The product package
``mjpdes_b200`` never imports it.  Only the interface structs (the ctypes mirror
of ``include/mjpdes_b200.h``) are shared with the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mjpdes_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmjp_oracle.so")
_LIB = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mjp_oracle.cpp")
    deps = [src, os.path.join(_HERE, "mjp_oracle.h"), os.path.join(_HERE, "..", "include", "mjpdes_b200.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmjp_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class Extra(C.Structure):
    _fields_ = [("act_counts", C.POINTER(C.c_uint64)), ("iterations", C.POINTER(C.c_uint64)),
                ("ticks", C.POINTER(C.c_uint64)), ("max_tick_msgs", C.POINTER(C.c_uint32)),
                ("max_lookahead", C.POINTER(C.c_uint32)), ("wide_ticks", C.POINTER(C.c_uint64)),
                ("wide_ticks_reordered", C.POINTER(C.c_uint64)), ("dets", C.POINTER(C.c_int64)),
                ("dets_capacity", C.c_uint64), ("dets_len", C.POINTER(C.c_uint64))]


class Basics(C.Structure):
    _fields_ = [("runtime", C.c_double), ("TP", C.c_float), ("residence", C.c_double), ("njobs", C.c_int64),
                ("blocked", C.c_double * abi.MAX_MACHINES), ("starved", C.c_double * abi.MAX_MACHINES),
                ("occupancy", C.c_double * abi.MAX_MACHINES)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        pd, pu8 = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        l.mjpo_run.argtypes = [C.POINTER(abi.LineDesc), C.c_uint64, C.POINTER(abi.RngSpec),
                               C.POINTER(abi.StatsOut), C.POINTER(abi.LogOut), C.POINTER(Extra), C.c_int]
        l.mjpo_set_tapes.argtypes = [C.POINTER(abi.DrawTapesStruct)]
        l.mjpo_set_tapes.restype = None
        l.mjpo_end_time.argtypes = [pu8, pd, C.c_int32, C.c_double, C.c_double]
        l.mjpo_end_time.restype = C.c_double
        l.mjpo_updown_events.argtypes = [C.c_double, C.c_double, C.POINTER(abi.RngSpec), C.c_uint64, C.c_int32,
                                         C.c_int32, pu8, pd, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.mjpo_log_script.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.c_double, C.c_uint8, pu8,
                                      C.POINTER(abi.LogRecord), C.c_int32, C.POINTER(abi.StatsOut),
                                      C.POINTER(abi.LogOut)]
        l.mjpo_calc_basics.argtypes = [C.POINTER(abi.LineDesc), C.POINTER(abi.StatsOut), C.c_uint64, C.c_uint64,
                                       C.POINTER(Basics)]
        l.mjpo_calc_bneck.argtypes = [C.c_int32, pd, pd]
        l.mjpo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.mjpo_philox4x32_10.restype = None
        l.mjpo_clj_keyword_hash.argtypes = [C.c_char_p]
        l.mjpo_clj_keyword_hash.restype = C.c_uint32
        l.mjpo_clj_trie_key.argtypes = [C.c_uint32]
        l.mjpo_clj_trie_key.restype = C.c_uint64
        for f in (l.mjpo_log, l.mjpo_exp):
            f.argtypes = [C.c_double]
            f.restype = C.c_double
        for f in (l.mjpo_uniform, l.mjpo_open_uniform):
            f.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
            f.restype = C.c_double
        _LIB = l
    return _LIB


class OracleResult:
    def __init__(self, stats: abi.HostStats, log: np.ndarray | None, extra: dict, rc: int, first_state=None, dets=None):
        self.stats, self.log, self.extra, self.rc = stats, log, extra, rc
        self.first_state = first_state      # [n][2M], mjp_log_out.first_state
        self.dets = dets                    # job_detail: one (run, bufs) pair per record, the literal log/detail snapshots


def run(line: abi.HostLine, n_inst: int, mode: int = abi.RNG_COUNTER, seed: int = 0, first_instance_id: int = 0,
        log_capacity: int = 0, n_threads: int = 1, tapes: abi.DrawTapes | None = None,
        job_detail: bool = False) -> OracleResult:
    """main-loop-loop (core.clj:729-744) for ``n_inst`` instances on the CPU.  ``job_detail``: [:report :job-detail?],
    every printed message comes with its literal (log/detail model) snapshot (util/log.clj:11-28)."""
    l = lib()
    n = int(n_inst)
    stats = abi.HostStats(line, n)
    rng = abi.RngSpec(mode, 0, seed, first_instance_id)
    desc = line.desc()
    hlog = abi.HostLog(log_capacity, n, line.M) if log_capacity else None
    ex = dict(act_counts=np.zeros((10, n), np.uint64), iterations=np.zeros(n, np.uint64),
              ticks=np.zeros(n, np.uint64), max_tick_msgs=np.zeros(n, np.uint32),
              max_lookahead=np.zeros(n, np.uint32), wide_ticks=np.zeros(n, np.uint64),
              wide_ticks_reordered=np.zeros(n, np.uint64))
    e = Extra(abi.ptr(ex["act_counts"], C.c_uint64), abi.ptr(ex["iterations"], C.c_uint64),
              abi.ptr(ex["ticks"], C.c_uint64), abi.ptr(ex["max_tick_msgs"], C.c_uint32),
              abi.ptr(ex["max_lookahead"], C.c_uint32), abi.ptr(ex["wide_ticks"], C.c_uint64),
              abi.ptr(ex["wide_ticks_reordered"], C.c_uint64))
    dets_buf, dets_len = None, C.c_uint64(0)
    if job_detail and hlog:
        M = line.M
        per = 2 * M + int(np.sum(line.N)) + M                 # a snapshot never holds more ids than the line holds jobs
        dets_buf = np.zeros(max(1, int(log_capacity) * per), np.int64)
        e.dets, e.dets_capacity, e.dets_len = abi.ptr(dets_buf, C.c_int64), len(dets_buf), C.pointer(dets_len)
    s = stats.struct()
    ts = tapes.struct() if tapes is not None else None
    l.mjpo_set_tapes(C.byref(ts) if ts is not None else None)
    rc = l.mjpo_run(C.byref(desc), n, C.byref(rng), C.byref(s),
                    C.byref(hlog.struct()) if hlog else None, C.byref(e), int(n_threads))
    l.mjpo_set_tapes(None)
    if rc not in (abi.OK, abi.LOG_TRUNCATED):
        raise abi.MjpError(rc, "oracle rejected the descriptor")
    dets = None
    if dets_buf is not None:
        assert dets_len.value <= len(dets_buf)
        dets, q, M = [], 0, line.M
        while q < dets_len.value:
            run_ = [int(x) if x >= 0 else None for x in dets_buf[q:q + M]]
            q += M
            bufs = []
            for _ in range(M - 1):
                c = int(dets_buf[q])
                bufs.append([int(x) for x in dets_buf[q + 1:q + 1 + c]])
                q += 1 + c
            dets.append((run_, bufs))
    return OracleResult(stats, hlog.view().copy() if hlog else None, ex, rc, hlog.first_state if hlog else None, dets)


def record_tapes(line: abi.HostLine, n_inst: int, seed: int, len_jobtype: int, len_uniform: int, len_exp: int,
                 first_instance_id: int = 0):
    """Run REPLAY and record every draw per consumer: what hooking rand / sample-exp in a JVM run would give."""
    l = lib()
    l.mjpo_record_tapes.argtypes = [C.POINTER(abi.LineDesc), C.c_uint64, C.POINTER(abi.RngSpec),
                                    C.POINTER(abi.DrawTapesStruct), C.POINTER(C.c_uint32)]
    n, M = int(n_inst), line.M
    tapes = abi.DrawTapes(np.full((n, len_jobtype), np.nan), np.full((n, M, len_uniform), np.nan),
                          np.full((n, M, len_exp), np.nan))
    used = np.zeros((n, 2 * M + 1), np.uint32)
    d, r, t = line.desc(), abi.RngSpec(abi.RNG_REPLAY, 0, seed, first_instance_id), tapes.struct()
    rc = l.mjpo_record_tapes(C.byref(d), n, C.byref(r), C.byref(t), abi.ptr(used, C.c_uint32))
    assert rc == 0, rc
    return tapes, used


def end_time(kinds, times, clock: float, dur: float) -> float:
    k = np.ascontiguousarray(kinds, np.uint8)
    t = np.ascontiguousarray(times, np.float64)
    return lib().mjpo_end_time(abi.ptr(k, C.c_uint8), abi.ptr(t, C.c_double), len(k), clock, dur)


def updown_events(lam: float, mu: float, n: int, mode: int, seed: int = 0, inst: int = 0, machine: int = 0):
    kinds = np.zeros(n, np.uint8)
    times = np.zeros(n, np.float64)
    du, de = C.c_uint32(0), C.c_uint32(0)
    rng = abi.RngSpec(mode, 0, seed, 0)
    got = lib().mjpo_updown_events(lam, mu, C.byref(rng), inst, machine, n, abi.ptr(kinds, C.c_uint8),
                                   abi.ptr(times, C.c_double), C.byref(du), C.byref(de))
    return kinds[:got], times[:got], du.value, de.value


def log_script(M: int, N, warm_up: float, script, up_down: bool = False, log_capacity: int = 256):
    """script: list of ("log", act, machine, clk[, n[, aux[, job]]]) or ("push",)."""
    Narr = np.ascontiguousarray(N, np.int32)
    ops = np.zeros(len(script), np.uint8)
    msgs = np.zeros(len(script), abi.LOG_RECORD_DTYPE)
    for j, s in enumerate(script):
        if s[0] == "push":
            ops[j] = 1
        else:
            _, act, m, clk, *rest = s
            msgs[j]["act"], msgs[j]["machine"], msgs[j]["clk"] = act, m, clk
            msgs[j]["n"] = rest[0] if len(rest) > 0 else 0
            msgs[j]["aux"] = rest[1] if len(rest) > 1 else np.nan
            msgs[j]["job"] = rest[2] if len(rest) > 2 else 0
    line = abi.HostLine(lam=np.zeros(M), mu=np.ones(M), W=np.ones(M), discipline=np.zeros(M, np.uint8), N=Narr,
                        portion=[1.0], w=np.ones((1, M)), warm_up=warm_up, run_to=1.0)
    stats = abi.HostStats(line, 1)
    hlog = abi.HostLog(log_capacity)
    st = stats.struct()
    lib().mjpo_log_script(M, abi.ptr(Narr, C.c_int32) if len(Narr) else None, warm_up, int(up_down),
                          abi.ptr(ops, C.c_uint8), C.cast(msgs.ctypes.data, C.POINTER(abi.LogRecord)), len(script),
                          C.byref(st), C.byref(hlog.struct()))
    return stats, hlog.view().copy()


def calc_basics(line: abi.HostLine, stats: abi.HostStats, inst: int) -> dict:
    b = Basics()
    d, s = line.desc(), stats.struct()
    lib().mjpo_calc_basics(C.byref(d), C.byref(s), stats.n_inst, inst, C.byref(b))
    M = line.M
    return dict(runtime=b.runtime, TP=b.TP, residence=b.residence, njobs=b.njobs, blocked=list(b.blocked[:M]),
                starved=list(b.starved[:M]), occupancy=list(b.occupancy[: M - 1]))


def calc_bneck(blocked, starved) -> int:
    bl = np.ascontiguousarray(blocked, np.float64)
    st = np.ascontiguousarray(starved, np.float64)
    return lib().mjpo_calc_bneck(len(bl), abi.ptr(bl, C.c_double), abi.ptr(st, C.c_double))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().mjpo_philox4x32_10(c, k, o)
    return list(o)
