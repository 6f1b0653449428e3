"""ctypes mirror of ``include/mjpdes_b200.h`` and the loader of ``libmjpdes_b200.so``.

This is the binding a host language writes against the C ABI (the Clojure/JNA
form of the same binding is in INTEGRATION.md).  There is no CPU fallback: if the
CUDA library is missing or fails to load, :func:`load_library` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ABI_VERSION = 3

# call status
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, LOG_TRUNCATED = range(6)
# per-instance status  ([:params :status], core.clj:735,739)
INST_NORMAL_END, INST_NO_RUNABLES, INST_CLOCK_BACK, INST_JOBTYPE_NIL, INST_LOGBUF_EMPTY, INST_TICK_OVERFLOW = range(6)
INST_ITER_LIMIT, INST_PAUSED, INST_TAPE_EXHAUSTED = 6, 7, 8
INST_NOT_RUN = 255
INST_STATUS_KEYWORD = {
    INST_NORMAL_END: ":normal-end",
    INST_NO_RUNABLES: ":no-runables",
    INST_CLOCK_BACK: ":clock-running-backwards",
    INST_JOBTYPE_NIL: ":jobtype-nil",
    INST_LOGBUF_EMPTY: ":log-buf-empty",
    INST_TICK_OVERFLOW: ":tick-overflow",
    INST_ITER_LIMIT: ":iteration-limit",
    INST_PAUSED: ":paused",
    INST_TAPE_EXHAUSTED: ":tape-exhausted",
    INST_NOT_RUN: ":not-run",
}
# actions
ACT_AJ, ACT_UP, ACT_DOWN, ACT_BJ, ACT_EJ, ACT_SM, ACT_ST, ACT_US, ACT_BL, ACT_UB = range(10)
ACT_KEYWORD = [":aj", ":up", ":down", ":bj", ":ej", ":sm", ":st", ":us", ":bl", ":ub"]
BAS, BBS = 0, 1
RNG_COUNTER, RNG_REPLAY, RNG_TAPE = 0, 1, 2
CFG_EVENT_HASH = 1
MAX_MACHINES = 64
MAX_JOBTYPES = 255
UINT64_MAX = (1 << 64) - 1

_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pu64 = C.POINTER(C.c_uint64)
_pu8 = C.POINTER(C.c_uint8)


class LineDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("K", C.c_int32),
        ("lam", _pd), ("mu", _pd), ("W", _pd), ("discipline", _pu8), ("N", _pi32),
        ("portion", _pd), ("w", _pd),
        ("warm_up", C.c_double), ("run_to", C.c_double), ("run_to_job", C.c_int64),
        ("jm2dm", C.c_uint8), ("log", C.c_uint8), ("up_down", C.c_uint8), ("reserved0", C.c_uint8),
        ("reserved1", C.c_uint32), ("max_lines", C.c_uint64),
        ("lam_inst", _pd), ("mu_inst", _pd), ("W_inst", _pd), ("N_inst", _pi32), ("w_inst", _pd),
        ("group_rank", _pi32),
    ]


class RngSpec(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64),
                ("first_instance_id", C.c_uint64)]


class DrawTapesStruct(C.Structure):
    _fields_ = [("len_jobtype", C.c_uint32), ("len_uniform", C.c_uint32), ("len_exp", C.c_uint32),
                ("reserved", C.c_uint32), ("jobtype", _pd), ("uniform", _pd), ("exp", _pd)]


class DrawTapes:
    """Recorded draws for RNG_TAPE: jobtype [n][Lj], uniform [n][M][Lu], exp [n][M][Le] (f64, instance-major)."""

    def __init__(self, jobtype, uniform, exp):
        self.jobtype = np.ascontiguousarray(jobtype, np.float64)
        self.uniform = np.ascontiguousarray(uniform, np.float64)
        self.exp = np.ascontiguousarray(exp, np.float64)
        assert self.jobtype.ndim == 2 and self.uniform.ndim == 3 and self.exp.ndim == 3

    def struct(self) -> "DrawTapesStruct":
        t = DrawTapesStruct()
        t.len_jobtype, t.len_uniform, t.len_exp = self.jobtype.shape[1], self.uniform.shape[2], self.exp.shape[2]
        t.jobtype, t.uniform, t.exp = ptr(self.jobtype, C.c_double), ptr(self.uniform, C.c_double), ptr(self.exp, C.c_double)
        return t


class StatsOut(C.Structure):
    _fields_ = [
        ("njobs", _pi64), ("residence_sum", _pd), ("blocked_time", _pd), ("starved_time", _pd),
        ("occ_time", _pd), ("final_clock", _pd), ("current_job", _pi64), ("n_events", _pu64),
        ("event_hash", _pu64), ("status", _pu8), ("aggregate", _pd),
    ]


class LogRecord(C.Structure):
    _fields_ = [("clk", C.c_double), ("aux", C.c_double), ("job", C.c_uint32), ("instance", C.c_uint32),
                ("line", C.c_uint32), ("n", C.c_uint8), ("ord", C.c_uint8), ("act", C.c_uint8), ("machine", C.c_uint8)]


LOG_RECORD_DTYPE = np.dtype([("clk", "<f8"), ("aux", "<f8"), ("job", "<u4"), ("instance", "<u4"),
                             ("line", "<u4"), ("n", "u1"), ("ord", "u1"), ("act", "u1"), ("machine", "u1")])
assert LOG_RECORD_DTYPE.itemsize == 32 and C.sizeof(LogRecord) == 32


class LogOut(C.Structure):
    _fields_ = [("records", C.POINTER(LogRecord)), ("capacity", C.c_uint64), ("count", C.c_uint64),
                ("dropped", C.c_uint64), ("first_state", C.POINTER(C.c_uint32))]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_int32), ("max_instances", C.c_uint64)]


REDUCE_PEER, REDUCE_NCCL = 0, 1


class MultiConfig(C.Structure):
    _fields_ = [("n_devices", C.c_int32), ("devices", C.POINTER(C.c_int32)), ("flags", C.c_int32),
                ("reduce", C.c_int32)]


def aggregate_len(M: int) -> int:
    return 1 + 2 * (3 * M + 1)


def ptr(a: np.ndarray | None, ctype):
    """Pointer to a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


class HostLine:
    """Owns the numpy arrays behind a :class:`LineDesc` (the flat SoA form of the
    preprocessed Model, core.clj:578-615) and keeps them alive."""

    def __init__(self, lam, mu, W, discipline, N, portion, w, warm_up, run_to, run_to_job=-1,
                 jm2dm=False, log=False, up_down=False, max_lines=0,
                 lam_inst=None, mu_inst=None, W_inst=None, N_inst=None, w_inst=None, group_rank=None):
        f8 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self.lam, self.mu, self.W = f8(lam), f8(mu), f8(W)
        self.discipline = np.ascontiguousarray(discipline, dtype=np.uint8)
        self.N = np.ascontiguousarray(N, dtype=np.int32)
        self.portion = f8(portion)
        self.w = f8(w)
        self.M = int(self.lam.shape[0])
        self.K = int(self.portion.shape[0])
        self.w = self.w.reshape(self.K, self.M)
        self.warm_up, self.run_to, self.run_to_job = float(warm_up), float(run_to), int(run_to_job)
        self.jm2dm, self.log, self.up_down = bool(jm2dm), bool(log), bool(up_down)
        if max_lines is None or max_lines == float("inf"):
            max_lines = UINT64_MAX
        self.max_lines = int(max_lines)
        self.lam_inst, self.mu_inst, self.W_inst, self.w_inst = f8(lam_inst), f8(mu_inst), f8(W_inst), f8(w_inst)
        self.N_inst = None if N_inst is None else np.ascontiguousarray(N_inst, dtype=np.int32)
        # hash-map rank of the machine names (util/log.clj:44 for ticks with > 8 machines); None = :m1 .. :mM
        self.group_rank = None if group_rank is None else np.ascontiguousarray(group_rank, dtype=np.int32)
        assert self.mu.shape == (self.M,) and self.W.shape == (self.M,) and self.discipline.shape == (self.M,)
        assert self.N.shape == (max(self.M - 1, 0),)

    @property
    def n_levels(self) -> int:
        return int((self.N + 1).sum())

    @property
    def level_offsets(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.N + 1)]).astype(np.int64)

    def desc(self) -> LineDesc:
        d = LineDesc()
        d.M, d.K = self.M, self.K
        d.lam, d.mu, d.W = ptr(self.lam, C.c_double), ptr(self.mu, C.c_double), ptr(self.W, C.c_double)
        d.discipline = ptr(self.discipline, C.c_uint8)
        d.N = ptr(self.N, C.c_int32)
        d.portion, d.w = ptr(self.portion, C.c_double), ptr(self.w, C.c_double)
        d.warm_up, d.run_to, d.run_to_job = self.warm_up, self.run_to, self.run_to_job
        d.jm2dm, d.log, d.up_down = int(self.jm2dm), int(self.log), int(self.up_down)
        d.max_lines = self.max_lines
        d.lam_inst, d.mu_inst = ptr(self.lam_inst, C.c_double), ptr(self.mu_inst, C.c_double)
        d.W_inst, d.w_inst = ptr(self.W_inst, C.c_double), ptr(self.w_inst, C.c_double)
        d.N_inst = ptr(self.N_inst, C.c_int32)
        d.group_rank = ptr(self.group_rank, C.c_int32)
        return d


class HostStats:
    """Caller-owned result buffers of one call (instance-fastest SoA)."""

    def __init__(self, line: HostLine, n_inst: int, want_hash: bool = True, want_aggregate: bool = True):
        n, M, L = int(n_inst), line.M, line.n_levels
        self.n_inst, self.M, self.L = n, M, L
        self.njobs = np.zeros(n, np.int64)
        self.residence_sum = np.zeros(n, np.float64)
        self.blocked_time = np.zeros((M, n), np.float64)
        self.starved_time = np.zeros((M, n), np.float64)
        self.occ_time = np.zeros((L, n), np.float64)
        self.final_clock = np.zeros(n, np.float64)
        self.current_job = np.zeros(n, np.int64)
        self.n_events = np.zeros(n, np.uint64)
        self.event_hash = np.zeros(n, np.uint64) if want_hash else None
        self.status = np.full(n, INST_NOT_RUN, np.uint8)
        self.aggregate = np.zeros(aggregate_len(M), np.float64) if want_aggregate else None

    def struct(self) -> StatsOut:
        s = StatsOut()
        s.njobs = ptr(self.njobs, C.c_int64)
        s.residence_sum = ptr(self.residence_sum, C.c_double)
        s.blocked_time = ptr(self.blocked_time, C.c_double)
        s.starved_time = ptr(self.starved_time, C.c_double)
        s.occ_time = ptr(self.occ_time, C.c_double)
        s.final_clock = ptr(self.final_clock, C.c_double)
        s.current_job = ptr(self.current_job, C.c_int64)
        s.n_events = ptr(self.n_events, C.c_uint64)
        s.event_hash = ptr(self.event_hash, C.c_uint64)
        s.status = ptr(self.status, C.c_uint8)
        s.aggregate = ptr(self.aggregate, C.c_double)
        return s


class HostLog:
    """Caller buffers of mjp_log_out.  ``n_inst`` and ``M`` given: also ``first_state`` [n_inst][2M], where the jobs
    were at each instance's first printed message (what the :dets snapshots start from, scada.DetsTracker)."""

    def __init__(self, capacity: int, n_inst: int = 0, M: int = 0):
        self.records = np.zeros(int(capacity), LOG_RECORD_DTYPE)
        self._s = LogOut()
        self._s.records = C.cast(self.records.ctypes.data, C.POINTER(LogRecord))
        self._s.capacity = int(capacity)
        self.first_state = None
        if n_inst and M:
            self.first_state = np.full((int(n_inst), 2 * int(M)), 0xFFFFFFFF, np.uint32)
            self._s.first_state = C.cast(self.first_state.ctypes.data, C.POINTER(C.c_uint32))

    def struct(self) -> LogOut:
        return self._s

    @property
    def count(self) -> int:
        return int(self._s.count)

    @property
    def dropped(self) -> int:
        return int(self._s.dropped)

    def view(self) -> np.ndarray:
        return self.records[: self.count]


_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libmjpdes_b200.so")


class MjpError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"mjpdes_b200 status {code}: {text}")
        self.code = code


def load_library(path: str | None = None) -> C.CDLL:
    """Load the CUDA library.  Raises if it is absent: there is no CPU path."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.environ.get("MJP_LIB") or LIB_PATH      # MJP_LIB: another build of the same library (A/B measurements)
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with `python __graft_entry__.py build` "
                          "(mjpdes_b200 has no CPU fallback)")
    lib = C.CDLL(p)
    H = C.c_void_p
    lib.mjp_abi_version.restype = C.c_int
    lib.mjp_device_count.restype = C.c_int
    lib.mjp_aggregate_len.argtypes = [C.c_int32]
    lib.mjp_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    lib.mjp_destroy.argtypes = [H]
    lib.mjp_destroy.restype = None
    lib.mjp_last_error.argtypes = [H]
    lib.mjp_last_error.restype = C.c_char_p
    lib.mjp_run.argtypes = [H, C.POINTER(LineDesc), C.c_uint64, C.POINTER(RngSpec), C.POINTER(StatsOut),
                            C.POINTER(LogOut)]
    lib.mjp_upload.argtypes = [H, C.POINTER(LineDesc), C.c_uint64]
    lib.mjp_set_tapes.argtypes = [H, C.POINTER(DrawTapesStruct)]
    lib.mjp_reserve_log.argtypes = [H, C.c_uint64]
    lib.mjp_launch.argtypes = [H, C.POINTER(RngSpec), C.c_int]
    lib.mjp_launch_slice.argtypes = [H, C.POINTER(RngSpec), C.c_int, C.c_double, C.c_int]
    lib.mjp_sync.argtypes = [H]
    lib.mjp_download.argtypes = [H, C.POINTER(StatsOut), C.POINTER(LogOut)]
    lib.mjp_last_kernel_ms.argtypes = [H, C.POINTER(C.c_float)]
    lib.mjp_last_launch_count.argtypes = [H, C.POINTER(C.c_int)]
    lib.mjp_probe_end_time.argtypes = [H, _pu8, _pd, C.c_int32, C.c_double, C.c_double, _pd]
    lib.mjp_device_aggregate.argtypes = [H, C.POINTER(_pd), C.POINTER(C.c_int32)]
    lib.mjp_multi_create.argtypes = [C.POINTER(MultiConfig), C.POINTER(H)]
    lib.mjp_multi_destroy.argtypes = [H]
    lib.mjp_multi_destroy.restype = None
    lib.mjp_multi_last_error.argtypes = [H]
    lib.mjp_multi_last_error.restype = C.c_char_p
    lib.mjp_multi_device_count.argtypes = [H]
    lib.mjp_multi_run.argtypes = [H, C.POINTER(LineDesc), C.c_uint64, C.POINTER(RngSpec), C.POINTER(StatsOut)]
    lib.mjp_multi_last_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    if lib.mjp_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI version mismatch: library {lib.mjp_abi_version()} vs binding {ABI_VERSION}")
    if path is None:
        _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "mjp_abi_version", "mjp_device_count", "mjp_aggregate_len", "mjp_create", "mjp_destroy", "mjp_last_error",
    "mjp_run", "mjp_upload", "mjp_set_tapes", "mjp_reserve_log", "mjp_launch", "mjp_launch_slice", "mjp_sync", "mjp_download", "mjp_last_kernel_ms",
    "mjp_last_launch_count", "mjp_probe_end_time", "mjp_device_aggregate",
    "mjp_multi_create", "mjp_multi_destroy", "mjp_multi_last_error", "mjp_multi_device_count", "mjp_multi_run",
    "mjp_multi_last_ms",
]
