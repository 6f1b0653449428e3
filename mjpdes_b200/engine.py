"""Thin object wrapper over the C ABI (``include/mjpdes_b200.h``).

``Handle.run`` is the call that replaces ``core/main-loop-loop`` (core.clj:729-744)
for a batch of independent line instances.  Everything here goes through
``libmjpdes_b200.so``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class RunResult:
    """Per-instance ``*log-steady*`` content plus loop status, as numpy arrays."""

    def __init__(self, line: abi.HostLine, stats: abi.HostStats, log: np.ndarray | None, rc: int,
                 kernel_ms: float, launches: int, first_state: np.ndarray | None = None):
        self.line, self.stats, self.log, self.rc = line, stats, log, rc
        self.first_state = first_state          # [n_inst][2M], see mjp_log_out.first_state (SCADA capture only)
        self.kernel_ms, self.launches = kernel_ms, launches

    @property
    def n_inst(self) -> int:
        return self.stats.n_inst

    @property
    def total_events(self) -> int:
        return int(self.stats.n_events.sum())


class Handle:
    """One CUDA device, one stream.  Not thread-safe; distinct handles are."""

    def __init__(self, device: int = 0, event_hash: bool = False, lib_path: str | None = None):
        self._lib = abi.load_library(lib_path)
        if self._lib.mjp_device_count() <= device:
            raise abi.MjpError(abi.ERR_CUDA, f"CUDA device {device} not present "
                               f"({self._lib.mjp_device_count()} visible); mjpdes_b200 has no CPU path")
        cfg = abi.Config(device, abi.CFG_EVENT_HASH if event_hash else 0, 0)
        self._h = C.c_void_p()
        rc = self._lib.mjp_create(C.byref(cfg), C.byref(self._h))
        if rc != abi.OK:
            raise abi.MjpError(rc, (self._lib.mjp_last_error(None) or b"").decode())
        self.device = device
        self.event_hash = event_hash
        self._line = None
        self._n = 0

    # -- lifecycle ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.mjp_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=(abi.OK,)):
        if rc not in allow:
            raise abi.MjpError(rc, (self._lib.mjp_last_error(self._h) or b"").decode())
        return rc

    # -- the drop-in call ----------------------------------------------------------
    def run(self, line: abi.HostLine, n_inst: int, mode: int = abi.RNG_COUNTER, seed: int = 0,
            first_instance_id: int = 0, log_capacity: int = 0, want_aggregate: bool = True) -> RunResult:
        """Blocking: host buffers in, host buffers out (copies included)."""
        n = int(n_inst)
        stats = abi.HostStats(line, n, want_hash=True, want_aggregate=want_aggregate)
        rng = abi.RngSpec(mode, 0, seed, first_instance_id)
        desc, st = line.desc(), stats.struct()
        hlog = abi.HostLog(log_capacity, n, line.M) if log_capacity else None
        rc = self._check(self._lib.mjp_run(self._h, C.byref(desc), n, C.byref(rng), C.byref(st),
                                           C.byref(hlog.struct()) if hlog else None),
                         allow=(abi.OK, abi.LOG_TRUNCATED))
        self._line, self._n = line, n
        return RunResult(line, stats, hlog.view().copy() if hlog else None, rc, self.kernel_ms(), self.launch_count(),
                         hlog.first_state if hlog else None)

    # -- device-resident variant (measurement) --------------------------------------
    def upload(self, line: abi.HostLine, n_inst: int):
        desc = line.desc()
        self._check(self._lib.mjp_upload(self._h, C.byref(desc), int(n_inst)))
        self._line, self._n = line, int(n_inst)

    def set_tapes(self, tapes: abi.DrawTapes):
        """Recorded draws for ``mode=abi.RNG_TAPE`` launches of the uploaded line."""
        t = tapes.struct()
        self._check(self._lib.mjp_set_tapes(self._h, C.byref(t)))

    def run_tape(self, line: abi.HostLine, n_inst: int, tapes: abi.DrawTapes, log_capacity: int = 0) -> "RunResult":
        """``run`` with every draw read from ``tapes`` (the reference's own stream, recorded in a JVM run)."""
        self.upload(line, n_inst)
        self.set_tapes(tapes)
        if log_capacity:
            self.reserve_log(log_capacity)
        self.launch(abi.RNG_TAPE, 0, 0, want_log=bool(log_capacity))
        self.sync()
        return self.download(log_capacity=log_capacity)

    def reserve_log(self, capacity: int):
        """Device buffer for SCADA records; needed before ``launch(want_log=True)``."""
        self._check(self._lib.mjp_reserve_log(self._h, int(capacity)))

    def launch(self, mode: int = abi.RNG_COUNTER, seed: int = 0, first_instance_id: int = 0, want_log: bool = False):
        rng = abi.RngSpec(mode, 0, seed, first_instance_id)
        self._check(self._lib.mjp_launch(self._h, C.byref(rng), int(want_log)))

    def launch_slice(self, until: float, resume: bool, mode: int = abi.RNG_COUNTER, seed: int = 0,
                     first_instance_id: int = 0, want_log: bool = False):
        """Advance every instance until its clock passes ``until`` (or the run ends) and park it in HBM;
        ``resume=True`` continues the instances parked by the previous slice."""
        rng = abi.RngSpec(mode, 0, seed, first_instance_id)
        self._check(self._lib.mjp_launch_slice(self._h, C.byref(rng), int(want_log), float(until), int(resume)))

    def sync(self):
        self._check(self._lib.mjp_sync(self._h))

    def download(self, stats: abi.HostStats | None = None, log_capacity: int = 0) -> RunResult:
        stats = stats or abi.HostStats(self._line, self._n)
        st = stats.struct()
        hlog = abi.HostLog(log_capacity, self._n, self._line.M) if log_capacity else None
        rc = self._check(self._lib.mjp_download(self._h, C.byref(st), C.byref(hlog.struct()) if hlog else None),
                         allow=(abi.OK, abi.LOG_TRUNCATED))
        return RunResult(self._line, stats, hlog.view().copy() if hlog else None, rc, self.kernel_ms(),
                         self.launch_count(), hlog.first_state if hlog else None)

    def log_count(self) -> tuple[int, int]:
        """(records captured, records dropped) by the last launch, without copying the records to the host."""
        lo = abi.LogOut(None, (1 << 62), 0, 0, None)
        self._check(self._lib.mjp_download(self._h, None, C.byref(lo)), allow=(abi.OK, abi.LOG_TRUNCATED))
        return int(lo.count), int(lo.dropped)

    def download_fields(self, names) -> abi.HostStats:
        """Copy back only the named planes of the last launch (e.g. ("n_events", "aggregate"))."""
        stats = abi.HostStats(self._line, self._n)
        st = stats.struct()
        for f, _ in abi.StatsOut._fields_:
            if f not in names:
                setattr(st, f, None)
        self._check(self._lib.mjp_download(self._h, C.byref(st), None))
        return stats

    def kernel_ms(self) -> float:
        ms = C.c_float(0)
        self._check(self._lib.mjp_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_int(0)
        self._check(self._lib.mjp_last_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def device_aggregate(self) -> tuple[int, int]:
        """(device address, length in doubles) of the aggregate vector of the last launch; valid after ``sync``."""
        p, n = C.POINTER(C.c_double)(), C.c_int32(0)
        self._check(self._lib.mjp_device_aggregate(self._h, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, int(n.value)

    def aggregate_tensor(self):
        """The aggregate vector of the last launch as a torch CUDA tensor that ALIASES the library's buffer (no
        copy): what a multi-process host hands to ``torch.distributed.all_reduce`` (NCCL) after ``sync``."""
        import torch
        addr, n = self.device_aggregate()

        class _Blob:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (addr, False), "version": 2}
        return torch.as_tensor(_Blob(), device=torch.device("cuda", self.device))

    def probe_end_time(self, kinds, times, clock: float, dur: float) -> float:
        k = np.ascontiguousarray(kinds, np.uint8)
        t = np.ascontiguousarray(times, np.float64)
        out = C.c_double(0)
        self._check(self._lib.mjp_probe_end_time(self._h, abi.ptr(k, C.c_uint8), abi.ptr(t, C.c_double), len(k),
                                                 clock, dur, C.byref(out)))
        return float(out.value)


class MultiHandle:
    """All (or the named) GPUs of one box behind one call: ``mjp_multi_run`` shards the batch into contiguous
    ranges of instance ids, one host thread and one stream per device, and sums the aggregate vectors on the
    devices (peer copies over NVLink, or NCCL).  The fan-out of ``main-loop-multi`` (core.clj:746-766)."""

    def __init__(self, devices=None, event_hash: bool = False, reduce: int = abi.REDUCE_PEER):
        self._lib = abi.load_library()
        if reduce == abi.REDUCE_NCCL:
            # the library binds whatever libnccl.so.2 the process has mapped; in a Python host that is torch's
            # bundled copy, which must be the first one loaded (a second libnccl under the same soname breaks torch)
            import torch  # noqa: F401
        devs = None if devices is None else np.ascontiguousarray(devices, np.int32)
        cfg = abi.MultiConfig(0 if devs is None else len(devs), abi.ptr(devs, C.c_int32),
                              abi.CFG_EVENT_HASH if event_hash else 0, int(reduce))
        self._m = C.c_void_p()
        rc = self._lib.mjp_multi_create(C.byref(cfg), C.byref(self._m))
        if rc != abi.OK:
            raise abi.MjpError(rc, (self._lib.mjp_multi_last_error(None) or b"").decode())
        self.n_devices = int(self._lib.mjp_multi_device_count(self._m))

    def close(self):
        if getattr(self, "_m", None):
            self._lib.mjp_multi_destroy(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def run(self, line: abi.HostLine, n_inst: int, mode: int = abi.RNG_COUNTER, seed: int = 0,
            first_instance_id: int = 0, stats: abi.HostStats | None = None) -> RunResult:
        n = int(n_inst)
        stats = stats or abi.HostStats(line, n)
        rng = abi.RngSpec(mode, 0, seed, first_instance_id)
        desc, st = line.desc(), stats.struct()
        rc = self._lib.mjp_multi_run(self._m, C.byref(desc), n, C.byref(rng), C.byref(st))
        if rc != abi.OK:
            raise abi.MjpError(rc, (self._lib.mjp_multi_last_error(self._m) or b"").decode())
        k, r = C.c_float(0), C.c_float(0)
        self._lib.mjp_multi_last_ms(self._m, C.byref(k), C.byref(r))
        res = RunResult(line, stats, None, rc, float(k.value), 3 * self.n_devices + 1)
        res.reduce_ms = float(r.value)
        return res
