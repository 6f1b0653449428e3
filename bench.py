#!/usr/bin/env python
"""bench.py -- the reference's headline workload on N B200s of one node, plus the other named shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-configs]

Headline workload (BASELINE.json configs[1]): resources/example.clj
(5 machines, 4 buffers, 2 job types, warm-up 2000, run-to 20000) x 100 000 independent
replications PER GPU (weak scaling), COUNTER RNG, seed 0x4D4A50646573, fresh instance ids every
step.  One "step" = one pass of the run loop over that batch.  Metric: simulated events/s
(event = one executed action = one run-action call, core.clj:533-536), whole job.

Printed keys beyond the base contract:
  value     events/s with the line descriptor already resident in HBM (device-event time:
            zero-fill of the accumulators + event-loop kernel + aggregate kernels + NCCL reduce of the
            device-resident aggregate vector)
  e2e       the same through the blocking C-ABI call mjp_run() with HOST buffers: descriptor
            upload, launch, and download of every per-instance statistic, wall clock
  roofline  event-loop kernel against the SM instruction-issue ceiling (the bound SURVEY 8d names;
            warp-instructions per event measured by ncu, profiles/issue_model.json -- `stale` says whether
            that capture was taken from the kernel source now in the tree), plus the statistics
            write-back against measured HBM bandwidth
  parity_spot  the first 256 instances of the LAST timed step, re-run on the CPU oracle outside the timed
            region and compared bit for bit (the benchmarked kernel instantiation, at the benchmarked size)
  configs   time-boxed runs of the other named shapes, each with its own roofline:
            C3  the 10-machine sweep, STRONG scaling: the 1M grid points divided over the N GPUs
            C4  20 machines, mixed BAS/BBS, 8 job types, 10 000 instances per GPU: statistics only (full
                horizon; also with 75 776 instances, which fill the GPU) and with SCADA capture over the window
                2000..3000 (record bytes/s over the window)
            C5  50 machines, 16 job types, 10M/8 = 1.25M resident instances per GPU, each advanced by a
                sim-time quantum through mjp_launch_slice with its state parked in HBM (weak scaling)
  cpu_baseline  the CPU oracle (a C++ restatement, NOT the JVM reference: no JVM exists in this
            image) on the box's host cores over a bounded sample of the same workload
--impl reference times that same CPU path as the reference arm.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x4D4A50646573
N_PER_GPU = 100_000
WORKLOAD = "example.clj x 100000 replications per GPU (5 machines, 4 buffers, 2 job types, warm-up 2000, run-to 20000)"
SM_COUNT, SMSP_PER_SM = 148, 4
KERNEL_SOURCES = ("mjp_line_kernel.cuh", "mjp_device.cuh", "mjp_params.h", "mjp_plan.h")


def example_line(run_to: float = 20000.0, warm_up: float = 2000.0):
    from mjpdes_b200 import workloads
    return workloads.example_line(warm_up=warm_up, run_to=run_to)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def kernel_source_hash() -> str:
    """sha256 over the sources the event-loop kernels are compiled from; stored in every
    profiles/issue_model*.json so that a capture of an older kernel is flagged, not silently reused."""
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "mjpdes_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def mangled_line_kernel(demangled: str) -> str | None:
    """'void mjp_line_kernel<unsigned int, 5, 0, 0, 0, 0, 0>(KParams)' (ncu's kernel name) -> its symbol."""
    import re
    m = re.search(r"mjp_line_kernel<unsigned (int|long), (\d+), (\d), (\d), (\d), (\d), (\d)>", demangled or "")
    if not m:
        return None
    mask = "j" if m.group(1) == "int" else "m"
    return (f"_ZN3mjp15mjp_line_kernelI{mask}Li{m.group(2)}E" + "".join(f"Lb{b}E" for b in m.groups()[2:]) +
            "EEvNS_7KParamsE")


def kernel_sass_hash(demangled: str, so: str | None = None) -> str | None:
    """sha256 over the instruction sequence (cuobjdump -sass) of ONE event-loop kernel in the built library: what an
    ncu capture was really taken from.  Opcodes, modifiers, predicates' sense, immediates and branch targets count;
    register NUMBERS do not (ptxas renames registers when unrelated source lines move, same instruction stream).  A
    source change that leaves this kernel's code as it was -- e.g. inside a branch that only other template
    instantiations compile -- does not make its issue model stale; any change of its instruction stream does.
    None when cuobjdump or the kernel is not available (the source hash then decides)."""
    import re
    sym = mangled_line_kernel(demangled)
    so = so or os.path.join(ROOT, "mjpdes_b200", "csrc", "libmjpdes_b200.so")
    if sym is None or not os.path.exists(so):
        return None
    try:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, so], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                             text=True, timeout=120).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    code = []
    for ln in out.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)              # "/*0a10*/   IMAD.MOV.U32 R3, RZ, RZ, R5 ;  /* encoding */"
        if m:
            code.append(re.sub(r"\b(U?)(R|P)\d+\b", r"\1\2", m.group(1)).strip())
    if len(code) < 100:                              # not found (cuobjdump only warns) or not a kernel
        return None
    return hashlib.sha256("\n".join(code).encode()).hexdigest()


def issue_model(name: str = ""):
    p = os.path.join(ROOT, "profiles", f"issue_model{'_' + name if name else ''}.json")
    if not os.path.exists(p):
        return None
    with open(p) as fh:
        m = json.load(fh)
    now = kernel_sass_hash(m.get("kernel", "")) if m.get("kernel_sass_sha256") else None
    if now is not None:
        m["stale"], m["stale_basis"] = m["kernel_sass_sha256"] != now, "kernel machine code"
    else:
        m["stale"], m["stale_basis"] = m.get("kernel_source_sha256") != kernel_source_hash(), "kernel sources"
    return m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(n_reps: int, threads: int, first_id: int):
    """events/s of the CPU oracle over n_reps replications of the workload on `threads` host threads."""
    from mjpdes_b200 import abi
    from oracle import binding as oracle
    line = example_line()
    t0 = time.perf_counter()
    r = oracle.run(line, n_reps, mode=abi.RNG_COUNTER, seed=SEED, first_instance_id=first_id, n_threads=threads)
    dt = time.perf_counter() - t0
    return float(r.stats.n_events.sum()) / dt, dt, int(r.stats.n_events.sum())


def headline_config(extra=None) -> dict:
    """`config` of the JSON line: the same dict for both arms (the driver compares them)."""
    c = {"workload": WORKLOAD, "instances_per_gpu": N_PER_GPU, "rng": "COUNTER", "seed": hex(SEED)}
    c.update(extra or {})
    return c


def run_reference(args, rank: int, world: int):
    """Reference arm: the reference's CPU implementation of the path.  The real thing is Clojure on a JVM
    and cannot run here (no java/lein in the image); the stand-in is the oracle restatement
    on all host cores.  Rank 0 alone works."""
    if rank != 0:
        return
    cores = host_cores()
    from oracle import binding as oracle
    oracle.build()
    reps = 64 * cores                                   # bounded sample per step (~2-4 s)
    for w in range(args.warmup):
        cpu_oracle_rate(max(cores, reps // 8), cores, 10 ** 9 + w * reps)
    ev, secs = 0, 0.0
    for s in range(args.steps):
        rate, dt, n_ev = cpu_oracle_rate(reps, cores, 2 * 10 ** 9 + s * reps)
        ev += n_ev; secs += dt
    value = ev / secs
    sample = (f"bounded sample of the workload: {reps} replications of example.clj per step x {args.steps} steps, "
              f"{cores} threads; CPU oracle restatement of core/main-loop-loop, not the JVM reference (no JVM in image)")
    print(json.dumps({
        "impl": "reference", "metric": "simulated_events_per_sec", "value": value, "unit": "events/s",
        "replications_per_sec": reps * args.steps / secs,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(),
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---- the other named shapes ---------------------------------------------------------

class Cluster:
    """barrier + max-over-ranks / sum-over-ranks of scalars (NCCL when world > 1)."""

    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, x: float, op: str) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())


def issue_roofline(name: str, events_per_sec_per_gpu: float, peak_gwis: float) -> dict:
    im = issue_model(name)
    if not im:
        return {"bound": "issue", "achieved": None, "peak": peak_gwis, "unit": "Gwarp-inst/s", "frac": None,
                "note": f"no ncu capture committed (profiles/issue_model_{name}.json)"}
    ach = events_per_sec_per_gpu * im["warp_inst_per_event"] / 1e9
    return {"bound": "issue", "achieved": ach, "peak": peak_gwis, "unit": "Gwarp-inst/s", "frac": ach / peak_gwis,
            "warp_inst_per_event": im["warp_inst_per_event"], "lane_efficiency": im.get("lane_efficiency"),
            "issue_slot_utilisation_in_capture": im.get("issue_slot_utilisation_in_capture"),
            "stale": im["stale"], "stale_basis": im["stale_basis"], "source": im.get("source")}


def bench_configs(args, cl: Cluster, h, peak_gwis: float, hbm_gbs: float) -> dict:
    """Time-boxed runs of C3 / C4 / C5 on this rank's shard; every number is max-over-ranks device time."""
    from mjpdes_b200 import abi, dist as mdist, workloads as wl
    out = {}
    rank, world = cl.rank, cl.world

    def timed_launch(repeat: int = 1, **kw):
        """device ms of one launch (the best of `repeat`: the cold layouts' DRAM behaviour varies by a few % from
        launch to launch with where the scratch happens to be allocated)"""
        best = None
        for _ in range(repeat):
            cl.barrier()
            h.launch(abi.RNG_COUNTER, SEED, **kw)
            h.sync()
            best = h.kernel_ms() if best is None else min(best, h.kernel_ms())
        return best

    def finish(name, ms, events, extra):
        ms_max = cl.reduce(ms, "max")
        ev = cl.reduce(events, "sum")
        rate = ev / (ms_max / 1e3)
        r = {"events": ev, "ms": ms_max, "events_per_sec": rate, "n_gpus": world}
        r.update(extra)
        r["roofline"] = issue_roofline(name, rate / world, peak_gwis)
        out[name] = r

    # -- C3: 10-machine sweep, per-instance W / N planes; STRONG scaling: the grid is divided over the ranks
    total = args.c3_points
    lo, hi = mdist.shard_range(total, world, rank)
    h.upload(wl.sweep10_line(4096, first=lo), 4096)                 # warm the variant (module load, attributes)
    h.launch(abi.RNG_COUNTER, SEED, lo); h.sync()
    h.upload(wl.sweep10_line(hi - lo, first=lo), hi - lo)
    ms = timed_launch(first_instance_id=lo)
    st = h.download_fields(("n_events", "status"))
    finish("C3_sweep10", ms, float(st.n_events.sum()),
           {"scaling": "strong", "points_total": total, "points_per_gpu": hi - lo,
            "all_normal_end": bool(np.all(st.status == abi.INST_NORMAL_END)),
            "shape": "M=10 K=1, per-instance W/N planes, warm-up 2000, run-to 20000, one replication per grid point"})

    # -- C4: 20 machines, mixed BAS/BBS, 8 job types, 10 000 instances per GPU
    n4 = args.c4_instances
    first4 = rank * n4
    h.upload(wl.mixed20_line(log=False), n4)
    timed_launch(first_instance_id=first4)                          # warm-up launch of the variant
    ms = timed_launch(2, first_instance_id=first4)
    st = h.download_fields(("n_events", "status"))
    finish("C4_mixed20_stats", ms, float(st.n_events.sum()),
           {"scaling": "weak", "instances_per_gpu": n4, "all_normal_end": bool(np.all(st.status == abi.INST_NORMAL_END)),
            "shape": "M=20 mixed BAS/BBS K=8, statistics only, warm-up 2000, run-to 20000"})
    # the same line with enough instances to fill the GPU (10 000 thread-per-instance lanes are 68 per SM)
    n4s = args.c4_saturated
    h.upload(wl.mixed20_line(log=False), n4s)
    ms = timed_launch(2, first_instance_id=10 ** 7 + rank * n4s)
    st = h.download_fields(("n_events", "status"))
    finish("C4_mixed20_stats_saturated", ms, float(st.n_events.sum()),
           {"scaling": "weak", "instances_per_gpu": n4s, "all_normal_end": bool(np.all(st.status == abi.INST_NORMAL_END)),
            "shape": "M=20 mixed BAS/BBS K=8, statistics only, warm-up 2000, run-to 20000; instance count that fills one GPU"})
    # SCADA capture over a window after the warm-up (C4's full horizon would be 173 GB of records per 10 000 instances):
    # window time = (run to the end of the window with capture) - (the capture-free run to 2000 of the same instances)
    def scada_window(name, mk_line, n, first, t_end, records_per_instance, model, note):
        h.upload(mk_line(2000.001, False), n)
        timed_launch(first_instance_id=first)
        ms_pre = timed_launch(2, first_instance_id=first)
        ev_pre = float(h.download_fields(("n_events",)).n_events.sum())
        h.upload(mk_line(t_end, True), n)
        h.reserve_log(int(n * records_per_instance))
        timed_launch(first_instance_id=first, want_log=True)
        ms_log = timed_launch(2, first_instance_id=first, want_log=True)
        ev_log = float(h.download_fields(("n_events",)).n_events.sum())
        records, dropped = h.log_count()
        ms_win = cl.reduce(ms_log - ms_pre, "max")
        rec = cl.reduce(float(records), "sum")
        ev_win = cl.reduce(ev_log - ev_pre, "sum")
        ev_all = cl.reduce(ev_log, "sum")
        bps = 32.0 * rec / (ms_win / 1e3)
        out[name] = {
            "scaling": "weak", "instances_per_gpu": n, "n_gpus": world, "window": f"sim time 2000..{t_end:g}, :log? true :up&down? true",
            "note": note, "records": rec, "dropped": cl.reduce(float(dropped), "sum"), "window_ms": ms_win,
            "events_in_window": ev_win, "events_per_sec": ev_win / (ms_win / 1e3), "records_per_sec": rec / (ms_win / 1e3),
            "record_bytes_per_sec": bps,
            "events_per_sec_whole_run": ev_all / (cl.reduce(ms_log, "max") / 1e3),
            "ms_with_capture": cl.reduce(ms_log, "max"), "ms_capture_free_0_to_2000": cl.reduce(ms_pre, "max"),
            "roofline": issue_roofline(model, ev_win / (ms_win / 1e3) / world, peak_gwis),
            "roofline_writeback": {"bound": "hbm", "achieved": bps / 1e9 / world, "peak": hbm_gbs, "unit": "GB/s",
                                   "frac": bps / 1e9 / world / hbm_gbs, "algorithmic_bytes": 32.0 * rec / world,
                                   "note": "32-byte records over the capture window only, per GPU"}}

    c4 = lambda t_end, log: wl.mixed20_line(warm_up=2000.0, run_to=t_end, log=log, up_down=True)
    scada_window("C4_mixed20_scada", c4, n4, first4, 3000.0, 40_000, "C4_mixed20_scada",
                 "the named configuration: 10 000 instances (68 lanes per SM: latency-bound)")
    scada_window("C4_mixed20_scada_saturated", c4, n4s, 10 ** 7 + rank * n4s, 2300.0, 12_500, "C4_mixed20_scada_saturated",
                 "enough instances to fill the GPU; shorter window so that the record stream (25 GB) stays in HBM")
    ex = lambda t_end, log: wl.example_line(warm_up=2000.0, run_to=t_end, log=log, up_down=True, max_lines=abi.UINT64_MAX)
    scada_window("example_scada", ex, args.instances, 2 * 10 ** 7 + rank * args.instances, 3000.0, 12_500, "example_scada",
                 "resources/example.clj with :log? true :up&down? true, 100 000 instances per GPU")

    # -- C5: 50 machines, 16 job types: every instance resident, advanced by one sim-time quantum per launch
    n5 = args.c5_instances
    first5 = rank * n5
    line5 = wl.long50_line(warm_up=1e5, run_to=1e6)
    h.upload(line5, n5)
    q0, q1 = args.c5_quantum0, args.c5_quantum0 + args.c5_quantum
    cl.barrier()
    h.launch_slice(q0, False, seed=SEED, first_instance_id=first5); h.sync()       # start-up transient, untimed
    ev0 = float(h.download_fields(("n_events",)).n_events.sum())
    cl.barrier()
    h.launch_slice(q1, True, seed=SEED, first_instance_id=first5); h.sync()        # the timed quantum
    ms = h.kernel_ms()
    st = h.download_fields(("n_events", "status"))
    finish("C5_long50_slice", ms, float(st.n_events.sum()) - ev0,
           {"scaling": "weak", "instances_per_gpu": n5, "quantum": f"sim time {q0}..{q1} of run-to 1e6, state parked in HBM "
            "between launches (mjp_launch_slice)", "all_paused": bool(np.all(st.status == abi.INST_PAUSED)),
            "shape": "M=50 K=16, aggregate-only, 10M replications / 8 GPUs resident per GPU",
            "extrapolated_full_run_hours": None})
    r5 = out["C5_long50_slice"]
    # the literal C5 (10M replications to run-to 1e6) is ~7e14 events: state the extrapolation, separately
    r5["extrapolated_full_run_hours"] = (7e7 * n5 * world) / r5["events_per_sec"] / 3600.0
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=N_PER_GPU, help="replications per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="headline only (no C3/C4/C5 runs)")
    ap.add_argument("--c3-points", type=int, default=1_000_000, help="grid points of the C3 sweep (whole job)")
    ap.add_argument("--c4-instances", type=int, default=10_000, help="C4 instances per GPU")
    ap.add_argument("--c4-saturated", type=int, default=75_776, help="C4 instances per GPU of the saturated run")
    ap.add_argument("--c5-instances", type=int, default=1_250_000, help="C5 resident instances per GPU")
    ap.add_argument("--c5-quantum0", type=float, default=20.0)
    ap.add_argument("--c5-quantum", type=float, default=60.0, help="sim time of the timed C5 slice")
    ap.add_argument("--run-to", type=float, default=20000.0, help="PROFILING ONLY: shorter horizon (invalid as a bench number)")
    ap.add_argument("--no-e2e", action="store_true", help="PROFILING ONLY: skip the end-to-end leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from mjpdes_b200 import abi, engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cl = Cluster(rank, world, dev)
    n = args.instances
    line = example_line(run_to=args.run_to, warm_up=args.run_to / 10.0)
    h = engine.Handle(device=local_rank, event_hash=False)
    h.upload(line, n)                                     # inputs resident in HBM before the timed region
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier = cl.barrier

    def one_step(step_idx: int):
        """one pass over this rank's batch; returns (device ms, events executed, all ended normally, first id)"""
        flush.fill_(step_idx & 0xFF)                       # flush L2 between timed iterations
        barrier()
        first = (step_idx * world + rank) * n              # disjoint instance ids per step and rank
        h.launch(abi.RNG_COUNTER, SEED, first)
        h.sync()
        ms = h.kernel_ms()                                 # CUDA events on the launching stream
        # the only collective of the path: the aggregate moments, reduced where they are (device-resident vector,
        # NCCL over NVLink) -- no host round trip
        agg = h.aggregate_tensor()
        if world > 1:
            e0.record()
            dist.all_reduce(agg)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        part = h.download_fields(("n_events", "status"))
        ok = bool(np.all(part.status == abi.INST_NORMAL_END)) and float(agg[0].item()) == float(n * world)
        return ms, float(part.n_events.sum()), ok, first

    for w in range(args.warmup):
        one_step(w)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    steps = [one_step(args.warmup + s) for s in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = h.launch_count() * args.steps
    ok = all(x[2] for x in steps)
    t = torch.tensor([x[0] for x in steps], dtype=torch.float64, device=dev)
    evt = torch.tensor([sum(x[1] for x in steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max over ranks, per step
        dist.all_reduce(evt)                               # whole-job events
    dev_secs = float(t.sum().item()) / 1e3
    events_per_step = float(evt.item()) / args.steps

    # ---- parity_spot (outside the timed region): the first 256 instances of the LAST timed step vs the oracle ----
    parity_spot = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import binding as oracle
        oracle.build()
        k = min(256, n)
        got = h.download().stats
        want = oracle.run(line, k, mode=abi.RNG_COUNTER, seed=SEED, first_instance_id=steps[-1][3],
                          n_threads=host_cores()).stats
        parity_spot = all(np.array_equal(getattr(got, f)[..., :k], getattr(want, f)) for f in
                          ("status", "n_events", "current_job", "njobs", "final_clock", "residence_sum",
                           "blocked_time", "starved_time", "occ_time"))

    # ---- e2e: blocking C-ABI call with host buffers, copies inside the timed region ----
    e2e_ev, e2e_secs = 0.0, 0.0
    for s in range(0 if args.no_e2e else max(1, min(args.steps, 2))):
        barrier()
        t0 = time.perf_counter()
        r = h.run(line, n, mode=abi.RNG_COUNTER, seed=SEED, first_instance_id=(10 ** 6 + s) * world * n + rank * n)
        dt = time.perf_counter() - t0
        e2e_secs += cl.reduce(dt, "max"); e2e_ev += cl.reduce(float(r.total_events), "sum")
    hs = abi.HostStats(line, n)
    d2h = sum(a.nbytes for a in (hs.njobs, hs.residence_sum, hs.blocked_time, hs.starved_time, hs.occ_time,
                                 hs.final_clock, hs.current_job, hs.n_events, hs.event_hash, hs.status, hs.aggregate))
    h2d = sum(a.nbytes for a in (line.lam, line.mu, line.W, line.discipline, line.N, line.portion, line.w)) + 64

    f_hz = 1e6 * ((clocks or {}).get("sm_max_mhz") or 1965.0)
    peak = SM_COUNT * SMSP_PER_SM * f_hz / 1e9                  # G warp-instructions/s at max clock
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    configs = None
    if not args.skip_configs:
        configs = bench_configs(args, cl, h, peak, hbm)

    if rank == 0:
        value = events_per_step * args.steps / dev_secs
        out = {
            "metric": "simulated_events_per_sec", "value": value, "unit": "events/s",
            "replications_per_sec": n * world * args.steps / dev_secs,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_secs / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": headline_config({"workload": WORKLOAD if args.run_to == 20000.0 else f"PROFILING run-to {args.run_to}",
                                       "instances_per_gpu": n,
                                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                                       "events_per_step": events_per_step, "all_normal_end": ok}),
            "clocks": clocks,
            "e2e": {"value": (e2e_ev / e2e_secs) if e2e_secs else None, "unit": "events/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
            "parity_spot": parity_spot,
        }
        im = issue_model()
        per_gpu_rate = value / world
        if im:
            ach = per_gpu_rate * im["warp_inst_per_event"] / 1e9
            out["roofline"] = {"bound": "issue", "achieved": ach, "peak": peak, "unit": "Gwarp-inst/s",
                               "frac": ach / peak,
                               # DRAM bytes of ONE launch of the bench configuration itself, measured by ncu
                               # (dram__bytes_read.sum + dram__bytes_write.sum); null if that capture is missing
                               "traffic": im.get("dram_bytes_per_launch_bench_config"),
                               "stale": im["stale"], "stale_basis": im["stale_basis"],
                               "warp_inst_per_event": im["warp_inst_per_event"],
                               "lane_efficiency": im.get("lane_efficiency"),
                               "peak_source": "148 SMs x 4 SMSPs x clocks.max.sm (SURVEY 8d); warp-inst/event from "
                                              + im.get("source", "profiles/")}
        else:
            out["roofline"] = {"bound": "issue", "achieved": None, "peak": peak, "unit": "Gwarp-inst/s",
                               "frac": None, "traffic": None,
                               "note": "no ncu instruction count committed yet (profiles/issue_model.json)"}
        # statistics write-back against HBM: algorithmic bytes = n x (8(2+2M+L)+24)
        stat_bytes = n * (8 * (2 + 2 * line.M + line.n_levels) + 24)
        out["roofline_writeback"] = {"bound": "hbm", "algorithmic_bytes_per_launch": stat_bytes,
                                     "achieved": stat_bytes / (dev_secs / args.steps) / 1e9, "peak": hbm,
                                     "unit": "GB/s", "frac": stat_bytes / (dev_secs / args.steps) / 1e9 / hbm,
                                     "note": "write-back is fused into the event loop (RED.ADD.F64 into the output "
                                             "planes); it is not a separate phase, so this fraction is tiny by design"}
        if configs is not None:
            out["configs"] = configs
        if not args.no_cpu_baseline:
            cores = host_cores()
            from oracle import binding as oracle
            oracle.build()
            reps = 192 * cores
            rate, dt, n_ev = cpu_oracle_rate(reps, cores, 3 * 10 ** 9)
            out["cpu_baseline"] = {"value": rate, "unit": "events/s", "cores": cores, "kind": "port",
                                   "sample": f"{reps} replications of example.clj ({n_ev} events, {dt:.1f} s), "
                                             f"CPU oracle on {cores} threads; not the JVM reference"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    h.close()


if __name__ == "__main__":
    main()
