#!/usr/bin/env python
"""bench.py -- the reference's headline workload on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): resources/example.clj
(5 machines, 4 buffers, 2 job types, warm-up 2000, run-to 20000) x 100 000 independent
replications PER GPU (weak scaling), COUNTER RNG, seed 0x4D4A50646573, fresh instance ids every
step.  One "step" = one pass of the run loop over that batch.  Metric: simulated events/s
(event = one executed action = one run-action call, core.clj:533-536), whole job.

Printed keys beyond the base contract:
  value     events/s with the line descriptor already resident in HBM (device-event time:
            zero-fill of the accumulators + event-loop kernel + aggregate kernels + NCCL reduce)
  e2e       the same through the blocking C-ABI call mjp_run() with HOST buffers: descriptor
            upload, launch, and download of every per-instance statistic, wall clock
  roofline  event-loop kernel against the SM instruction-issue ceiling (the bound SURVEY 8d names;
            warp-instructions per event measured by ncu, profiles/issue_model.json), plus the
            statistics write-back against measured HBM bandwidth
  cpu_baseline  the CPU oracle (a C++ restatement, NOT the JVM reference: no JVM exists in this
            image) on the box's host cores over a bounded sample of the same workload
--impl reference times that same CPU path as the reference arm.
"""
from __future__ import annotations

import argparse
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


def example_line(run_to: float = 20000.0, warm_up: float = 2000.0):
    from mjpdes_b200 import abi
    return abi.HostLine(lam=[0.1] * 5, mu=[0.9] * 5, W=[1.2, 1.0, 1.1, 1.05, 1.2], discipline=[0] * 5,
                        N=[3, 5, 1, 1], portion=[0.8, 0.2],
                        w=[[1.0, 1.0, 1.0, 1.0, 1.0], [1.0, 2.0, 1.5, 1.0, 1.0]], warm_up=warm_up, run_to=run_to)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
    sample = f"{reps} replications of example.clj per step x {args.steps} steps, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "simulated_events_per_sec", "value": value, "unit": "events/s",
        "replications_per_sec": reps * args.steps / secs,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rng": "COUNTER", "seed": hex(SEED),
                   "note": "CPU oracle restatement of core/main-loop-loop, not the JVM reference (no JVM in image)"},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def issue_model():
    p = os.path.join(ROOT, "profiles", "issue_model.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=N_PER_GPU, help="replications per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    n = args.instances
    line = example_line(run_to=args.run_to, warm_up=args.run_to / 10.0)
    h = engine.Handle(device=local_rank, event_hash=False)
    h.upload(line, n)                                     # inputs resident in HBM before the timed region
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    agg_dev = torch.zeros(abi.aggregate_len(line.M), dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(step_idx: int):
        """one pass over this rank's batch; returns (device ms, events executed)"""
        flush.fill_(step_idx & 0xFF)                       # flush L2 between timed iterations
        barrier()
        first = (step_idx * world + rank) * n              # disjoint instance ids per step and rank
        h.launch(abi.RNG_COUNTER, SEED, first)
        h.sync()
        ms = h.kernel_ms()                                 # CUDA events on the launching stream
        part = h.download_fields(("n_events", "status", "aggregate"))
        # the only collective of the path: reduce the aggregate moments across ranks (NCCL over NVLink)
        agg_dev.copy_(torch.from_numpy(part.aggregate))
        if world > 1:
            e0.record()
            dist.all_reduce(agg_dev)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ok = bool(np.all(part.status == abi.INST_NORMAL_END))
        return ms, float(part.n_events.sum()), ok

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

    # ---- e2e: blocking C-ABI call with host buffers, copies inside the timed region ----
    e2e_ev, e2e_secs = 0.0, 0.0
    for s in range(0 if args.no_e2e else max(1, min(args.steps, 2))):
        barrier()
        t0 = time.perf_counter()
        r = h.run(line, n, mode=abi.RNG_COUNTER, seed=SEED, first_instance_id=(10 ** 6 + s) * world * n + rank * n)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        ee = torch.tensor([float(r.total_events)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ee)
        e2e_secs += float(tt.item()); e2e_ev += float(ee.item())
    hs = abi.HostStats(line, n)
    d2h = sum(a.nbytes for a in (hs.njobs, hs.residence_sum, hs.blocked_time, hs.starved_time, hs.occ_time,
                                 hs.final_clock, hs.current_job, hs.n_events, hs.event_hash, hs.status, hs.aggregate))
    h2d = sum(a.nbytes for a in (line.lam, line.mu, line.W, line.discipline, line.N, line.portion, line.w)) + 64

    if rank == 0:
        value = events_per_step * args.steps / dev_secs
        out = {
            "metric": "simulated_events_per_sec", "value": value, "unit": "events/s",
            "replications_per_sec": n * world * args.steps / dev_secs,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_secs / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.run_to == 20000.0 else f"PROFILING run-to {args.run_to}", "instances_per_gpu": n, "rng": "COUNTER", "seed": hex(SEED),
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "events_per_step": events_per_step, "all_normal_end": ok},
            "clocks": clocks,
            "e2e": {"value": (e2e_ev / e2e_secs) if e2e_secs else None, "unit": "events/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
        }
        im = issue_model()
        f_hz = 1e6 * ((clocks or {}).get("sm_max_mhz") or 1965.0)
        peak = SM_COUNT * SMSP_PER_SM * f_hz / 1e9                  # G warp-instructions/s at max clock
        per_gpu_rate = value / world
        if im:
            ach = per_gpu_rate * im["warp_inst_per_event"] / 1e9
            out["roofline"] = {"bound": "issue", "achieved": ach, "peak": peak, "unit": "Gwarp-inst/s",
                               "frac": ach / peak,
                               "traffic": im.get("dram_bytes_per_event", 0.0) * events_per_step / world,
                               "warp_inst_per_event": im["warp_inst_per_event"],
                               "lane_efficiency": im.get("lane_efficiency"),
                               "peak_source": "148 SMs x 4 SMSPs x clocks.max.sm (SURVEY 8d); warp-inst/event from "
                                              + im.get("source", "profiles/")}
        else:
            out["roofline"] = {"bound": "issue", "achieved": None, "peak": peak, "unit": "Gwarp-inst/s",
                               "frac": None, "traffic": None,
                               "note": "no ncu instruction count committed yet (profiles/issue_model.json)"}
        # statistics write-back against HBM: algorithmic bytes = n x (8(2+2M+L)+24)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
        stat_bytes = n * (8 * (2 + 2 * line.M + line.n_levels) + 24)
        out["roofline_writeback"] = {"bound": "hbm", "algorithmic_bytes_per_launch": stat_bytes,
                                     "achieved": stat_bytes / (dev_secs / args.steps) / 1e9, "peak": hbm,
                                     "unit": "GB/s", "frac": stat_bytes / (dev_secs / args.steps) / 1e9 / hbm,
                                     "note": "write-back is fused into the event loop (RED.ADD.F64 into the output "
                                             "planes); it is not a separate phase, so this fraction is tiny by design"}
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
