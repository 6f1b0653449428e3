"""Host-side mirror of the reference's model build and result post-processing.

Mirrors, with the same names and argument meaning:

  map->Model / map->ExpoMachine / map->Buffer / map->JobType   core.clj:38-67
  spec-check-model                                             core.clj:75-121,538-545
  preprocess-model (defaults, :bounds sampling, machine/buffer vectors)   core.clj:547-615
  calc-basics / calc-bneck / analyze-results / postprocess-model          core.clj:617-718
  main-loop (once / multi)                                                core.clj:746-797

What changes is only the middle: ``main-loop-loop`` (core.clj:729) for all replications
at once is one call into the CUDA library.  Maps are plain dicts keyed by keyword
strings (":line", ":m1", ...), which is also what :mod:`mjpdes_b200.edn` produces, so the
reference's own model files load unchanged.
"""
from __future__ import annotations

import math
import time

import numpy as np

from . import abi, cljhash, edn
from .philox import uniform53

DEFAULTS = {":jobs-move-to-down-machines?": False, ":time-format": "~10,4f"}      # core.clj:32-34
DEFAULT_SEED = 0x4D4A50646573
BOUNDS_CONSUMER = 0x40000000          # substreams used for {:dist :uniform :bounds [lo hi]} (core.clj:547-549)


class ModelError(ValueError):
    """ex-info "The model is not well-formed:" (core.clj:544)."""


class Record(dict):
    tag = "Record"

    def __repr__(self):
        return f"(map->{self.tag} {dict.__repr__(self)})"


class Model(Record):
    tag = "Model"


class ExpoMachine(Record):
    tag = "ExpoMachine"


class Buffer(Record):
    tag = "Buffer"


class JobType(Record):
    tag = "JobType"


_CTORS = {"map->Model": Model, "map->ExpoMachine": ExpoMachine, "map->Buffer": Buffer, "map->JobType": JobType}


def _kw(d: dict) -> dict:
    """Accept python-style keys (lambda_, W, N ...) as well as keyword strings."""
    out = {}
    for k, v in d.items():
        k = str(k)
        if not k.startswith(":"):
            k = ":" + k.rstrip("_").replace("_", "-")
        out[k] = v
    return out


def map_to_Model(m: dict) -> Model:
    return Model(_kw(m))


def map_to_ExpoMachine(m: dict | None = None, **kw) -> ExpoMachine:
    return ExpoMachine(_kw({**(m or {}), **kw}))


def map_to_Buffer(m: dict | None = None, **kw) -> Buffer:
    return Buffer(_kw({**(m or {}), **kw}))


def map_to_JobType(m: dict | None = None, **kw) -> JobType:
    return JobType(_kw({**(m or {}), **kw}))


def eval_form(x):
    """What ``(eval form)`` in ns gov.nist.MJPdes.core does to a model file (main.clj:12-14):
    only the four record constructors, optionally ns-qualified, are understood."""
    if isinstance(x, edn.Form):
        if not x:
            return x
        name = str(x[0]).split("/")[-1]
        if name in _CTORS:
            if len(x) != 2 or not isinstance(x[1], dict):
                raise ModelError(f"{name} expects one map")
            return _CTORS[name]({k: eval_form(v) for k, v in x[1].items()})
        raise ModelError(f"cannot evaluate ({x[0]} ...): only map->Model/ExpoMachine/Buffer/JobType are supported")
    if isinstance(x, dict):
        return {k: eval_form(v) for k, v in x.items()}
    if isinstance(x, list):
        return [eval_form(v) for v in x]
    return x


def read_model(path: str) -> Model:
    """``-i <file>`` of the reference CLI (main.clj:8-14): read ONE form, evaluate it."""
    m = eval_form(edn.read_file(path))
    if not isinstance(m, Model):
        raise ModelError(f"{path}: the form is not a (map->Model ...)")
    return m


def _num(x) -> bool:
    return isinstance(x, (int, float)) and not isinstance(x, bool)


def _bounds(x):
    return isinstance(x, dict) and ":bounds" in x and len(x[":bounds"]) == 2 and all(_num(v) for v in x[":bounds"])


def is_machine(e) -> bool:
    return isinstance(e, ExpoMachine) or (isinstance(e, dict) and ":lambda" in e and ":mu" in e)


def is_buffer(e) -> bool:
    return isinstance(e, Buffer) or (isinstance(e, dict) and ":N" in e and ":lambda" not in e)


def spec_check_model(model: dict) -> dict:
    """The clojure.spec checks of core.clj:75-121 that concern the run loop; raises ModelError."""
    def bad(msg):
        raise ModelError("The model is not well-formed: " + msg)
    for k in (":line", ":topology", ":entry-point", ":jobmix", ":params"):      # ::Model core.clj:121
        if k not in model:
            bad(f"missing {k}")
    line, topo = model[":line"], model[":topology"]
    if not isinstance(line, dict) or len(line) < 3:
        bad(":line needs at least 3 pieces of equipment (core.clj:120)")
    if not isinstance(topo, list) or set(topo) != set(line.keys()) or len(topo) != len(line):
        bad(":topology must list every key of :line exactly once")
    for i, name in enumerate(topo):
        e = line[name]
        if i % 2 == 0:
            if not is_machine(e):
                bad(f"{name}: machines and buffers must alternate, starting and ending with a machine")
            if not (_num(e[":lambda"]) and e[":lambda"] >= 0):
                bad(f"{name} :lambda must be non-negative (core.clj:89)")
            if not (_num(e[":mu"]) and e[":mu"] > 0):
                bad(f"{name} :mu must be positive (core.clj:88)")
            W = e.get(":W", 1.0)
            if not ((_num(W) and W > 0) or _bounds(W)):
                bad(f"{name} :W must be positive or {{:dist :uniform :bounds [lo hi]}} (core.clj:87,563)")
            if e.get(":discipline", ":BAS") not in (":BAS", ":BBS"):
                bad(f"{name} :discipline must be :BAS or :BBS (core.clj:91)")
        else:
            if not is_buffer(e):
                bad(f"{name}: machines and buffers must alternate")
            if not (isinstance(e[":N"], int) and e[":N"] > 0):
                bad(f"{name} :N must be a positive integer (core.clj:95)")
    if len(topo) % 2 == 0:
        bad("the line must end with a machine")
    if model[":entry-point"] != topo[0]:
        bad(":entry-point must be the first machine of :topology (the only layout the run loop handles, utils.clj:82-87)")
    jm = model[":jobmix"]
    if not isinstance(jm, dict) or len(jm) < 1:
        bad(":jobmix needs at least one job type (core.clj:117)")
    machines = topo[0::2]
    for jt, v in jm.items():
        if not (_num(v.get(":portion")) and v[":portion"] > 0):
            bad(f"{jt} :portion must be positive (core.clj:109)")
        w = v.get(":w")
        if not isinstance(w, dict):
            bad(f"{jt} :w must be a map (core.clj:108)")
        for m in machines:
            if m not in w or not ((_num(w[m]) and w[m] > 0) or _bounds(w[m])):
                bad(f"{jt} :w {m} must be positive")
    p = model[":params"]
    if not (_num(p.get(":warm-up-time", 0)) and p.get(":warm-up-time", 0) >= 0):
        bad(":warm-up-time must be non-negative (core.clj:104)")
    if not (_num(p.get(":run-to-time")) and p[":run-to-time"] > 0):
        bad(":run-to-time must be positive (core.clj:105)")
    rep = model.get(":report") or {}
    ml = rep.get(":max-lines", 0)
    if not ((isinstance(ml, int) and ml >= 0) or ml == math.inf):
        bad(":max-lines must be a non-negative integer or ##Inf (core.clj:112)")
    return model


class CompiledLine:
    """The preprocessed model in the flat form the C ABI takes, plus the names needed to
    turn results back into the reference's maps."""

    def __init__(self, model, host: abi.HostLine, machines, buffers, jobtypes, n_inst, seed, first_instance_id):
        self.model, self.host = model, host
        self.machines, self.buffers, self.jobtypes = machines, buffers, jobtypes
        self.n_inst, self.seed, self.first_instance_id = n_inst, seed, first_instance_id


def preprocess_model(model: dict, n_inst: int | None = None, seed: int = DEFAULT_SEED, first_instance_id: int = 0,
                     check: bool = True) -> CompiledLine:
    """core.clj:578-615 for all replications at once.

    * :discipline defaults to :BAS (core.clj:556); :jm2dm to false (core.clj:582);
    * an empty :report becomes {:log? false :max-lines 0} (core.clj:586-588);
    * ``W`` / ``w`` given as {:dist :uniform :bounds [lo hi]} are drawn PER REPLICATION
      (rand-range, core.clj:547-549,563,599-608) -> per-instance planes W_inst / w_inst;
      draw = lo + u * (hi - lo) with u from substream BOUNDS_CONSUMER + index of the instance.
    """
    if check:
        spec_check_model(model)
    n = int(n_inst if n_inst is not None else (model.get(":number-of-simulations") or 1))
    line, topo = model[":line"], model[":topology"]
    machines, buffers = topo[0::2], topo[1::2]
    # iteration order of the :jobmix map, which new-job walks (core.clj:332): insertion order up to 8 job types,
    # Clojure hash-map order of the job-type keywords from the 9th on
    jobtypes = cljhash.map_order(list(model[":jobmix"].keys()))
    M, K = len(machines), len(jobtypes)
    ids = np.arange(first_instance_id, first_instance_id + n, dtype=np.uint64)

    def draw(spec, index):
        lo, hi = (float(v) for v in spec[":bounds"])
        return lo + uniform53(seed, ids, BOUNDS_CONSUMER + index, 0) * (hi - lo)

    W = np.ones(M)
    W_inst = None
    for i, m in enumerate(machines):
        v = line[m].get(":W", 1.0)                       # (or (:W mach) 1.0) utils.clj:99
        if _num(v):
            W[i] = float(v)
        else:
            if W_inst is None:
                W_inst = np.tile(W[:, None], (1, n))
            W_inst[i] = draw(v, i)
    if W_inst is not None:
        for i in range(M):
            if _num(line[machines[i]].get(":W", 1.0)):
                W_inst[i] = W[i]
            else:
                W[i] = float(np.mean(line[machines[i]][":W"][":bounds"]))
    w = np.ones((K, M))
    w_inst = None
    for k, jt in enumerate(jobtypes):
        for i, m in enumerate(machines):
            v = model[":jobmix"][jt][":w"][m]
            if _num(v):
                w[k, i] = float(v)
            else:
                if w_inst is None:
                    w_inst = np.zeros((K * M, n))
                    w_inst[:] = np.nan
                w_inst[k * M + i] = draw(v, M + k * M + i)
                w[k, i] = float(np.mean(v[":bounds"]))
    if w_inst is not None:
        for k in range(K):
            for i in range(M):
                if np.isnan(w_inst[k * M + i, 0]):
                    w_inst[k * M + i] = w[k, i]
    rep = dict(model.get(":report") or {})
    if not rep:
        rep = {":log?": False, ":max-lines": 0}          # core.clj:586-588
    params = model[":params"]
    host = abi.HostLine(
        lam=[float(line[m][":lambda"]) for m in machines], mu=[float(line[m][":mu"]) for m in machines], W=W,
        discipline=[1 if line[m].get(":discipline", ":BAS") == ":BBS" else 0 for m in machines],
        N=[int(line[b][":N"]) for b in buffers],
        portion=[float(model[":jobmix"][jt][":portion"]) for jt in jobtypes], w=w,
        warm_up=float(params.get(":warm-up-time", 0)), run_to=float(params[":run-to-time"]),
        run_to_job=int(params[":run-to-job"]) if params.get(":run-to-job") is not None else -1,
        jm2dm=bool(model.get(":jm2dm", DEFAULTS[":jobs-move-to-down-machines?"])),
        log=bool(rep.get(":log?", False)), up_down=bool(rep.get(":up&down?", False)),
        max_lines=rep.get(":max-lines", 0), W_inst=W_inst, w_inst=w_inst,
        # (group-by :m ...) in clean-log-buf iterates in hash-map order of these names once > 8 machines log in a tick
        group_rank=cljhash.hash_rank(machines))
    return CompiledLine(model, host, machines, buffers, jobtypes, n, seed, first_instance_id)


# ---- results ------------------------------------------------------------------------------------

def calc_basics(cl: CompiledLine, stats: abi.HostStats, inst: int) -> dict:
    """core.clj:617-634 on the accumulators of one replication."""
    params = cl.model[":params"]
    warm_up = params.get(":warm-up-time", 0)
    run_time = params[":run-to-time"] - warm_up
    njobs = int(stats.njobs[inst])
    lv = cl.host.level_offsets
    occ = {}
    for b, name in enumerate(cl.buffers):
        acc = 0.0
        for level in range(int(cl.host.N[b]) + 1):
            acc = acc + level * float(stats.occ_time[lv[b] + level, inst])
        occ[name] = acc / run_time
    return {
        ":runtime": run_time,
        ":TP": float(np.float32(njobs / run_time)),                                          # (float ...) core.clj:625
        ":observed-residence-time": ":na" if njobs == 0 else float(stats.residence_sum[inst]) / njobs,
        ":number-of-jobs": njobs,
        ":blocked": {m: float(stats.blocked_time[i, inst]) / run_time for i, m in enumerate(cl.machines)},
        ":starved": {m: float(stats.starved_time[i, inst]) / run_time for i, m in enumerate(cl.machines)},
        ":avg-buffer-occupancy": occ,
    }


def calc_bneck(res: dict, machines: list, faithful: bool = True) -> dict:
    """core.clj:636-661.  The reference reads bl(i) / st(i) as ``(nth (vals (:blocked res)) (dec i))``: the i-th
    VALUE of a map built by ``into {}`` over the machines (core.clj:628-629).  Up to 8 machines that map iterates in
    line order and bl(i) is machine i's figure; from 9 machines on it is a hash map and bl(i) is the figure of the
    i-th machine NAME in hash order, so the reference's answer no longer follows the line.  The
    severity map (``zipmap``, core.clj:650) is walked in the same order to pick the first maximum.
    ``faithful=True`` reproduces exactly that (:bottleneck-machine is what the reference prints); for M > 8 the
    rule as intended -- indexed by position in the line -- is reported beside it as
    ``:bottleneck-machine-by-position``."""
    M = len(machines)

    def rule(order):
        bl = [res[":blocked"][m] for m in order]
        st = [res[":starved"][m] for m in order]
        candidates = [i for i in range(M - 1) if bl[i] < st[i + 1]]
        if len(candidates) == 1:
            return machines[candidates[0]]
        sev = {}
        for i, m in enumerate(machines):
            if i == 0:
                sev[m] = abs(bl[0] - st[1])
            elif i == M - 1:
                sev[m] = abs(bl[M - 2] - st[M - 1])
            else:
                sev[m] = abs(bl[i - 1] - st[i]) + abs(bl[i] - st[i + 1])
        top = max(sev.values())
        return next(m for m in order if sev[m] == top)             # first maximum in the severity map's iteration order

    by_position = rule(machines)
    if M <= 8 or not faithful:
        return {**res, ":bottleneck-machine": by_position}
    return {**res, ":bottleneck-machine": rule(cljhash.map_order(machines)),
            ":bottleneck-machine-by-position": by_position}


def calc_residence_time(res: dict, cl: CompiledLine, inst: int = 0) -> dict:
    """core.clj:663-700 (dormant in the reference: ``#_`` in analyze-results, core.clj:706).  Adds
    ``:computed-residence-time`` per job type from machine efficiencies mu/(lambda+mu), w/W (job-requires,
    utils.clj:95-101), the measured blocked fractions and the occupancy of each machine's feed buffer."""
    host, M = cl.host, cl.host.M
    lam = host.lam_inst[:, inst] if host.lam_inst is not None else host.lam
    mu = host.mu_inst[:, inst] if host.mu_inst is not None else host.mu
    W = host.W_inst[:, inst] if host.W_inst is not None else host.W
    w = host.w_inst[:, inst].reshape(host.K, M) if host.w_inst is not None else host.w
    eff = [float(mu[i]) / (float(lam[i]) + float(mu[i])) for i in range(M)]
    requires = [[float(w[k, i]) / float(W[i]) for i in range(M)] for k in range(host.K)]
    virt = [sum(float(host.portion[k]) * (1.0 / eff[i]) * requires[k][i] for k in range(host.K)) for i in range(M)]
    bl = [res[":blocked"][m] for m in cl.machines]
    computed = {}
    for k, jt in enumerate(cl.jobtypes):
        processing = sum((1.0 / eff[i]) * (1.0 + bl[i]) * requires[k][i] for i in range(M))
        waiting = sum((0.5 + res[":avg-buffer-occupancy"][cl.buffers[i - 1]]) * virt[i] * (1.0 + bl[i])
                      for i in range(1, M))                        # 0 for the entry machine, core.clj:693-694
        computed[jt] = processing + waiting
    return {**res, ":computed-residence-time": computed}


def analyze_results(cl: CompiledLine, stats: abi.HostStats, inst: int) -> dict:
    return calc_bneck(calc_basics(cl, stats, inst), cl.machines)          # core.clj:702-706


def postprocess_model(cl: CompiledLine, stats: abi.HostStats, inst: int, process_id: int, wall_seconds: float,
                      log: np.ndarray | None = None, first_state=None) -> dict:
    """core.clj:708-718, plus the README's key aliases.  With ``:report {:diag-log-buf? true}`` the
    result carries ``:diag-log-buf``: every message print-lines saw, without ``:line`` (util/log.clj:134-135)."""
    r = analyze_results(cl, stats, inst)
    r[":process-id"] = process_id
    if (cl.model.get(":report") or {}).get(":diag-log-buf?"):
        from . import scada
        msgs = scada.pull_data(cl, log, None, instance=inst, first_state=first_state) if log is not None else []
        r[":diag-log-buf"] = [{k: v for k, v in m.items() if k != ":line"} for m in msgs]
    r[":jobmix"] = {jt: dict(cl.model[":jobmix"][jt]) for jt in cl.jobtypes}
    r[":status"] = abi.INST_STATUS_KEYWORD.get(int(stats.status[inst]), ":unknown")
    r[":runtime"] = wall_seconds
    r[":average-buffer-occupancy"] = r[":avg-buffer-occupancy"]
    r[":bneck"] = r[":bottleneck-machine"]
    return r


def main_loop(model: dict, handle=None, out_stream=None, seed: int = DEFAULT_SEED, mode: int = abi.RNG_COUNTER,
              log_capacity: int | None = None):
    """core/main-loop (core.clj:789-797): one result map for a single simulation, a list for
    ``:number-of-simulations`` > 1.  SCADA lines (if ``:report {:log? true}``) and the ``#_``-prefixed
    result maps are written to ``out_stream`` as the reference prints them to ``*out*``.
    """
    from . import engine, scada
    n = model.get(":number-of-simulations") or 1
    if (model.get(":report") or {}).get(":continuous?"):
        return main_loop_continuous(model, handle=handle, seed=seed, mode=mode)          # core.clj:794-795
    start = time.time()
    cl = preprocess_model(model, n_inst=n, seed=seed)
    own = handle is None
    h = handle or engine.Handle(device=0)
    try:
        if cl.host.log:
            fixed = log_capacity is not None
            if not fixed:
                # :max-lines is tested per TICK (util/log.clj:220-226), so an instance can overshoot it by one
                # tick: at most 9 messages per machine plus the entry/exit extras (the device's staging bound)
                ml, tick = cl.host.max_lines, 9 * cl.host.M + 4
                per = (ml + tick) if ml != abi.UINT64_MAX else int(4.0 * cl.host.M * cl.host.run_to) + 1024
                log_capacity = max(1024, min(per * n, 1 << 27))
            while True:
                res = h.run(cl.host, n, mode=mode, seed=seed, log_capacity=log_capacity)
                if res.rc != abi.LOG_TRUNCATED:
                    break
                # never hand back a log with silent holes: the run is deterministic, so repeat it with more room
                if fixed or log_capacity >= (1 << 31):
                    raise abi.MjpError(res.rc, f"SCADA record buffer too small ({log_capacity} records)")
                log_capacity *= 4
        else:
            res = h.run(cl.host, n, mode=mode, seed=seed)
    finally:
        if own:
            h.close()
    wall = time.time() - start
    results = [postprocess_model(cl, res.stats, g, g if n > 1 else 1, wall, log=res.log, first_state=res.first_state)
               for g in range(n)]
    if out_stream is not None:
        out_stream.write("#_" + scada.pprint_edn(scada.pretty_model(model), print_length=10) + "\n")   # core.clj:750-752,773-775
        if res.log is not None:
            for text in scada.render_lines(cl, res.log, first_state=res.first_state):
                out_stream.write(text + "\n")
        for r in results:                                                                 # util/log.clj:305-319
            out_stream.write(scada.result_block(r) + "\n")
    return results[0] if n == 1 else results


# ---- continuous / streaming mode ---------------------------------------------------------------------

class ContinuousModel:
    """core/main-loop-continuous (core.clj:801-818): a model from which log messages can be pulled
    indefinitely.  The reference steps the model one micro-step at a time in a go-loop and hands messages
    over a 30-slot channel; here the device advances the (never-ending) run one time slice at a time with
    ``mjp_launch_slice`` -- the instance state stays parked in HBM between slices -- and the SCADA records of
    each slice are queued on the host.  As in the reference, ``:log?`` is forced true and ``:max-lines``
    infinite (core.clj:812-813), and ``:run-to-time`` does not end the run (one-des-step has no end test).
    """

    def __init__(self, model: dict, handle=None, seed: int = DEFAULT_SEED, mode: int = abi.RNG_COUNTER,
                 slice_time: float | None = None, log_capacity: int = 1 << 20):
        from . import engine
        m = dict(model)
        rep = dict(m.get(":report") or {})
        rep.update({":log?": True, ":max-lines": math.inf})
        rep.pop(":continuous?", None)
        m[":report"] = rep
        self.cl = preprocess_model(m, n_inst=1, seed=seed)
        self.cl.host.run_to = math.inf                       # never :normal-end
        self._own = handle is None
        self.h = handle or engine.Handle(device=0)
        self.seed, self.mode, self.capacity = seed, mode, int(log_capacity)
        # a slice should yield a few hundred messages: ~1.4 events per machine per time unit
        self.dt = float(slice_time) if slice_time else max(1.0, 400.0 / (1.4 * self.cl.host.M))
        self.until, self.started, self.queue = 0.0, False, []
        self._trackers = {}                                  # :job-detail?: the :dets state carried from slice to slice
        self.h.upload(self.cl.host, 1)
        self.h.reserve_log(self.capacity)

    def _advance(self):
        from . import scada
        self.until += self.dt
        self.h.launch_slice(self.until, self.started, mode=self.mode, seed=self.seed, want_log=True)
        self.h.sync()
        self.started = True
        res = self.h.download(log_capacity=self.capacity)
        if res.rc == abi.LOG_TRUNCATED:                      # slice too long for the buffer: cannot happen silently
            raise abi.MjpError(res.rc, "SCADA buffer too small for one slice; lower slice_time")
        st = int(res.stats.status[0])
        if st != abi.INST_PAUSED:
            raise abi.MjpError(abi.ERR_INVALID, f"continuous run stopped with status {abi.INST_STATUS_KEYWORD.get(st, st)}")
        recs = np.sort(res.log, order="line")
        self.queue.extend(scada.message_maps(self.cl, recs, res.first_state, self._trackers))

    def pull_data(self, n: int) -> list:
        """core/pull-data! (core.clj:845-849): the next ``n`` messages."""
        while len(self.queue) < n:
            self._advance()
        out, self.queue = self.queue[:n], self.queue[n:]
        return out

    def close(self):
        if self._own and self.h is not None:
            self.h.close()
            self.h = None


def main_loop_continuous(model: dict, handle=None, seed: int = DEFAULT_SEED, mode: int = abi.RNG_COUNTER,
                         **kw) -> ContinuousModel:
    return ContinuousModel(model, handle=handle, seed=seed, mode=mode, **kw)


def pull_data(continuous_model: ContinuousModel, n: int) -> list:
    """(core/pull-data! continuous-model n)"""
    return continuous_model.pull_data(n)
