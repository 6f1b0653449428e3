"""Text rendering of the SCADA record stream, as util/log.clj prints it.

  print-lines         util/log.clj:111-137   "{:clk~10,4f :act ~18A~{ ~A~}}~%"
  shorten-msg-floats  util/log.clj:228-234   times re-read from their "~10,4f" rendering
  mjp2pretty-name     util/log.clj:248-261   :m1-start-job, :m2-complete-job, ...
  msg-key-order       util/log.clj:272-273   [:clk :act :m :mjpact :bf :jt :n :ent :ends :j :dets :line]
  detail (:dets)      util/log.clj:11-28     job-per-machine / jobs-per-buffer snapshots, DetsTracker
  pretty-model        util/log.clj:293-303

The device emits compact binary records (mjp_log_record); this module is the host-side
"next" row of SURVEY.md 8(f).  Result / model maps are laid out as clojure.pprint does (pprint_edn), keys in Clojure map order.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi

_PRETTY = {abi.ACT_AJ: "start-job", abi.ACT_SM: "start-job", abi.ACT_BJ: "complete-job", abi.ACT_EJ: "complete-job",
           abi.ACT_BL: "blocked", abi.ACT_UB: "unblocked", abi.ACT_ST: "starved", abi.ACT_US: "unstarved",
           abi.ACT_UP: "up", abi.ACT_DOWN: "down"}


def java_double(x: float) -> str:
    """Double/toString as Clojure prints a double."""
    if math.isnan(x):
        return "##NaN"
    if math.isinf(x):
        return "##Inf" if x > 0 else "##-Inf"
    if x == 0:
        return "0.0"
    a = abs(x)
    if 1e-3 <= a < 1e7:
        s = repr(float(x))
        if "e" in s or "E" in s:
            s = f"{x:.17f}".rstrip("0")
            if s.endswith("."):
                s += "0"
        return s
    m, e = f"{x:.16e}".split("e")
    m = repr(float(m))
    # shortest mantissa that round-trips
    for digits in range(1, 18):
        cand = f"{x:.{digits - 1}e}"
        if float(cand) == x:
            m = cand.split("e")[0]
            break
    if "." not in m:
        m += ".0"
    return f"{m}E{int(e)}"


def shorten(x: float) -> float:
    """(read-string (cl-format nil "~10,4f" x)), util/log.clj:230."""
    return float(f"{x:.4f}")


_NO_DETS = object()


def job_detail_mode(model: dict) -> str:
    """How :dets shows in the printed messages.  log/detail reads [:report :job-detail?] (util/log.clj:14) while
    push-log strips :dets unless the TOP-LEVEL :job-detail? is set (util/log.clj:101): "strip" (no :dets key),
    "nil" (every message prints ``:dets nil``) or "full" (the snapshots)."""
    if not model.get(":job-detail?"):
        return "strip"
    return "full" if (model.get(":report") or {}).get(":job-detail?") else "nil"


class DetsTracker:
    """(log/detail model), util/log.clj:11-28, for the messages of ONE instance: which job sits on which machine and
    in which buffer just BEFORE each logged action.  Jobs pass through a serial line in order, so the state is the
    number of the last job started per machine, which machines hold a job, and the buffer counts.  Nothing before the
    warm-up is in the record stream, so the device reports that state once: ``mjp_log_out.first_state``, taken at the
    END of the tick whose records are the instance's first; the moves of that tick are undone (in reverse execution
    order) to get the state before it, and from there the state is advanced through the job movements in EXECUTION
    order, which ``ord`` restores within a tick (clean-log-buf regroups a tick by machine; a cancelled :bl/:ub or
    :st/:us pair moves no job)."""

    def __init__(self, cl, first_state_row):
        M = len(cl.machines)
        row = [int(x) for x in first_state_row]
        if len(row) != 2 * M or row[0] == 0xFFFFFFFF:
            raise ValueError("first_state: nothing was printed for this instance")
        self.cl, self.M = cl, M
        self.started = row[:M]
        self.occ = [bool(x & 1) for x in row[M:]]
        self.cnt = [(x >> 8) & 0xFF for x in row[M:2 * M - 1]]
        self._rewound = False                                     # first_state is the state AFTER the first tick

    def snapshot(self) -> dict:
        from . import cljhash
        cl, M = self.cl, self.M
        run = {cl.machines[i]: (self.started[i] if self.occ[i] else None) for i in range(M)}            # util/log.clj:18-22
        bufs = {cl.buffers[b]: [self.started[b + 1] + 1 + q for q in range(self.cnt[b])] for b in range(M - 1)}   # :24-28
        order = lambda d: {k: d[k] for k in cljhash.map_order(list(d.keys()))}                            # zipmap: a plain map
        return {":run": order(run), ":bufs": order(bufs)}

    def _apply(self, rec):
        act, m = int(rec["act"]), int(rec["machine"])
        if act in (abi.ACT_AJ, abi.ACT_SM):                       # add-job core.clj:240-254, advance2machine :263-277
            self.started[m], self.occ[m] = int(rec["job"]), True
            if act == abi.ACT_SM:
                self.cnt[m - 1] -= 1
        elif act == abi.ACT_BJ:                                   # advance2buffer core.clj:285-300
            self.occ[m] = False
            self.cnt[m] += 1
        elif act == abi.ACT_EJ:
            self.occ[m] = False

    def _undo(self, rec):
        act, m = int(rec["act"]), int(rec["machine"])
        if act in (abi.ACT_AJ, abi.ACT_SM):                       # job j came after job j-1 on this machine
            self.started[m], self.occ[m] = int(rec["job"]) - 1, False
            if act == abi.ACT_SM:
                self.cnt[m - 1] += 1
        elif act == abi.ACT_BJ:
            self.occ[m] = True
            self.cnt[m] -= 1
        elif act == abi.ACT_EJ:
            self.occ[m] = True

    def tick(self, recs) -> list:
        """:dets of the records of one clock tick (given in cleaned order), in that order."""
        ords = [int(r["ord"]) for r in recs]
        if len(set(ords)) != len(ords):
            raise ValueError("more than 256 messages in one clock tick: the 8-bit execution ordinal wrapped")
        out = [None] * len(recs)
        by_ord = sorted(range(len(recs)), key=lambda q: ords[q])
        if not self._rewound:                                     # the instance's first tick: back to where it began
            for q in reversed(by_ord):
                self._undo(recs[q])
            self._rewound = True
        for q in by_ord:
            out[q] = self.snapshot()
            self._apply(recs[q])
        return out

    def run(self, recs) -> list:
        """:dets for the records of this instance in :line order (any number of whole ticks)."""
        out, a = [], 0
        clk = recs["clk"].view(np.uint64) if len(recs) else []
        for b in range(1, len(recs) + 1):
            if b == len(recs) or clk[b] != clk[a]:
                out.extend(self.tick(recs[a:b]))
                a = b
        return out


def message_map(cl, rec, dets=_NO_DETS) -> dict:
    """One record as the message map push-log hands to print-lines (keys in msg-key-order).  ``dets``: the value
    of :dets when the model keeps it (None prints as nil)."""
    act, m = int(rec["act"]), int(rec["machine"])
    mname = cl.machines[m]
    out = {":clk": shorten(float(rec["clk"])), ":act": f"{mname}-{_PRETTY[act]}", ":m": mname,
           ":mjpact": abi.ACT_KEYWORD[act]}
    if act == abi.ACT_BJ:
        out[":bf"] = cl.buffers[m]
    elif act == abi.ACT_SM:
        out[":bf"] = cl.buffers[m - 1]
    if act == abi.ACT_AJ:
        out[":jt"] = cl.jobtypes[int(rec["n"])]
    if act in (abi.ACT_BJ, abi.ACT_SM):
        out[":n"] = int(rec["n"])
    if act == abi.ACT_EJ:
        out[":ent"] = shorten(float(rec["aux"]))
    if act == abi.ACT_AJ:
        out[":ends"] = shorten(float(rec["aux"]))
    if act in (abi.ACT_AJ, abi.ACT_BJ, abi.ACT_EJ, abi.ACT_SM):
        out[":j"] = int(rec["job"])
    if dets is not _NO_DETS:
        out[":dets"] = dets
    out[":line"] = int(rec["line"])
    return out


def _show(v) -> str:
    if isinstance(v, float):
        return java_double(v)
    if v is None or isinstance(v, dict):
        return _flat(v)                  # ~A of nil / of a map: print-str
    return str(v)


def message_maps(cl, recs, first_state=None, trackers: dict | None = None) -> list:
    """Message maps of records sorted by (instance, line), with :dets as the model's :job-detail? flags say
    (job_detail_mode).  ``first_state``: RunResult.first_state; ``trackers``: {instance: DetsTracker} carried over
    between successive parts of one stream (streaming)."""
    mode = job_detail_mode(cl.model)
    if mode == "strip":
        return [message_map(cl, r) for r in recs]
    if mode == "nil":
        return [message_map(cl, r, None) for r in recs]
    if first_state is None:
        raise ValueError(":job-detail? needs RunResult.first_state")
    trackers = {} if trackers is None else trackers
    out = []
    inst = recs["instance"] if len(recs) else []
    a = 0
    for b in range(1, len(recs) + 1):
        if b == len(recs) or inst[b] != inst[a]:
            g = int(inst[a])
            if g not in trackers:
                trackers[g] = DetsTracker(cl, first_state[g])
            part = recs[a:b]
            out.extend(message_map(cl, r, d) for r, d in zip(part, trackers[g].run(part)))
            a = b
    return out


def format_line(msg: dict) -> str:
    """cl-format "{:clk~10,4f :act ~18A~{ ~A~}}" (util/log.clj:114-125)."""
    rest = " ".join(f"{k} {_show(v)}" for k, v in msg.items() if k not in (":clk", ":act"))
    return "{:clk%10.4f :act %-18s %s}" % (msg[":clk"], msg[":act"], rest)


def render_lines(cl, records: np.ndarray, instance: int | None = None, first_state=None):
    """Text lines of one instance (default: every instance, in instance order), in :line order."""
    recs = np.sort(records, order=["instance", "line"])
    if instance is not None:
        recs = recs[recs["instance"] == instance]
    return [format_line(m) for m in message_maps(cl, recs, first_state)]


def pull_data(cl, records: np.ndarray, n: int | None, instance: int = 0, first_state=None):
    """core/pull-data! (core.clj:845-849) over a captured stream: the first n message maps (all if n is None)."""
    recs = np.sort(records[records["instance"] == instance], order="line")
    return message_maps(cl, recs, first_state)[:n]


# defrecord field lists (core.clj:38-66): a record prints its declared fields first, in this order and nil when unset,
# then whatever else was put into it
RECORD_FIELDS = {"Model": [":line", ":entry-point", ":jobmix", ":params", ":topology", ":print"],
                 "ExpoMachine": [":name", ":lambda", ":mu", ":status", ":W", ":up&down"],
                 "Buffer": [":N", ":holding"], "JobType": [":w"]}
_PRETTY_DROPS = (":name", ":status", ":mchain", ":up&down", ":holding")          # util/log.clj:300


def clj_order(x, force_hash: bool = False):
    """`x` with every map re-keyed in the order Clojure would iterate (and so print) it: records by declared
    fields then extras; plain maps in insertion order up to 8 entries and in hash order beyond (cljhash.py).
    ``force_hash``: the top-level map was a hash map once (it held more than 8 entries) and stayed one."""
    from . import cljhash
    if isinstance(x, dict):
        tag = getattr(x, "tag", None)
        if tag in RECORD_FIELDS:
            basis = RECORD_FIELDS[tag]
            keys = basis + cljhash.map_order([k for k in x if k not in basis])
            return {k: clj_order(x.get(k)) for k in keys}
        keys = list(x.keys())
        if force_hash:
            keys = sorted(keys, key=lambda k: cljhash.trie_key(cljhash.keyword_hash(k)))
        elif all(isinstance(k, str) and k.startswith(":") for k in keys):
            keys = cljhash.map_order(keys)
        return {k: clj_order(x[k]) for k in keys}
    if isinstance(x, (list, tuple)):
        return [clj_order(v) for v in x]
    return x


def pretty_model(model: dict) -> dict:
    """util/log.clj:293-303: drop :name :status :mchain :up&down :holding from the equipment (dissoc of a declared
    field turns the record into a plain map of its remaining fields, in field order), everything in print order."""
    from . import cljhash
    line = model[":line"]
    pretty_line = {}
    for name in cljhash.map_order(list(line.keys())):
        equip = clj_order(line[name]) if getattr(line[name], "tag", None) in RECORD_FIELDS else dict(line[name])
        pretty_line[name] = {a: clj_order(b) for a, b in equip.items() if a not in _PRETTY_DROPS}
    out = clj_order(model)
    out[":line"] = pretty_line
    return out


RESULT_KEYS = (":runtime", ":TP", ":observed-residence-time", ":number-of-jobs", ":blocked", ":starved",
               ":avg-buffer-occupancy", ":bottleneck-machine", ":process-id", ":diag-log-buf", ":status")


def result_block(res: dict) -> str:
    """The ``#_`` result block of one simulation as output-sims / main-loop-once print it (util/log.clj:305-319,
    core.clj:785-786): the reference's own keys (no README aliases, :jobmix dissoc'ed), in the order of the hash map
    the result became when its ninth key was assoc'ed (core.clj:711-718); :TP is a float (core.clj:625);
    *print-length* is 10 (core.clj:750,773)."""
    shown = {k: res[k] for k in RESULT_KEYS if k in res}
    if ":TP" in shown:
        shown[":TP"] = np.float32(shown[":TP"])
    return "#_" + pprint_edn(clj_order(shown, force_hash=True), print_length=10)


RIGHT_MARGIN = 72          # clojure.pprint/*print-right-margin*


def java_float(x) -> str:
    """Float/toString: what Clojure prints for ``(float ...)`` (:TP, core.clj:625)."""
    f = np.float32(x)
    if f == 0:
        return "0.0"
    a = abs(float(f))
    digits = np.format_float_scientific(f, unique=True, trim="0")        # shortest digits that round-trip as a float
    mant, exp = digits.split("e")
    if 1e-3 <= a < 1e7:
        return np.format_float_positional(f, unique=True, trim="0")
    return f"{mant}E{int(exp)}"


def _atom(x) -> str:
    if isinstance(x, bool):
        return "true" if x else "false"
    if x is None:
        return "nil"
    if isinstance(x, np.float32):
        return java_float(x)
    if isinstance(x, (float, np.floating)):
        return java_double(float(x))
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    if isinstance(x, str) and not x.startswith(":") and not getattr(x, "is_symbol", False) and type(x) is str:
        return '"' + x.replace("\\", "\\\\").replace('"', '\\"') + '"'
    return str(x)


class _More:
    """What clojure.pprint writes in place of the elements beyond *print-length*."""

    def __repr__(self):
        return "..."


_MORE = _More()


def _entries(x, limit):
    """(key, value) pairs of a map, cut at *print-length* (the cut shows as a bare ``...``)."""
    items = list(x.items())
    if limit is not None and len(items) > limit:
        return items[:limit] + [(_MORE, _MORE)]
    return items


def _elements(x, limit):
    items = list(x)
    if limit is not None and len(items) > limit:
        return items[:limit] + [_MORE]
    return items


def _flat(x, limit=None) -> str:
    """The one-line rendering (what clojure.pprint produces when everything fits)."""
    if isinstance(x, dict):
        return "{" + ", ".join("..." if k is _MORE else f"{_flat(k, limit)} {_flat(v, limit)}" for k, v in _entries(x, limit)) + "}"
    if isinstance(x, (list, tuple)):
        return "[" + " ".join(_flat(v, limit) for v in _elements(x, limit)) + "]"
    if x is _MORE:
        return "..."
    return _atom(x)


def pprint_edn(x, col: int = 0, trailing: int = 0, print_length: int | None = None) -> str:
    """``clojure.pprint/pprint`` of maps, vectors and atoms, laid out as the reference's ``#_`` blocks are
    (core.clj:774-775,785-786, util/log.clj:305-319).

    Restated from clojure.pprint (pprint-map / pprint-vector, pretty_writer.clj; Clojure 1.9): a map is a logical
    block ``{ ... }`` whose entries are separated by ``", "`` + a :linear conditional newline; every entry is its own
    block ``key`` + space + :linear newline + ``value``; a vector has a :linear newline after every element.  A
    :linear newline of block B is taken iff B already had a line break inside it, or the text from the newline up
    to the next conditional newline of an ENCLOSING block (so: through the end of B and any closing brackets and
    separators that follow) does not fit before column 72 (``column + length < 72``).  Continuation lines of a
    block start in the column after its opening bracket; an entry's value that is pushed to its own line starts in
    the column of the entry's key.  Pinned on the reference's recorded outputs (``data/new-out.clj``,
    ``data/out-10.clj``, ``data/submodel-1-out.clj``): re-printing the parsed files reproduces them byte for byte
    (tests/test_host_model.py).  ``col``: column the value starts in; ``trailing``: characters that follow it
    before the next newline opportunity of an enclosing block; ``print_length``: ``*print-length*`` (the reference
    binds it to 10 around its ``#_`` blocks, core.clj:750,773: longer collections end in ``...``)."""
    return _pp(x, col, trailing, print_length)[0]


def _pp(x, col: int, trailing: int, limit=None):
    """-> (text, a line break was emitted inside)"""
    flat = _flat(x, limit)
    fits = lambda c, n: c + n < RIGHT_MARGIN
    if isinstance(x, dict) and x:
        items = _entries(x, limit)
        indent = col + 1
        out, cur, broke = "{", indent, False
        # flat lengths of the entries still to come (with their separators), for the map-level newlines
        flats = [3 if k is _MORE else len(_flat(k, limit)) + 1 + len(_flat(v, limit)) for k, v in items]
        for idx, (k, v) in enumerate(items):
            last = idx == len(items) - 1
            if k is _MORE:                                               # always the last one
                out += "..."
                break
            ktxt = _flat(k, limit)
            entry_col = cur
            etrail = (1 + trailing) if last else 2                       # "}" + what follows the map, or ", "
            out += ktxt + " "
            cur += len(ktxt) + 1
            if fits(cur, len(_flat(v, limit)) + etrail):                  # the entry's own :linear newline
                vtxt, vbroke = _flat(v, limit), False
                out += vtxt
                cur += len(vtxt)
            else:
                out = out[:-1] + "\n" + " " * entry_col                  # (trailing white space is dropped at a break)
                vtxt, _ = _pp(v, entry_col, etrail, limit)
                out += vtxt
                cur = len(vtxt.rsplit("\n", 1)[-1]) + (entry_col if "\n" not in vtxt else 0)
                broke = True
            if not last:
                out += ","
                rest = sum(flats[idx + 1:]) + 2 * (len(items) - idx - 2) + 1 + trailing   # entries, ", " between, "}" + after
                if broke or not fits(cur + 2, rest):
                    out += "\n" + " " * indent
                    cur = indent
                    broke = True
                else:
                    out += " "
                    cur += 2
        return out + "}", broke
    if isinstance(x, (list, tuple)) and len(x):
        indent = col + 1
        out, cur, broke = "[", indent, False
        x = _elements(x, limit)
        flats = [len(_flat(v, limit)) for v in x]
        for idx, v in enumerate(x):
            last = idx == len(x) - 1
            vtrail = (1 + trailing) if last else 1
            if isinstance(v, (dict, list, tuple)) and not fits(cur, flats[idx] + vtrail):
                vtxt, vb = _pp(v, cur, vtrail, limit)
                broke = broke or vb
            else:
                vtxt = _flat(v, limit)
            out += vtxt
            cur = len(vtxt.rsplit("\n", 1)[-1]) + (cur if "\n" not in vtxt else 0)
            if not last:
                rest = sum(flats[idx + 1:]) + (len(x) - idx - 2) + 1 + trailing
                if broke or not fits(cur + 1, rest):
                    out += "\n" + " " * indent
                    cur = indent
                    broke = True
                else:
                    out += " "
                    cur += 1
        return out + "]", broke
    return flat, False
