"""Text rendering of the SCADA record stream, as util/log.clj prints it.

  print-lines         util/log.clj:111-137   "{:clk~10,4f :act ~18A~{ ~A~}}~%"
  shorten-msg-floats  util/log.clj:228-234   times re-read from their "~10,4f" rendering
  mjp2pretty-name     util/log.clj:248-261   :m1-start-job, :m2-complete-job, ...
  msg-key-order       util/log.clj:272-273   [:clk :act :m :mjpact :bf :jt :n :ent :ends :j :dets :line]
  pretty-model        util/log.clj:293-303

The device emits compact binary records (mjp_log_record); this module is the host-side
"next" row of SURVEY.md 8(f).  Result / model maps are printed as EDN with one key per line;
the reference's clojure.pprint line-breaking is not reproduced byte for byte.
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


def message_map(cl, rec) -> dict:
    """One record as the message map push-log hands to print-lines (keys in msg-key-order)."""
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
    out[":line"] = int(rec["line"])
    return out


def _show(v) -> str:
    if isinstance(v, float):
        return java_double(v)
    return str(v)


def format_line(msg: dict) -> str:
    """cl-format "{:clk~10,4f :act ~18A~{ ~A~}}" (util/log.clj:114-125)."""
    rest = " ".join(f"{k} {_show(v)}" for k, v in msg.items() if k not in (":clk", ":act"))
    return "{:clk%10.4f :act %-18s %s}" % (msg[":clk"], msg[":act"], rest)


def render_lines(cl, records: np.ndarray, instance: int | None = None):
    """Text lines of one instance (default: every instance, in instance order), in :line order."""
    recs = np.sort(records, order=["instance", "line"])
    if instance is not None:
        recs = recs[recs["instance"] == instance]
    return [format_line(message_map(cl, r)) for r in recs]


def pull_data(cl, records: np.ndarray, n: int | None, instance: int = 0):
    """core/pull-data! (core.clj:845-849) over a captured stream: the first n message maps (all if n is None)."""
    recs = np.sort(records[records["instance"] == instance], order="line")[:n]
    return [message_map(cl, r) for r in recs]


def pretty_model(model: dict) -> dict:
    """util/log.clj:293-303: drop :name :status :mchain :up&down :holding from the equipment."""
    out = dict(model)
    out[":line"] = {k: {a: b for a, b in dict(v).items() if a not in (":name", ":status", ":mchain", ":up&down", ":holding")}
                    for k, v in model[":line"].items()}
    return out


def pprint_edn(x, indent: int = 0) -> str:
    pad = " " * (indent + 1)
    if isinstance(x, dict):
        if not x:
            return "{}"
        items = [f"{k} {pprint_edn(v, indent + len(str(k)) + 2)}" for k, v in x.items()]
        return "{" + (",\n" + pad).join(items) + "}"
    if isinstance(x, (list, tuple)):
        return "[" + " ".join(pprint_edn(v, indent + 1) for v in x) + "]"
    if isinstance(x, bool):
        return "true" if x else "false"
    if x is None:
        return "nil"
    if isinstance(x, float):
        return java_double(x)
    return str(x)
