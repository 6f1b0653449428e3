#!/usr/bin/env python
"""Warp-instructions of one ncu capture grouped by section of mjp_line_kernel.cuh.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_sections.py sass.csv <mangled kernel> [events]
"""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_by_line as N

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SO = os.path.join(ROOT, "mjpdes_b200", "csrc", "libmjpdes_b200.so")
SRC = os.path.join(ROOT, "mjpdes_b200", "csrc", "mjp_line_kernel.cuh")

def main():
    sass, mangled = sys.argv[1:3]
    events = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(sass)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    ni = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
    lsb = hdr.index("stall_long_sb") if "stall_long_sb" in hdr else None
    inst = [r for r in rows[hi + 1:] if len(r) > te]
    lm = N.line_map(SO, mangled)
    assert len(lm) == len(inst), (len(lm), len(inst), "rebuild mismatch: profile and .so differ")
    src = open(SRC).read().splitlines()
    find = lambda s: next(i + 1 for i, l in enumerate(src) if s in l)
    marks = [("prologue/init", 1), ("loop head + guard masks", find("for (;;) {")),
             ("scan (min candidate time)", find("Machines whose job is already over")),
             ("t*, m_f / m_ud (group scans)", find("// future :ends equal to t*")),
             ("batch masks", find("const MaskT keep = live ?")), ("clock + flush vote", find("// ---- advance-clock")),
             ("AJ section", find("if (__any_sync(am, b_aj != 0))")), ("UP/DOWN section", find("bring-machine-up, then -down")),
             ("TOBUF section", find("// advance2buffer core.clj:285-300")), ("TOMACH section", find("// advance2machine core.clj:263-277")),
             ("record-actions section", find("record-actions core.clj:304-326, in the order")),
             ("loop tail", find("rare: one vote on the common path"))]
    helpers = [("flush_tick", find("auto flush_tick"), find("auto start_job")),
               ("SCADA stage/emit", find("auto stage_msg"), find("// push-log for the tick that just ended")),
               ("byte get/set, params", find("auto getb"), find("double r_end[")),
               ("get_end/set_end", find("double r_end["), find("auto px =")),
               ("next_event_time", find("auto next_event_time"), find("// ---- preprocess-model")),
               ("group minima helpers / rescan_fut", find("constexpr int GSH = "), find("if (fresh_start) {")),
               ("mark_seen/toggle/type_from", find("auto mark_seen"), find("// ---- SCADA capture")),
               ("job-type draw (whole warp)", find("keep at least one drawn job type"), find("// ---- runables")),
               ("up/down generation (whole warp)", find("Pre-generate the event AFTER"), find("bring-machine-up, then -down")),
               ("out-of-line generators", find("sojourn_counter(uint32_t idx"), find("#define MJP_BIT"))]
    agg = {}
    for r, loc in zip(inst, lm):
        f, l, inl = loc
        if f != "mjp_line_kernel.cuh":
            key = "philox / ln / exp (mjp_device.cuh)" if f == "mjp_device.cuh" else "cuda intrinsics headers"
        else:
            key = None
            for name, a, b in helpers:
                if a <= l < b:
                    key = name
            if key is None:
                for name, a in marks:
                    if l >= a:
                        key = name
        x = agg.setdefault(key, [0, 0, 0, 0, 0])
        x[0] += int(r[ie] or 0); x[1] += int(r[te] or 0); x[2] += int(r[sm] or 0)
        if ni is not None: x[3] += int(r[ni] or 0)
        if lsb is not None: x[4] += int(r[lsb] or 0)
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[2] for v in agg.values())
    print(f"{'section':40s} {'% warp-inst':>11s} {'lanes':>6s} {'% samples':>9s} {'no_inst':>8s} {'long_sb':>8s}" + (f" {'inst/event':>10s}" if events else ""))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:40s} {100 * v[0] / ti:10.1f}% {v[1] / max(v[0], 1):6.1f} {100 * v[2] / max(ts, 1):8.1f}% {100 * v[3] / max(ts, 1):7.1f}% {100 * v[4] / max(ts, 1):7.1f}%"
              + (f" {v[0] / events:10.2f}" if events else ""))

if __name__ == "__main__":
    main()
