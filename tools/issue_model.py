#!/usr/bin/env python
"""profiles/issue_model[_<name>].json from one ncu report + the event count of the profiled launch:

    python tools/issue_model.py gpurun_out/<name>.ncu-rep gpurun_out/<name>_events.json [--name <name>] \\
        [--bench-dram-csv gpurun_out/dram.csv] [--summary profiles/r02_<name>_ncu_summary.txt]

The JSON carries the sha256 of the kernel sources (bench.kernel_source_hash): bench.py prints
roofline.stale = true when the capture was not taken from the kernel that is in the tree.
"""
import argparse, csv, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(m, k):
    return float(m[k][0].replace(",", ""))


def to_bytes(m, k):
    v, u = m[k]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("events_json")
    ap.add_argument("--name", default="")
    ap.add_argument("--bench-dram-csv", help="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv log of the bench configuration")
    ap.add_argument("--summary", help="also write tools/ncu_summary.py output here")
    a = ap.parse_args()
    ev = json.load(open(a.events_json))
    m = raw_metrics(a.rep)
    wi = num(m, "smsp__inst_executed.sum")
    lanes = num(m, "smsp__thread_inst_executed_per_inst_executed.ratio")
    ce = num(m, "smsp__cycles_elapsed.sum")
    dram = to_bytes(m, "dram__bytes_read.sum") + to_bytes(m, "dram__bytes_write.sum")
    out = {
        "source": f"{a.summary or a.rep} (ncu --set full --clock-control none, {m['Kernel Name'][0]}, "
                  f"{ev['instances']} instances, workload {ev['name']}, B200)",
        "kernel": m["Kernel Name"][0], "workload": ev["name"], "events_in_capture": ev["events"],
        "warp_inst_per_event": wi / ev["events"], "lane_inst_per_event": wi * lanes / ev["events"],
        "lane_efficiency": lanes / 32.0, "issue_slot_utilisation_in_capture": wi / ce,
        "dram_bytes_per_launch_in_capture": dram, "dram_bytes_per_event": dram / ev["events"],
        "registers_per_thread": int(num(m, "launch__registers_per_thread")),
        "gpu_time_ms": num(m, "gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3, "msecond": 1, "usecond": 1e-3, "second": 1e3}.get(m["gpu__time_duration.sum"][1], 1),
        "stall_no_instruction_per_issue": num(m, "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        "stall_long_scoreboard_per_issue": num(m, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "kernel_source_sha256": bench.kernel_source_hash(),
        "kernel_sass_sha256": bench.kernel_sass_hash(m["Kernel Name"][0]),      # of the built library the capture ran on
    }
    if ev.get("records"):
        out["records_in_capture"] = ev["records"]
        out["dram_bytes_per_record_byte"] = dram / (32.0 * ev["records"])
    if a.bench_dram_csv:
        tot, k = 0.0, 0
        for r in csv.DictReader(l for l in open(a.bench_dram_csv) if l.startswith('"')):
            if "mjp_line_kernel" in r.get("Kernel Name", "") and r.get("Metric Name", "").startswith("dram__bytes"):
                tot += float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r["Metric Unit"]]
                k += 1
        if k:
            out["dram_bytes_per_launch_bench_config"] = tot / (k / 2)
    p = os.path.join(ROOT, "profiles", f"issue_model{'_' + a.name if a.name else ''}.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out, indent=1))
    if a.summary:
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), a.rep, str(ev["events"])],
                             stdout=subprocess.PIPE, text=True).stdout
        open(a.summary, "w").write(txt)


if __name__ == "__main__":
    main()
