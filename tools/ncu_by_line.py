#!/usr/bin/env python
"""Attribute an ncu SASS source page to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_by_line.py sass.csv mjpdes_b200/csrc/libmjpdes_b200.so '<mangled kernel name>' [top]

The per-instruction rows of the ncu csv are in address order; nvdisasm -g prints the same
instructions with `//## File "...", line N` markers (needs -lineinfo at compile time)."""
import csv, os, re, subprocess, sys, tempfile

def line_map(so, mangled):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
    # one cubin per translation unit (mjp_variant.cu is compiled once per kernel family): find the kernel's
    txt = ""
    for cub in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):
        syms = subprocess.run(["cuobjdump", "-elf", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
        if ".text." + mangled in syms:
            txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
            break
    out, cur, on = [], (None, 0), False
    for ln in txt.splitlines():
        if ln.startswith("//--------------------- .text."):
            on = ln.split(".text.")[1].split()[0] == mangled
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (os.path.basename(m.group(1)), int(m.group(2)), (os.path.basename(inl.group(1)), int(inl.group(2))) if inl else None)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    return out

def main():
    sass, so, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    rows = list(csv.reader(open(sass)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ie, te, sm, sc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    inst = [r for r in rows[hi + 1:] if len(r) > te]
    lm = line_map(so, mangled)
    if len(lm) != len(inst):
        print(f"warning: {len(inst)} profiled instructions vs {len(lm)} disassembled", file=sys.stderr)
    agg = {}
    for r, loc in zip(inst, lm):
        key = loc[2] if (loc[2] and loc[0] != "mjp_line_kernel.cuh") else loc[:2]   # charge inlined helpers to the call site
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += int(r[ie] or 0); a[1] += int(r[te] or 0); a[2] += int(r[sm] or 0)
    ti, tt, ts = (sum(a[k] for a in agg.values()) for k in range(3))
    print(f"total warp-inst {ti}  thread-inst {tt}  lanes/inst {tt / ti:.2f}  samples {ts}")
    src = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(so)), f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
        print(f"{100 * a[0] / ti:5.1f}% inst {a[1] / max(a[0], 1):5.1f} lanes {100 * a[2] / max(ts, 1):5.1f}% smp | {f}:{l:<4} {text[:100]}")

if __name__ == "__main__":
    main()
