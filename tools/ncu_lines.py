#!/usr/bin/env python
"""Per-source-line summary of an ncu report's source page (needs -lineinfo and --import-source on):

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [--top 40] [--ranges file:lo-hi:name,...]

Prints instructions executed, stall samples and the dominant stall reasons per source line, sorted by
samples, and optionally per named line range."""
import csv
import subprocess
import sys
from collections import defaultdict


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines = {}
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("", "Function Name", "Kernel Name"):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        lines[(fname, ln)] = (r[1], d)
    return lines


def num(s):
    try:
        return float(s)
    except (ValueError, TypeError):
        return 0.0


def main():
    rep = sys.argv[1]
    top = 40
    ranges = []
    a = sys.argv[2:]
    while a:
        if a[0] == "--top":
            top = int(a[1]); a = a[2:]
        elif a[0] == "--ranges":
            for spec in a[1].split(","):
                f, lr, name = spec.split(":")
                lo, hi = lr.split("-")
                ranges.append((f, int(lo), int(hi), name))
            a = a[2:]
        else:
            a = a[1:]
    lines = load(rep)
    tot_i = sum(num(d["Instructions Executed"]) for _, d in lines.values())
    tot_s = sum(num(d["# Samples"]) for _, d in lines.values())
    print(f"total instructions {tot_i:.4g}  samples {tot_s:.0f}")
    stall_keys = [k for k in next(iter(lines.values()))[1] if k.startswith("stall_") and "Not Issued" not in k]
    if ranges:
        print("\nper range:  name  inst%  samples%  top stalls")
        for f, lo, hi, name in ranges:
            sel = [(k, v) for k, v in lines.items() if k[0] == f and lo <= k[1] <= hi]
            i = sum(num(v[1]["Instructions Executed"]) for _, v in sel)
            s = sum(num(v[1]["# Samples"]) for _, v in sel)
            st = defaultdict(float)
            for _, v in sel:
                for sk in stall_keys:
                    st[sk] += num(v[1][sk])
            tops = sorted(st.items(), key=lambda kv: -kv[1])[:5]
            print(f"{name:28s} {100 * i / tot_i:6.2f} {100 * s / tot_s:6.2f}  " +
                  " ".join(f"{k[6:]}={100 * v / max(s, 1):.0f}%" for k, v in tops))
    print("\nper line (by samples):")
    for (f, ln), (src, d) in sorted(lines.items(), key=lambda kv: -num(kv[1][1]["# Samples"]))[:top]:
        s = num(d["# Samples"]); i = num(d["Instructions Executed"])
        st = sorted(((sk, num(d[sk])) for sk in stall_keys), key=lambda kv: -kv[1])[:3]
        print(f"{f}:{ln:<5d} inst {100 * i / tot_i:5.2f}% smp {100 * s / tot_s:5.2f}%  "
              + " ".join(f"{k[6:]}={v:.0f}" for k, v in st) + "  | " + src.strip()[:90])


if __name__ == "__main__":
    main()
