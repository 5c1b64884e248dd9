#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (source page, needs -lineinfo).
usage: ncu_lines.py report.ncu-rep [kernel-index] [top-N]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # split into kernels: a kernel section starts with a "Function Name" row sequence; files repeat per kernel
    sections, cur, fname, kern = [], None, None, None
    kernels = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1]
        elif r[0] == "Function Name":
            kern = r[1]
            if not kernels or kernels[-1] != kern:
                kernels.append(kern)
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit():
            sections.append((kern, fname, int(r[0]), r[1], hdr, r))
    want = kernels[kidx]
    print("kernel:", want)
    tot_i = tot_s = 0
    items = []
    for kern, fname, line, src, hdr, r in sections:
        if kern != want:
            continue
        d = dict(zip(hdr, r))
        def num(k):
            v = d.get(k, "0")
            return int(v) if v.isdigit() else 0
        inst, samp = num("Instructions Executed"), num("# Samples")
        tot_i += inst
        tot_s += samp
        items.append((inst, samp, fname.split("/")[-1], line, src.strip()))
    items.sort(reverse=True)
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for inst, samp, f, line, src in items[:top]:
        print(f"{100.0 * inst / tot_i:5.1f}% inst {100.0 * samp / max(tot_s, 1):5.1f}% stall  {f}:{line}  {src[:110]}")


if __name__ == "__main__":
    main()
