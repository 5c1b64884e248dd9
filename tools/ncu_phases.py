#!/usr/bin/env python
"""Instruction and stall-sample shares per phase of the fused band kernel from an .ncu-rep source page.
usage: ncu_phases.py report.ncu-rep 'name:lo-hi,lo-hi;name2:lo-hi' [kernel-index]   (line ranges of fvvdp_fused.cuh)"""
import csv
import subprocess
import sys


def main():
    rep, spec = sys.argv[1], sys.argv[2]
    kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    phases = []
    for part in spec.split(";"):
        name, rng = part.split(":")
        phases.append((name, [tuple(int(v) for v in r.split("-")) for r in rng.split(",")]))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, kern, fname, hdr = [], None, None, None
    acc = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1]
        elif r[0] == "Function Name":
            kern = r[1]
            if kern not in kernels:
                kernels.append(kern)
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit() and kernels.index(kern) == kidx:
            d = dict(zip(hdr, r))
            inst = int(d["Instructions Executed"]) if d["Instructions Executed"].isdigit() else 0
            samp = int(d["# Samples"]) if d["# Samples"].isdigit() else 0
            line = int(r[0])
            name = "other(" + fname.split("/")[-1] + ")"
            if fname.endswith("fvvdp_fused.cuh"):
                name = "unassigned"
                for pn, rngs in phases:
                    if any(lo <= line <= hi for lo, hi in rngs):
                        name = pn
                        break
            a = acc.setdefault(name, [0, 0])
            a[0] += inst
            a[1] += samp
    ti = sum(a[0] for a in acc.values())
    ts = sum(a[1] for a in acc.values())
    print("kernel:", kernels[kidx])
    for name, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:28s} inst {100.0 * a[0] / ti:5.1f}%   stall samples {100.0 * a[1] / ts:5.1f}%")


if __name__ == "__main__":
    main()
