#!/usr/bin/env python
"""Producer / consumer split of the warp-specialised band kernel from an .ncu-rep (SASS source page): instructions, stall
samples and stall reasons per code segment.  Segments are cut at the marker instructions of the kernel: USETMAXREG (role
entry), BAR.SYNC 0x1 (the producers' two named barriers), SYNCS.ARRIVE (full / empty hand-off), EXIT.
usage: ncu_roles.py report.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

REASONS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_math", "stall_mio", "stall_wait", "stall_not_selected", "stall_selected",
           "stall_branch_resolving", "stall_no_inst", "stall_dispatch", "stall_lg"]


def main():
    rep = sys.argv[1]
    kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            kernels.append((r[1], []))
        elif r[0] == "Address":
            hdr = r
        elif r[0].startswith("0x") and kernels:
            kernels[-1][1].append(dict(zip(hdr, r)))
    name, ins = kernels[kidx]
    print("kernel:", name)
    segs, seg = [], dict(name="setup", rows=[])
    n_bar = 0
    for d in ins:
        src = d["Source"].strip()
        seg["rows"].append(d)
        cut = None
        if "USETMAXREG.DEALLOC" in src:
            cut = "producer: prologue"
        elif "USETMAXREG.TRY_ALLOC" in src:
            cut = "consumer: prologue + wait full + ring/filters"
        elif src.startswith("BAR.SYNC") and "0x1" in src:
            n_bar += 1
            cut = "producer: rows" if n_bar == 1 else "producer: columns + coarse filters"
        elif src.startswith("BAR.SYNC") and "0x2" in src:
            cut = "consumer: final sums"
        if cut:
            segs.append(seg)
            seg = dict(name=cut, rows=[])
    segs.append(seg)

    def num(d, k):
        v = d.get(k, "0")
        try:
            return int(v)
        except ValueError:
            return 0
    ti = sum(num(d, "Instructions Executed") for d in ins)
    ts = sum(num(d, "# Samples") for d in ins)
    print(f"total warp-instructions {ti}, samples {ts}")
    for s in segs:
        i = sum(num(d, "Instructions Executed") for d in s["rows"])
        sm = sum(num(d, "# Samples") for d in s["rows"])
        rs = {k: sum(num(d, k) for d in s["rows"]) for k in REASONS}
        top = sorted(rs.items(), key=lambda kv: -kv[1])[:5]
        print(f"{s['name']:48s} inst {100.0 * i / ti:5.1f}%  samples {100.0 * sm / ts:5.1f}%   " +
              " ".join(f"{k[6:]}={100.0 * v / max(sm, 1):.0f}%" for k, v in top))
    # hottest instructions by samples
    hot = sorted(ins, key=lambda d: -num(d, "# Samples"))[:25]
    for d in hot:
        rs = sorted(((k, num(d, k)) for k in REASONS), key=lambda kv: -kv[1])[:2]
        print(f"  {100.0 * num(d, '# Samples') / ts:5.2f}%  {d['Address'][-5:]}  {d['Source'].strip()[:70]:70s} " + " ".join(f"{k[6:]}={v}" for k, v in rs))


if __name__ == "__main__":
    main()
