#!/usr/bin/env python
"""Table of one ncu pass with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch (csv log):
kernel, grid, block, ms, DRAM GB, GB/s, % of the measured copy peak.   usage: ncu_table.py log.csv [peak GB/s]"""
import csv
import json
import os
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
launches = {}
for r in rows:
    d = launches.setdefault(r[0], {"name": r[4], "block": r[7], "grid": r[8]})
    d[r[12]] = (float(r[14]), r[13])
print(f"{'kernel':84s} {'grid':>16s} {'block':>14s} {'ms':>9s} {'DRAM GB':>9s} {'GB/s':>8s} {'% peak':>7s}")
for d in launches.values():
    name = re.sub(r"\(fvvdp::fused::BandParams\)|fvvdp::|void |\(anonymous namespace\)::", "", d["name"])
    if name.startswith("at::") or not any(k in name for k in ("band", "front", "level", "final", "pool", "recon", "vis_", "yuv", "pu_", "luminance")):
        continue
    t, tu = d.get("gpu__time_duration.sum", (0, "ns"))
    ms = t / 1e6 if tu in ("ns", "nsecond") else (t / 1e3 if tu.startswith("us") else t)
    def gb(k):
        v, u = d.get(k, (0, "byte"))
        return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
    g = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    rate = g / (ms / 1e3) if ms > 0 else 0
    print(f"{name[:84]:84s} {d['grid']:>16s} {d['block']:>14s} {ms:9.4f} {g:9.4f} {rate:8.1f} {100 * rate / peak:6.1f}%")
