#!/usr/bin/env python
"""Opcode histogram per kernel of libfvvdp_b200.so (cuobjdump -sass): the SASS evidence behind the Blackwell-native claims --
TMA (UTMALDG), mbarriers (SYNCS), register re-allocation between warp roles (USETMAXREG), packed fp32x2 arithmetic
(FFMA2 / FADD2 / FMUL2), three-input min/max (FMNMX3), MUFU special-function ops, named barriers (BAR), spills (STL/LDL).
usage: sass_summary.py [library] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fovvideovdp_b200", "_lib", "libfvvdp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTMALDG", "UTMASTG", "SYNCS", "USETMAXREG", "BAR", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "FMNMX", "MUFU.LG2", "MUFU.EX2", "MUFU.RCP",
        "MUFU.SQRT", "MUFU.RSQ", "LDS", "STS", "LDG", "STG", "LDGSTS", "SHFL", "LDCU", "STL", "LDL", "BRX", "HMMA", "UTCHMMA"]
kern, counts, arch = None, {}, set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[kern][k] += 1
                break
print(f"# {os.path.relpath(lib, ROOT)}: cubin architectures {sorted(arch)}; static instruction counts per kernel (cuobjdump -sass)")
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
for name, pretty in sorted(zip(counts, demangle), key=lambda kv: -counts[kv[0]]["total"]):
    c = counts[name]
    if c["total"] < 50:
        continue
    short = re.sub(r"\(fvvdp::fused::BandParams\)|fvvdp::|\(anonymous namespace\)::", "", pretty)[:110]
    parts = [f"{k}={c[k]}" for k in KEYS if c[k]]
    print(f"{short}\n    total={c['total']}  " + " ".join(parts))
