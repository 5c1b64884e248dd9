#!/usr/bin/env python
"""Build experiment variants of libfvvdp_b200.so: tools/build_variants.py name=-DFLAG,-DFLAG2 ...
Each variant lands in fovvideovdp_b200/_lib/variants/<name>/libfvvdp_b200.so (select with FVVDP_B200_LIB).  Only the
warp-specialised units (ws_*) are recompiled with the variant's flags; the other objects come from the main build."""
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fovvideovdp_b200 import build as b

main_obj = b.OBJ_DIR
b.build_native()
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    d = os.path.join(b.PKG, "_lib", "variants", name)
    obj = os.path.join(d, "obj")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(obj)
    rebuild_all = "FUSED" in flags or "WS_TH16" in flags  # flags that reach beyond the warp-specialised units: every unit is recompiled
    for f in os.listdir(main_obj):
        if not f.startswith("ws_") and not rebuild_all:
            shutil.copy2(os.path.join(main_obj, f), os.path.join(obj, f))
    b.OBJ_DIR = obj
    b.LIB = os.path.join(d, "libfvvdp_b200.so")
    saved = list(b.FLAGS)
    b.FLAGS = saved + [f for f in flags.split(",") if f]
    os.environ.pop("FVVDP_B200_LIB", None)
    orig = b.is_stale
    b.is_stale = lambda: True
    print(name, b.build_native())
    b.is_stale = orig
    b.FLAGS = saved
