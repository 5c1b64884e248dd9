#!/usr/bin/env python
"""Heat-map throughput at 3840x2160: predict(heatmap=...) including the per-frame device->host copy of the map into the result tensor.
python tools/time_heatmap.py [--frames 32]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=32)
a = ap.parse_args()
dev = torch.device("cuda:0")
t, r = synth_pair_torch(a.frames, 2160, 3840, dev)
for kind in ("raw", "threshold"):
    fv = m.fvvdp(display_name="standard_4k", device=dev, heatmap=kind)
    for _ in range(2):
        jod, st = fv.predict(t, r, frames_per_second=30)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        jod, st = fv.predict(t, r, frames_per_second=30)
        hm = st["heatmap"]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"heatmap={kind}: {a.frames / dt:.1f} frames/s ({dt * 1e3 / a.frames:.2f} ms per frame), result {tuple(hm.shape)} {hm.dtype}, "
          f"{hm.numel() * 2 / dt / 1e9:.1f} GB/s into the host tensor, JOD={float(jod):.4f}")
