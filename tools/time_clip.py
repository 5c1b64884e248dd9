#!/usr/bin/env python
"""Time predict() on a resident synthetic clip: python tools/time_clip.py [--fps 60] [--size 3840x2160] [--frames 64] [--foveated] [--display NAME]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

ap = argparse.ArgumentParser()
ap.add_argument("--fps", type=float, default=60)
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--display", default="standard_4k")
ap.add_argument("--foveated", action="store_true")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
W, H = [int(v) for v in a.size.split("x")]
dev = torch.device("cuda:0")
t, r = synth_pair_torch(a.frames, H, W, dev)
fv = m.fvvdp(display_name=a.display, device=dev, foveated=a.foveated)
fix = None
if a.foveated:
    import numpy as np
    fix = np.stack([np.linspace(0, W - 1, a.frames), np.linspace(0, H - 1, a.frames)], 1).astype(np.float32)
for _ in range(2):
    jod, _ = fv.predict(t, r, frames_per_second=a.fps, fixation_point=fix)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    jod, _ = fv.predict(t, r, frames_per_second=a.fps, fixation_point=fix)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.steps
print(f"fps={a.fps} {W}x{H}x{a.frames} display={a.display} foveated={a.foveated}: {a.frames / dt:.1f} frames/s ({dt * 1e3:.2f} ms/clip) JOD={float(jod):.4f} launches={fv.last_run['gpu_launches']}")
