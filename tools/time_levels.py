#!/usr/bin/env python
"""Per-kernel-class device time of one predict() (CUDA events recorded inside the library): python tools/time_levels.py [--fps 120] ..."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

ap = argparse.ArgumentParser()
ap.add_argument("--fps", type=float, default=120)
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--display", default="standard_4k")
ap.add_argument("--foveated", action="store_true")
a = ap.parse_args()
W, H = [int(v) for v in a.size.split("x")]
dev = torch.device("cuda:0")
t, r = synth_pair_torch(a.frames, H, W, dev)
fv = m.fvvdp(display_name=a.display, device=dev, foveated=a.foveated)
import numpy as np
gaze = np.stack([np.linspace(0, W - 1, a.frames), np.linspace(0, H - 1, a.frames)], 1).astype(np.float32) if a.foveated else None
for _ in range(2):
    fv.predict(t, r, frames_per_second=a.fps, fixation_point=gaze)
fv._ctx.profile(True)
jod, _ = fv.predict(t, r, frames_per_second=a.fps, fixation_point=gaze)
prof = fv._ctx.profile_read()
print(f"fps={a.fps} {W}x{H}x{a.frames} {a.display}{' foveated' if a.foveated else ''}: " + "  ".join(f"{k}={v[0]:.3f}ms" for k, v in prof.items()) + f"  total={sum(v[0] for v in prof.values()):.3f}ms")
