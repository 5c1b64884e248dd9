#!/usr/bin/env python
"""A/B of the warp-specialised band kernel against the fused band kernel of the same library (FVVDP_B200_PATH=fused):
per-band relative difference of Q_per_ch on a few clip shapes / paddings / block cuts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

dev = torch.device("cuda:0")
cases = [(12, 270, 480, "replicate", 30, None), (12, 270, 480, "pingpong", 30, None), (12, 270, 480, "circular", 30, None),
         (20, 540, 960, "replicate", 30, None), (20, 540, 960, "pingpong", 24, 7), (16, 1080, 1920, "replicate", 30, None),
         (9, 135, 240, "circular", 25, 4), (10, 2160, 3840, "replicate", 30, None), (10, 1081, 1922, "pingpong", 30, None),
         (20, 270, 480, "replicate", 60, None), (20, 540, 960, "pingpong", 50, 6), (24, 1080, 1920, "circular", 60, None),
         (10, 2160, 3840, "replicate", 60, None), (12, 1081, 1922, "pingpong", 60, 5)]
worst = 0.0
for (N, H, W, pad, fps, T) in cases:
    t, r = synth_pair_torch(N, H, W, dev)
    out = []
    for path in ("ws", "fused"):
        if path == "fused":
            os.environ["FVVDP_B200_PATH"] = "fused"
        else:
            os.environ.pop("FVVDP_B200_PATH", None)
        fv = m.fvvdp(display_name="standard_4k", device=dev, temp_padding=pad, block_frames=T)
        jod, st = fv.predict(t, r, frames_per_second=fps)
        out.append((float(jod), st["Q_per_ch"]))
    q0, q1 = out[0][1], out[1][1]
    rel = np.abs(q0 - q1).max(axis=2) / np.maximum(np.abs(q1).max(axis=2), 1e-12)   # per band and channel
    worst = max(worst, float(rel.max()))
    print(f"{N}x{H}x{W} {pad} {fps}fps T={T}: JOD ws={out[0][0]:.6f} fused={out[1][0]:.6f}  per-band rel diff: " +
          " ".join(f"{v:.1e}" for v in rel.max(axis=1)))
print("worst", worst)
