#!/usr/bin/env python
"""Host-side cost of predict_video_source(): cProfile over repeated calls on a small resident clip (GPU time negligible)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

dev = torch.device("cuda:0")
t, r = synth_pair_torch(64, 270, 480, dev)
fv = m.fvvdp(display_name="standard_4k", device=dev)
vs = m.fvvdp_video_source_array(t, r, 30, display_photometry=fv.display_photometry)
for _ in range(5):
    fv.predict_video_source(vs)
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    fv.predict_video_source(vs)
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / N * 1e3:.3f} ms per predict_video_source (64 frames 480x270)")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    fv.predict_video_source(vs)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
