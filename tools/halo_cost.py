#!/usr/bin/env python
"""Calibration of HALO_SLOT_COST (fvvdp.frame_block): device time of the block each rank of a sharded 4K clip scores, one
emulated rank at a time on ONE GPU (torch.distributed's rank / world size are stubbed; the all-reduce is a no-op).
usage: halo_cost.py [--world 2] [--frames-per-rank 64] [--cost 0.6 0.8 1.0]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import fovvideovdp_b200 as m
import importlib
F = importlib.import_module("fovvideovdp_b200.fvvdp")
from fovvideovdp_b200.synthetic import synth_pair_torch

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--frames-per-rank", type=int, default=64)
ap.add_argument("--cost", type=float, nargs="+", default=[0.6, 0.8, 1.0])
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
H, W, FPS, fl = 2160, 3840, 30, 8
world, n_total = a.world, a.world * a.frames_per_rank
state = dict(rank=0)
dist.is_initialized = lambda: True
dist.get_world_size = lambda *x, **k: world
dist.get_rank = lambda *x, **k: state["rank"]
dist.all_reduce = lambda t, *x, **k: None
fv = m.fvvdp(display_name="standard_4k", device=dev, shard_frames=True)
for cost in a.cost:
    F.frame_block.__defaults__ = F.frame_block.__defaults__[:-1] + (cost,)
    line = []
    for rank in sorted({0, 1, world - 1}):
        state["rank"] = rank
        first, last = F.frame_block(n_total, rank, world, halo=fl - 1, first_halo=1)
        halo = min(first, fl - 1)
        t, r = synth_pair_torch(last - first + halo, H, W, dev, first_frame=first - halo)
        vs = m.fvvdp_video_source_array(t, r, FPS, display_photometry=fv.display_photometry, first_frame=first - halo, total_frames=n_total)
        for _ in range(3):
            fv.predict_video_source(vs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fv.predict_video_source(vs)
        e1.record()
        torch.cuda.synchronize()
        line.append(f"rank {rank}: frames [{first},{last}) + {halo} halo = {e0.elapsed_time(e1) / a.steps:.3f} ms")
        del t, r, vs
    print(f"world {world} cost {cost}: " + "   ".join(line))
