#!/usr/bin/env python
"""End-to-end rate of a raw .yuv pair through fvvdp_video_source_yuv_file + predict_video_source(): the files are written once
(synthetic clip, 10-bit 4:2:0 BT.709 by default) and read back from the page cache; host->device copies of the raw frames and
the conversion kernel are inside the timed region.   python tools/time_yuv.py [--size 3840x2160] [--frames 64] [--dir /dev/shm]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200 import video_source_yuv as vy
from fovvideovdp_b200.synthetic import synth_yuv_pair

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--fps", type=int, default=30)
ap.add_argument("--bits", type=int, default=10)
ap.add_argument("--dir", default="/dev/shm")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--resize", default=None, help="full-screen resize mode (nearest, bilinear, bicubic, area) to --resize-to")
ap.add_argument("--resize-to", default="3840x2160")
a = ap.parse_args()
W, H = [int(v) for v in a.size.split("x")]
props = dict(width=W, height=H, bit_depth=a.bits, color_space="709", chroma_ss="420", fps=a.fps)
ft, fr = [os.path.join(a.dir, vy.create_yuv_fname(n, props)) for n in ("fvvdp_b200_test", "fvvdp_b200_ref")]
with open(ft, "wb") as f1, open(fr, "wb") as f2:
    for f0 in range(0, a.frames, 8):  # written in chunks to bound memory
        t, r = synth_yuv_pair(min(8, a.frames - f0), H, W, a.bits, "420")
        t.tofile(f1)
        r.tofile(f2)
try:
    dev = torch.device("cuda:0")
    fv = m.fvvdp(display_name="standard_4k", device=dev)
    rw, rh = [int(v) for v in a.resize_to.split("x")]
    vs = m.fvvdp_video_source_yuv_file(ft, fr, display_photometry="standard_4k", full_screen_resize=a.resize, resize_resolution=(rw, rh) if a.resize else None)
    for _ in range(2):
        jod, _ = fv.predict_video_source(vs)
        float(jod)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        jod, _ = fv.predict_video_source(vs)
        jod = float(jod)
    dt = (time.perf_counter() - t0) / a.steps
    nbytes = os.path.getsize(ft) + os.path.getsize(fr)
    print(f".yuv pair {W}x{H}x{a.frames} {a.bits}-bit 4:2:0{(' ' + a.resize + ' -> ' + a.resize_to) if a.resize else ''} from {a.dir}: {a.frames / dt:.1f} frames/s ({dt * 1e3:.1f} ms/clip, "
          f"{nbytes / dt / 1e9:.1f} GB/s of file data), JOD={jod:.4f}, block={fv.last_run['block_frames']}, launches={fv.last_run['gpu_launches']}")
finally:
    os.remove(ft)
    os.remove(fr)
