#!/usr/bin/env python
"""Every kernel family once at 3840x2160 (few frames), for an ncu pass with duration + DRAM-byte metrics:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python tools/all_kernels.py"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200 import video_source_yuv as vy
from fovvideovdp_b200.synthetic import synth_pair_torch, synth_yuv_pair

dev = torch.device("cuda:0")
N, H, W = 16, 2160, 3840
t, r = synth_pair_torch(N, H, W, dev)
gaze = np.stack([np.linspace(0, W - 1, N), np.linspace(0, H - 1, N)], 1).astype(np.float32)
m.fvvdp(display_name="standard_4k", device=dev).predict(t, r, frames_per_second=30)                       # ws level 0 (7-position ring) + fused levels 1..
m.fvvdp(display_name="standard_4k", device=dev).predict(t, r, frames_per_second=60)                       # ws16 on every level (15-position ring)
m.fvvdp(display_name="standard_4k", device=dev).predict(t[:, :, :8], r[:, :, :8], frames_per_second=120)  # 32 taps: front_pairs_kernel + two-plane fused band kernels
m.fvvdp(display_name="standard_hdr_pq", device=dev, foveated=True).predict(t, r, frames_per_second=30, fixation_point=gaze)
m.fvvdp(display_name="standard_4k", device=dev).predict(t[0, 0, 0], r[0, 0, 0], dim_order="HW")           # image
t8 = (t[0, 0, :4, :, :, None] * 255).round().to(torch.uint8).expand(4, H, W, 3).contiguous()
r8 = (r[0, 0, :4, :, :, None] * 255).round().to(torch.uint8).expand(4, H, W, 3).contiguous()
m.fvvdp(display_name="standard_4k", device=dev).predict(t8, r8, dim_order="FHWC", frames_per_second=30)   # luminance front end (u8 RGB)
m.fvvdp(display_name="standard_4k", device=dev, heatmap="threshold").predict(t[:, :, :3], r[:, :, :3], frames_per_second=30)  # heat map + colour map
m.pu_psnr(device=dev).predict(t[:, :, :2], r[:, :, :2], frames_per_second=30)                             # PU21-PSNR, block kernel (EOTF + PU21 fused)
yt, yr = synth_yuv_pair(2, H, W, 10, "420")
with tempfile.TemporaryDirectory() as d:
    props = dict(width=W, height=H, bit_depth=10, color_space="2020", chroma_ss="420", fps=30)
    ft, fr = os.path.join(d, vy.create_yuv_fname("t", props)), os.path.join(d, vy.create_yuv_fname("r", props))
    yt.tofile(ft)
    yr.tofile(fr)
    vs = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry="standard_hdr_pq")
    vs.get_test_frame(0, dev)                                                                              # yuv conversion kernel (per frame)
    m.fvvdp(display_name="standard_hdr_pq", device=dev).predict_video_source(vs)                             # yuv block path (yuv_planes_kernel)
torch.cuda.synchronize()
print("done")
