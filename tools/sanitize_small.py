#!/usr/bin/env python
"""Small clips through every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_numpy, synth_pair_torch

dev = torch.device("cuda:0")
t, r = synth_pair_numpy(10, 70, 150)
td, rd = synth_pair_torch(10, 72, 160, dev)
gaze = np.stack([np.linspace(0, 149, 10), np.linspace(0, 69, 10)], 1).astype(np.float32)
runs = [
    ("30 fps, host clip (cp.async staging)", dict(display_name="standard_fhd"), (t, r), dict(frames_per_second=30)),
    ("30 fps, resident clip (TMA staging)", dict(display_name="standard_fhd"), (td, rd), dict(frames_per_second=30)),
    ("60 fps (16-frame ring)", dict(display_name="standard_fhd"), (td, rd), dict(frames_per_second=60)),
    ("120 fps (general kernels)", dict(display_name="standard_fhd"), (td, rd), dict(frames_per_second=120)),
    ("image", dict(display_name="standard_fhd"), (td[0, 0, 0], rd[0, 0, 0]), dict(dim_order="HW")),
    ("foveated PQ", dict(display_name="standard_hdr_pq", foveated=True), (0.1 + 0.65 * t, 0.1 + 0.65 * r), dict(frames_per_second=30, fixation_point=gaze)),
    ("heat map, colour map", dict(display_name="standard_fhd", heatmap="threshold"), (t, r), dict(frames_per_second=30)),
    ("uint8 RGB (luminance front end)", dict(display_name="standard_fhd"),
     ((t[0, 0, :, :, :, None] * 255).astype(np.uint8).repeat(3, -1), (r[0, 0, :, :, :, None] * 255).astype(np.uint8).repeat(3, -1)),
     dict(dim_order="FHWC", frames_per_second=30)),
]
for name, ctor, (a, b), kw in runs:
    jod, _ = m.fvvdp(device=dev, **ctor).predict(a, b, **kw)
    torch.cuda.synchronize()
    print(f"{name}: JOD {float(jod):.4f}")
q, _ = m.pu_psnr(device=dev).predict(t, r, frames_per_second=30)
print(f"PU21-PSNR {float(q):.3f} dB")
# raw .yuv clips: per-frame conversion, block path, full-screen resize in every mode (up and down)
import tempfile

from fovvideovdp_b200 import video_source_yuv as vy
from fovvideovdp_b200.synthetic import synth_yuv_pair

for bits, css, cs, disp in ((10, "420", "2020", "standard_hdr_pq"), (8, "444", "709", "standard_fhd")):
    yt, yr = synth_yuv_pair(5, 70, 150 if css == "444" else 152, bits, css)
    with tempfile.TemporaryDirectory() as d:
        props = dict(width=150 if css == "444" else 152, height=70, bit_depth=bits, color_space=cs, chroma_ss=css, fps=30)
        ft, fr = os.path.join(d, vy.create_yuv_fname("t", props)), os.path.join(d, vy.create_yuv_fname("r", props))
        yt.tofile(ft)
        yr.tofile(fr)
        fv = m.fvvdp(device=dev, display_name=disp)
        jod, _ = fv.predict_video_source(vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp))
        print(f"yuv {bits}b {css}: JOD {float(jod):.4f}")
        for mode in ("nearest", "bilinear", "bicubic", "area"):
            for res in ((211, 97), (90, 41)):
                vs = vy.fvvdp_video_source_yuv_file(ft, fr, display_photometry=disp, full_screen_resize=mode, resize_resolution=res)
                vs.get_test_frame(0, dev)
                jod, _ = fv.predict_video_source(vs)
                torch.cuda.synchronize()
                print(f"yuv {bits}b {css} resize {mode} -> {res[0]}x{res[1]}: JOD {float(jod):.4f}")
