#!/usr/bin/env python
"""Device time of the .yuv conversion kernels with and without full-screen resize (one frame, one stream per launch):
python tools/time_resize.py [--size 1920x1080] [--to 3840x2160]"""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fovvideovdp_b200 import _native
from fovvideovdp_b200 import video_source_yuv as vy
from fovvideovdp_b200.display_model import fvvdp_display_photometry, photometry_kernel_spec
from fovvideovdp_b200.synthetic import synth_yuv_pair

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--to", default="3840x2160")
a = ap.parse_args()
W, H = [int(v) for v in a.size.split("x")]
ow, oh = [int(v) for v in a.to.split("x")]
dev = torch.device("cuda:0")
t, _ = synth_yuv_pair(1, H, W, 10, "420")
spec = photometry_kernel_spec(fvvdp_display_photometry.load("standard_4k"))
with tempfile.TemporaryDirectory() as d:
    props = dict(width=W, height=H, bit_depth=10, color_space="709", chroma_ss="420", fps=30)
    f = os.path.join(d, vy.create_yuv_fname("t", props))
    t.tofile(f)
    rd = vy.YUVReader(f)
    raw = torch.from_numpy(t[0].view("int16")).to(dev)
    esz = 2
    for mode in (None, "nearest", "bilinear", "bicubic", "area"):
        resize = None if mode is None else (mode, (ow, oh))
        w, h = (W, H) if mode is None else (ow, oh)
        lum = torch.empty((h, w), dtype=torch.float32, device=dev)
        desc = rd._desc(spec, [0.2126, 0.7152, 0.0722], resize)
        st = torch.cuda.current_stream(dev).cuda_stream

        def run():
            _native.yuv_to_luminance(desc, raw.data_ptr(), raw.data_ptr() + rd.y_pixels * esz, raw.data_ptr() + (rd.y_pixels + rd.uv_pixels) * esz,
                                     lum.data_ptr(), 0, 0, st)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{W}x{H} -> {w}x{h} {mode or 'no resize'}: {ms:.4f} ms per frame and stream ({w * h / ms / 1e6:.1f} Gpx/s out)")
