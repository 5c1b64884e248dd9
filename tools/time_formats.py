#!/usr/bin/env python
"""Resident-clip throughput for other input formats (uint8 / uint16, RGB, channel-last): the generic staging path."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

dev = torch.device("cuda:0")
N, H, W = 32, 2160, 3840
t, r = synth_pair_torch(N, H, W, dev)
cases = {
    "f32 1ch BCFHW": (t, r, "BCFHW"),
    "u8 1ch BCFHW": ((t * 255).round().to(torch.uint8), (r * 255).round().to(torch.uint8), "BCFHW"),
    "u8 RGB BCFHW (planar)": ((t * 255).round().to(torch.uint8).expand(1, 3, N, H, W).contiguous(), (r * 255).round().to(torch.uint8).expand(1, 3, N, H, W).contiguous(), "BCFHW"),
    "u8 RGB FHWC (interleaved)": ((t[0, 0, :, :, :, None] * 255).round().to(torch.uint8).expand(N, H, W, 3).contiguous(),
                                  (r[0, 0, :, :, :, None] * 255).round().to(torch.uint8).expand(N, H, W, 3).contiguous(), "FHWC"),
    "f32 RGB BCFHW": (t.expand(1, 3, N, H, W).contiguous(), r.expand(1, 3, N, H, W).contiguous(), "BCFHW"),
}
fv = m.fvvdp(display_name="standard_4k", device=dev)
for name, (a, b, order) in cases.items():
    for _ in range(2):
        jod, _ = fv.predict(a, b, dim_order=order, frames_per_second=30)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        jod, _ = fv.predict(a, b, dim_order=order, frames_per_second=30)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {N / dt:.1f} frames/s ({dt * 1e3:.2f} ms/clip) JOD={float(jod):.4f}")
