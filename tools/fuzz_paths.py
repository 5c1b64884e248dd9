#!/usr/bin/env python
"""Randomised A/B of the fused band kernels against the general kernels (FVVDP_B200_PATH=v1) of the same library:
random frame sizes (edge tiles, odd sizes, the row-parity quirk), frame rates, paddings, dtypes, channel counts, foveation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fovvideovdp_b200 as m

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda:0")
worst_j, worst_q, bad = 0.0, 0.0, 0
for case in range(n_cases):
    H, W = int(rng.integers(17, 320)), int(rng.integers(17, 420))
    N = int(rng.integers(1, 24))
    fps = [24, 25, 30, 50, 60, 90, 120][int(rng.integers(0, 7))]
    pad = ["replicate", "circular", "pingpong"][int(rng.integers(0, 3))]
    C = [1, 3][int(rng.integers(0, 2))]
    kind = ["f32", "u8", "u16"][int(rng.integers(0, 3))]
    fov = bool(rng.integers(0, 3) == 0)
    disp = ["standard_4k", "standard_fhd", "standard_hdr_pq", "standard_hmd"][int(rng.integers(0, 4))]
    resident = bool(rng.integers(0, 2))
    base = rng.random((1, C, N, H, W), dtype=np.float32)
    base = 0.5 * base + 0.5 * np.roll(base, 1, axis=4)  # some spatial correlation
    test = np.clip(base + 0.05 * rng.standard_normal(base.shape).astype(np.float32), 0, 1)
    if kind == "u8":
        base, test = (base * 255).astype(np.uint8), (test * 255).astype(np.uint8)
    elif kind == "u16":
        base, test = (base * 65535).astype(np.uint16), (test * 65535).astype(np.uint16)
    a, b = test, base
    if resident and kind != "u16":
        a, b = torch.from_numpy(test).to(dev), torch.from_numpy(base).to(dev)
    kw = dict(frames_per_second=fps) if N > 1 else {}
    if fov:
        kw["fixation_point"] = np.stack([rng.uniform(0, W - 1, N), rng.uniform(0, H - 1, N)], 1).astype(np.float32) if N > 1 else \
            np.array([rng.uniform(0, W - 1), rng.uniform(0, H - 1)], np.float32)
    out = []
    for path in ("fused", "v1"):
        if path == "v1":
            os.environ["FVVDP_B200_PATH"] = "v1"
        else:
            os.environ.pop("FVVDP_B200_PATH", None)
        try:
            fv = m.fvvdp(display_name=disp, device=dev, foveated=fov, temp_padding=pad)
            jod, st = fv.predict(a, b, **kw)
            out.append((float(jod), st["Q_per_ch"]))
        except RuntimeError as e:
            out.append(("error", str(e)))
    if out[0][0] == "error" or out[1][0] == "error":
        same_err = out[0][0] == out[1][0] == "error"
        print(f"case {case}: {H}x{W}x{N} C={C} {kind} fps={fps} -> {out[0][1] if out[0][0]=='error' else 'ok'} | {out[1][1] if out[1][0]=='error' else 'ok'}")
        bad += 0 if same_err else 1
        continue
    dj = abs(out[0][0] - out[1][0]) / out[1][0]
    scale = np.maximum(np.abs(out[1][1]).max(axis=(0, 2), keepdims=True), 1e-6)
    dq = float((np.abs(out[0][1] - out[1][1]) / scale).max())
    worst_j, worst_q = max(worst_j, dj), max(worst_q, dq)
    if dj > 1e-5 or dq > 5e-4:
        bad += 1
        print(f"case {case}: MISMATCH {H}x{W}x{N} C={C} {kind} fps={fps} pad={pad} fov={fov} {disp} resident={resident}: dJOD={dj:.2e} dQ={dq:.2e}")
print(f"{n_cases} cases, {bad} mismatches, worst relative JOD difference {worst_j:.2e}, worst Q difference / channel max {worst_q:.2e}")
sys.exit(1 if bad else 0)
