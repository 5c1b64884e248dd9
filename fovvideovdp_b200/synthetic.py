"""Deterministic analytic test/reference clips (no RNG => identical wherever they are generated).

ref  = clip(0.5 + 0.25 sin(2pi(x/37 + f/11)) cos(2pi y/53) + 0.15 sin(2pi(x+2y)/7 + 0.3 f), 0, 1)
test = clip(ref_unclipped + 0.04 sin(2pi(x/3 + y/5) + 1.7 f) + 0.02 cos(2pi f/4), 0, 1)

Evaluated in float64, stored as float32 display-encoded values in [0,1], layout (1,1,N,H,W)
(SURVEY.md section 8d).  `first_frame` lets a rank generate only its own frame block of a longer clip.
"""
import math

import numpy as np


def synth_pair_numpy(n_frames, height, width, first_frame=0):
    f = (np.arange(n_frames, dtype=np.float64) + first_frame)[:, None, None]
    y = np.arange(height, dtype=np.float64)[None, :, None]
    x = np.arange(width, dtype=np.float64)[None, None, :]
    tp = 2.0 * math.pi
    r = 0.5 + 0.25 * np.sin(tp * (x / 37.0 + f / 11.0)) * np.cos(tp * y / 53.0) + 0.15 * np.sin(tp * (x + 2.0 * y) / 7.0 + 0.3 * f)
    t = r + 0.04 * np.sin(tp * (x / 3.0 + y / 5.0) + 1.7 * f) + 0.02 * np.cos(tp * f / 4.0)
    ref = np.clip(r, 0.0, 1.0).astype(np.float32)[None, None]
    test = np.clip(t, 0.0, 1.0).astype(np.float32)[None, None]
    return test, ref


def synth_pair_torch(n_frames, height, width, device, first_frame=0, chunk=8):
    """Same pattern generated on `device` (float64 maths, float32 result), frame-chunked to bound memory."""
    import torch

    test = torch.empty((1, 1, n_frames, height, width), dtype=torch.float32, device=device)
    ref = torch.empty_like(test)
    y = torch.arange(height, dtype=torch.float64, device=device)[None, :, None]
    x = torch.arange(width, dtype=torch.float64, device=device)[None, None, :]
    tp = 2.0 * math.pi
    for f0 in range(0, n_frames, chunk):
        f1 = min(n_frames, f0 + chunk)
        f = (torch.arange(f0, f1, dtype=torch.float64, device=device) + first_frame)[:, None, None]
        r = 0.5 + 0.25 * torch.sin(tp * (x / 37.0 + f / 11.0)) * torch.cos(tp * y / 53.0) + 0.15 * torch.sin(tp * (x + 2.0 * y) / 7.0 + 0.3 * f)
        t = r + 0.04 * torch.sin(tp * (x / 3.0 + y / 5.0) + 1.7 * f) + 0.02 * torch.cos(tp * f / 4.0)
        ref[0, 0, f0:f1] = r.clamp(0.0, 1.0).float()
        test[0, 0, f0:f1] = t.clamp(0.0, 1.0).float()
    return test, ref


def synth_yuv_pair(n_frames, height, width, bit_depth=10, chroma_ss="420"):
    """The same clips as limited-range planar Y'CbCr code values (test, ref), each (n_frames, frame_pixels) uint8 / uint16
    in file order (Y plane, Cb plane, Cr plane per frame): luma from the pattern above, smooth analytic chroma."""
    test, ref = synth_pair_numpy(n_frames, height, width)
    sc = float(2 ** (bit_depth - 8))
    ch, cw = (height // 2, width // 2) if chroma_ss == "420" else (height, width)
    f = np.arange(n_frames, dtype=np.float64)[:, None, None]
    y = np.arange(ch, dtype=np.float64)[None, :, None]
    x = np.arange(cw, dtype=np.float64)[None, None, :]
    tp = 2.0 * math.pi
    cb = 0.30 * np.sin(tp * (x / 23.0 + y / 31.0) + 0.2 * f)
    cr = 0.25 * np.cos(tp * (x / 17.0 - y / 41.0) + 0.1 * f)
    dtype = np.uint16 if bit_depth > 8 else np.uint8
    out = []
    for luma, dc in ((test[0, 0], 0.02), (ref[0, 0], 0.0)):
        Y = np.round((16.0 + 219.0 * luma.astype(np.float64)) * sc)
        U = np.round((128.0 + 224.0 * (cb + dc * np.sin(tp * x / 5.0))) * sc)
        V = np.round((128.0 + 224.0 * cr) * sc)
        out.append(np.concatenate([Y.reshape(n_frames, -1), U.reshape(n_frames, -1), V.reshape(n_frames, -1)], 1).astype(dtype))
    return out[0], out[1]
