"""PU21-PSNR, the second metric of the reference's command line (pyfvvdp/pupsnr.py, utils.PU: utils.py:157-202):
both luminance frames are encoded with the perceptually uniform PU21 transfer function and compared by PSNR, frame by
frame.  The per-frame sum of squared differences is one CUDA kernel (fvvdp_b200_pu_sq_err, include/fvvdp_b200.h); there is
no CPU fallback.  Interface = the reference's `pu_psnr` class; `predict()` takes the display model explicitly because
the reference's relies on attributes its constructor never sets (pupsnr.py:43)."""
import math

import torch

from . import _native
from .display_model import fvvdp_display_photometry
from .display_model import photometry_kernel_spec
from .fvvdp import _DTYPES, _FrameSet
from .video_source import fvvdp_video_source_array, is_array_source

PU21_BANDING_GLARE = [234.0235618, 216.9339286, 0.0001091864237, 0.893206924, 0.06733984121, 1.444718567, 567.6315065]  # utils.py:175


class pu_psnr:
    def __init__(self, device=None, display_name="standard_4k", display_photometry=None, color_space="sRGB"):
        if device is None:
            if not (torch.cuda.is_available() and torch.cuda.device_count() > 0):
                raise RuntimeError("fovvideovdp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda:0")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"fovvideovdp_b200 runs on CUDA devices only (got '{device}'); there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        _native.load_library()
        self.display_photometry = fvvdp_display_photometry.load(display_name) if display_photometry is None else display_photometry
        self.color_space = color_space
        self.L_min, self.L_max = 0.005, 10000.0
        p = PU21_BANDING_GLARE
        self.peak = p[6] * (((p[0] + p[1] * self.L_max ** p[3]) / (1 + p[2] * self.L_max ** p[3])) ** p[4] - p[5])  # utils.py:185
        self._params = _native.PuParams()
        for i, v in enumerate(p):
            self._params.p[i] = v
        self._params.L_min, self._params.L_max = self.L_min, self.L_max

    def predict(self, test_cont, reference_cont, dim_order="BCFHW", frames_per_second=0, fixation_point=None, frame_padding="replicate"):
        vs = fvvdp_video_source_array(test_cont, reference_cont, frames_per_second, dim_order=dim_order,
                                      display_photometry=self.display_photometry, color_space_name=self.color_space)
        return self.predict_video_source(vs, fixation_point=fixation_point, frame_padding=frame_padding)

    def predict_video_source(self, vid_source, fixation_point=None, frame_padding="replicate"):
        height, width, N_frames = vid_source.get_video_size()
        N_frames = int(N_frames)
        dev = self.device
        spec = photometry_kernel_spec(vid_source.dm_photometry) if is_array_source(vid_source) else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            acc = torch.zeros(N_frames, dtype=torch.float64, device=dev)
            n = 0
            if spec is not None:
                # array source with a stock display model: whole blocks of frames in one launch, read in their own dtype and
                # layout (EOTF, RGB -> Y and PU21 in the kernel; no luminance frames, no per-frame torch arithmetic)
                frames = _FrameSet(vid_source, dev, True)
                fmt = _native.FrameFormat()
                fmt.width, fmt.height = int(width), int(height)
                fmt.in_dtype, fmt.in_channels = _DTYPES[frames.dtype], 3 if vid_source.is_color else 1
                fmt.eotf = _native.EOTF_CODES[spec["kind"]]
                fmt.Y_peak, fmt.Y_black = spec.get("Y_peak", 0.0), spec.get("Y_black", 0.0)
                fmt.gamma, fmt.L_min, fmt.L_max = spec.get("gamma", 2.2), spec.get("L_min", 0.0), spec.get("L_max", 0.0)
                w = vid_source.color_to_luminance if vid_source.is_color else [1.0, 0.0, 0.0]
                for i in range(3):
                    fmt.rgb2y[i] = float(w[i])
                first = getattr(vid_source, "first_frame", 0)
                B = 32
                for f0 in range(0, N_frames, B):
                    idx = list(range(f0, min(N_frames, f0 + B)))
                    frames.fetch_all(idx)
                    ready = frames.uploaded()
                    if ready is not None:
                        torch.cuda.current_stream(dev).wait_event(ready)
                    tp, rp = frames.block_pointers(idx)
                    _native.pu_sq_err_frames(fmt, tp, rp, frames.strides, self._params, acc.data_ptr() + 8 * f0, dev.index, stream)
                    scored = None
                    if frames.up_stream is not None:
                        scored = torch.cuda.Event()
                        scored.record(torch.cuda.current_stream(dev))
                    frames.retain_only(set(), scored)
                n = int(height) * int(width)
                N_loop = 0
            else:
                N_loop = N_frames
            for ff in range(N_loop):
                T = vid_source.get_test_frame(ff, device=dev).to(device=dev, dtype=torch.float32).contiguous()
                R = vid_source.get_reference_frame(ff, device=dev).to(device=dev, dtype=torch.float32).contiguous()
                n = T.numel()
                _native.pu_sq_err(T.data_ptr(), R.data_ptr(), n, self._params, acc.data_ptr() + 8 * ff, dev.index, stream)
                T.record_stream(torch.cuda.current_stream(dev))
                R.record_stream(torch.cuda.current_stream(dev))
            sq = acc.cpu().numpy()  # one device->host read
        # pupsnr.py:66-79; identical frames (zero error) give +inf like the reference's torch arithmetic, not a ZeroDivisionError
        psnr = sum((20.0 * math.log10(self.peak / math.sqrt(float(s) / n)) if s > 0 else math.inf) for s in sq) / N_frames
        return torch.tensor(psnr, dtype=torch.float32, device=dev), None

    def short_name(self):
        return "PU21-PSNR"

    def quality_unit(self):
        return "dB"

    def get_info_string(self):
        return None
