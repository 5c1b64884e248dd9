"""Raw planar .yuv clips as frame providers.  Interface and names follow the reference's pyfvvdp/video_source_yuv.py:

  decode_video_props / create_yuv_fname   clip properties carried by the file name (:6-65)
  YUVReader                               memory-mapped reader of one file (:68-234)
  fvvdp_video_source_yuv_file             test / reference pair of .yuv files behind a display model (:238-302)

What differs is where the pixels are converted: the reference unpacks a frame with ~20 torch ops (fixed2float, bilinear
chroma upsampling, a 3x3 matmul, clip, EOTF, RGB2Y; video_source_yuv.py:157-228,299-302).  Here the three planes are
uploaded as they are in the file (uint8 / uint16) and ONE CUDA kernel (fvvdp_b200_yuv_to_luminance, include/fvvdp_b200.h)
turns them into the luminance frame the metric's kernels read.  There is no CPU fallback.
"""
import logging
import os
import re

import numpy as np
import torch

from . import _native
from .display_model import photometry_kernel_spec
from .video_source import fvvdp_video_source_dm

# display-encoded Y'CbCr -> R'G'B' (video_source_yuv.py:171-179)
YCBCR2RGB = {"2020": [1, 0, 1.47460, 1, -0.16455, -0.57135, 1, 1.88140, 0],
             "709": [1, 0, 1.402, 1, -0.344136, -0.714136, 1, 1.772, 0]}


def decode_video_props(fname):
    """Clip properties from a file name such as `clip_1920x1080_10b_420_2020_25fps.yuv`; defaults 1920x1080, 24 fps,
    8 bit, BT.2020, 4:2:0 (video_source_yuv.py:6-53)."""
    vprops = dict(width=1920, height=1080, fps=24, bit_depth=8, color_space="2020", chroma_ss="420")
    bname = os.path.splitext(os.path.basename(fname))[0]
    res_match = re.compile(r"(\d+)x(\d+)p?")
    for field in bname.split("_"):
        if res_match.match(field):
            res = field.split("x")
            if len(res) != 2:
                raise ValueError("Cannot decode the resolution")
            vprops["width"], vprops["height"] = int(res[0]), int(res[1].rstrip("p"))
            continue
        if field.endswith("fps"):
            vprops["fps"] = float(field[:-3])
        if field in ("444", "420"):
            vprops["chroma_ss"] = field
        if field in ("10", "10b"):
            vprops["bit_depth"] = 10
        if field in ("8", "8b"):
            vprops["bit_depth"] = 8
        if field in ("2020", "709"):
            vprops["color_space"] = field
        if field == "bt709":
            vprops["color_space"] = "709"
        if field in ("ct2020", "pq2020"):
            vprops["color_space"] = "2020"
    return vprops


def create_yuv_fname(basename, vprops):
    """File name that encodes the clip properties (video_source_yuv.py:56-65)."""
    fps = vprops["fps"]
    fps = round(fps, 3) if round(fps) != fps else int(fps)
    return f"{basename}_{vprops['width']}x{vprops['height']}_{vprops['bit_depth']}b_{vprops['chroma_ss']}_{vprops['color_space']}_{fps}fps.yuv"


class YUVReader:
    def __init__(self, file_name):
        self.file_name = file_name
        if not os.path.isfile(file_name):
            raise FileNotFoundError("File {} not found".format(file_name))
        vprops = decode_video_props(file_name)
        self.width, self.height, self.fps = vprops["width"], vprops["height"], vprops["fps"]
        self.color_space, self.chroma_ss, self.bit_depth = vprops["color_space"], vprops["chroma_ss"], vprops["bit_depth"]
        self.y_pixels = int(self.width * self.height)
        self.y_shape = (self.height, self.width)
        if self.chroma_ss == "444":
            self.uv_pixels, self.uv_shape = self.y_pixels, self.y_shape
        else:
            self.uv_pixels = int(self.y_pixels / 4)
            self.uv_shape = (int(self.height / 2), int(self.width / 2))
        self.frame_pixels = self.y_pixels + 2 * self.uv_pixels
        self.dtype = np.uint16 if self.bit_depth > 8 else np.uint8
        self.frame_bytes = self.frame_pixels * np.dtype(self.dtype).itemsize
        self.frame_count = int(os.stat(file_name).st_size / self.frame_bytes)
        self.mm = None

    def get_frame_count(self):
        return int(self.frame_count)

    def _planes(self, frame_index):
        if frame_index < 0 or frame_index >= self.frame_count:
            raise RuntimeError("The frame index is outside the range of available frames")
        if self.mm is None:  # mem-map as needed
            self.mm = np.memmap(self.file_name, self.dtype, mode="r")
        o = int(frame_index * self.frame_pixels)
        return (self.mm[o:o + self.y_pixels], self.mm[o + self.y_pixels:o + self.y_pixels + self.uv_pixels],
                self.mm[o + self.y_pixels + self.uv_pixels:o + self.y_pixels + 2 * self.uv_pixels])

    def get_frame_yuv(self, frame_index):
        Y, u, v = self._planes(frame_index)
        return np.reshape(Y, self.y_shape, "C"), np.reshape(u, self.uv_shape, "C"), np.reshape(v, self.uv_shape, "C")

    def _desc(self, spec=None, rgb2y=None, resize=None):
        """`resize` = (mode, (out_width, out_height)) of a full-screen resize, or None."""
        d = _native.YuvDesc()
        d.width, d.height, d.bit_depth = self.width, self.height, self.bit_depth
        d.chroma_420 = 1 if self.chroma_ss == "420" else 0
        for i, m in enumerate(YCBCR2RGB["2020" if self.color_space == "2020" else "709"]):
            d.ycbcr2rgb[i] = float(m)
        spec = spec or dict(kind="none")
        d.eotf = _native.EOTF_CODES[spec["kind"]]
        d.Y_peak, d.Y_black, d.gamma = spec.get("Y_peak", 0.0), spec.get("Y_black", 0.0), spec.get("gamma", 2.2)
        d.L_min, d.L_max = spec.get("L_min", 0.0), spec.get("L_max", 0.0)
        for i, w in enumerate(rgb2y or [0.0, 0.0, 0.0]):
            d.rgb2y[i] = float(w)
        if resize is not None:
            if resize[0] not in _native.RESIZE_CODES:
                raise ValueError(f"Unknown full_screen_resize mode '{resize[0]}' (nearest, bilinear, bicubic or area)")
            d.resize, d.out_width, d.out_height = _native.RESIZE_CODES[resize[0]], int(resize[1][0]), int(resize[1][1])
        return d

    def _convert(self, frame_index, device, spec, rgb2y, want_lum, want_rgb, resize=None):
        """Upload the frame's planes in their file layout and run the conversion (+ resize) kernel on `device`."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("fovvideovdp_b200 converts .yuv frames on CUDA devices only; there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        o = int(frame_index * self.frame_pixels)
        self._planes(frame_index)  # range check + mem-map
        raw = np.array(self.mm[o:o + self.frame_pixels])  # one contiguous, writable copy of the frame
        host = torch.from_numpy(raw.view(np.int16) if self.dtype == np.uint16 else raw)  # the bit pattern is what travels
        with torch.cuda.device(device):
            dev = host.to(device, non_blocking=True)
            esz = dev.element_size()
            ow, oh = (self.width, self.height) if resize is None else (int(resize[1][0]), int(resize[1][1]))
            lum = torch.empty((1, 1, 1, oh, ow), dtype=torch.float32, device=device) if want_lum else None
            rgb = torch.empty((oh, ow, 3), dtype=torch.float32, device=device) if want_rgb else None
            _native.yuv_to_luminance(self._desc(spec, rgb2y, resize), dev.data_ptr(), dev.data_ptr() + self.y_pixels * esz,
                                     dev.data_ptr() + (self.y_pixels + self.uv_pixels) * esz, lum.data_ptr() if want_lum else 0,
                                     rgb.data_ptr() if want_rgb else 0, device.index, torch.cuda.current_stream(device).cuda_stream)
            dev.record_stream(torch.cuda.current_stream(device))
        return lum, rgb

    def get_frame_rgb_tensor(self, frame_index, device, resize=None):
        """Display-encoded RGB (H,W,3) float32 in [0,1] on `device` (video_source_yuv.py:157-182); with `resize` = (mode,
        (width, height)) resampled like torch.nn.functional.interpolate(mode=...) and clipped (:293-297), in the same kernel."""
        return self._convert(frame_index, device, None, None, False, True, resize)[1]

    def get_frame_luminance(self, frame_index, device, spec, rgb2y, resize=None):
        """Luminance (1,1,1,H,W) float32 in cd/m^2 for a stock photometry `spec` (display_model.photometry_kernel_spec)."""
        return self._convert(frame_index, device, spec, rgb2y, True, False, resize)[0]

    def __enter__(self):
        return self

    def __exit__(self, type, value, tb):
        self.mm = None


class fvvdp_video_source_yuv_file(fvvdp_video_source_dm):
    def __init__(self, test_fname, reference_fname, display_photometry="standard_4k", color_space_name="auto", frames=-1,
                 full_screen_resize=None, resize_resolution=None, verbose=False):
        self.reference_vidr = YUVReader(reference_fname)
        self.test_vidr = YUVReader(test_fname)
        self.frames = self.test_vidr.frame_count if frames == -1 else min(self.test_vidr.frame_count, frames)
        self.full_screen_resize = full_screen_resize
        self.resize_resolution = resize_resolution
        if color_space_name == "auto":
            color_space_name = "BT.2020" if self.test_vidr.color_space == "2020" else "sRGB"
        super().__init__(display_photometry=display_photometry, color_space_name=color_space_name)
        self._spec = photometry_kernel_spec(self.dm_photometry)  # None: custom photometry, applied through its forward()
        for vr, what in ((self.test_vidr, "Test"), (self.reference_vidr, "Reference")):
            rs = "" if full_screen_resize is None else f"->[{resize_resolution[0]}x{resize_resolution[1]}]"
            logging.debug(f"{what} video '{vr.file_name}': [{vr.width}x{vr.height}]{rs}, colorspace: {vr.color_space}, fps: {vr.fps}, "
                          f"{vr.bit_depth} bit {vr.chroma_ss}, frames: {self.frames}")

    def get_video_size(self):
        if self.full_screen_resize is not None:
            return [self.resize_resolution[1], self.resize_resolution[0], self.frames]
        return [self.test_vidr.height, self.test_vidr.width, self.frames]

    def get_frames_per_second(self):
        return self.test_vidr.fps

    def resize_of(self, vid_reader):
        """(mode, (width, height)) when this reader's frames are resampled to the display resolution (video_source_yuv.py:293), else None."""
        if self.full_screen_resize is None or (vid_reader.height == self.resize_resolution[1] and vid_reader.width == self.resize_resolution[0]):
            return None
        return (self.full_screen_resize, (int(self.resize_resolution[0]), int(self.resize_resolution[1])))

    def get_test_frame(self, frame, device):
        return self._get_frame(self.test_vidr, frame, device)

    def get_reference_frame(self, frame, device):
        return self._get_frame(self.reference_vidr, frame, device)

    def _get_frame(self, vid_reader, frame, device):
        resize = self.resize_of(vid_reader)
        if self._spec is not None:
            return vid_reader.get_frame_luminance(frame, device, self._spec, self.color_to_luminance, resize)
        # custom photometry plugins: (resized) RGB from the kernel, the plugin's forward() as the reference does it (:290-302)
        RGB = vid_reader.get_frame_rgb_tensor(frame, device, resize).permute(2, 0, 1)[None]
        RGB_lin = self.dm_photometry.forward(RGB[:, :, None])
        w = self.color_to_luminance
        return RGB_lin[:, 0:1] * w[0] + RGB_lin[:, 1:2] * w[1] + RGB_lin[:, 2:3] * w[2]
